/* distance3d_b200 -- C ABI of the CUDA library (libd3d_b200.so).
 *
 * The reference (AlexanderFabisch/distance3d, pure Python + numba) has no FFI
 * seam; the seam a maintainer would bind is its Python API.  Each entry point
 * below names the reference function it replaces (file:line relative to the
 * reference repository).  Conventions:
 *   - every pointer is a DEVICE pointer unless marked "host";
 *   - the caller owns all buffers (inputs, outputs, workspaces);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*), no
 *     host synchronisation happens inside a call;
 *   - return value 0 = ok, < 0 = API misuse / CUDA error, text available from
 *     d3d_last_error_string(); data-dependent outcomes are per-pair status
 *     codes (include/d3d_types.h).
 * INTEGRATION.md shows the ctypes binding used by distance3d_b200/_lib.py.
 */
#ifndef D3D_B200_H
#define D3D_B200_H

#include <stddef.h>
#include "d3d_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* host: message of the last failing call on this host thread */
const char *d3d_last_error_string(void);
/* host: SM count and compute capability of the current device */
int d3d_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* Measurement / test hooks (no reference counterpart):
 * d3d_fp64_peak_probe runs blocks x 256 threads x 8*iters dependent-free FP64 FMAs
 * (roofline denominator for the narrow phase, timed by the caller with CUDA events);
 * d3d_debug_norm evaluates the device emulation of the reference's np.linalg.norm
 * (BLAS dnrm2 on the x87 FPU; mode 0 = production path, 1 = integer emulation only). */
int d3d_fp64_peak_probe(double *scratch, int blocks, int iters, void *stream);
int d3d_debug_norm(const double *v, int64_t n, double *out, int mode, void *stream);
/* out[k,:] = v[k,:] / s[k] through the device's vector division (three IEEE quotients that
 * share one reciprocal refinement); tests pin it to host division bit for bit. */
int d3d_debug_vdiv(const double *v, const double *s, int64_t n, double *out, void *stream);

/* geometry.py:138-157 convert_box_to_vertices (called from colliders.py:161-176 on
 * construction and update_pose): writes the 8 world-frame vertices of every BOX of
 * `c` into verts_out[vert_off .. vert_off+8). */
int d3d_prepare(const d3d_colliders *c, double *verts_out, void *stream);

/* Compact wire format for collider sets that arrive over PCIe (stream.GjkDistanceStream): the
 * reference's objects carry a 4x4 pose per collider (colliders.py:243-646, 128 B) although a
 * sphere is four numbers.  One record of doubles per collider at wire[wire_off[i]], layout by
 * wire_type[i] (= the type tag):
 *   sphere 4: centre, radius                      capsule / cylinder / cone 14: pose rows 0-2, r, h
 *   ellipsoid 15: pose rows 0-2, radii            box 16: pose rows 0-2, size, {vert_off, vert_len}
 *   hull 1: {vert_off, vert_len}                  mesh 13: pose rows 0-2, {vert_off, vert_len}
 *   disk 7: centre, normal, radius                ellipse 11: centre, axis 0, axis 1, radii
 * ({a, b} = two int32 in one 8-byte slot).  99 B per collider on the five-primitive mix instead of
 * 164 B.  Expands the records into the d3d_colliders arrays (type, pose, param, vert_off,
 * vert_len; margin / MeshGraph fields are passed to d3d_colliders as they are). */
int d3d_unpack_colliders(const uint8_t *wire_type, const int32_t *wire_off, const double *wire,
                         int64_t n, int32_t *type, double *pose, double *param, int32_t *vert_off,
                         int32_t *vert_len, void *stream);

/* wire_off[i] = sum of the record sizes of colliders 0 .. i-1, computed on the device from the
 * types alone: the offsets d3d_pack_wire_host wrote on the host do not have to be uploaded
 * (4 of ~104 bytes per collider on the primitive mix). */
int d3d_wire_offsets(const uint8_t *wire_type, int64_t n, int32_t *wire_off, void *stream);
/* HOST functions (no CUDA call): number of doubles the records of `type[n]` take (-1: unknown
 * type), and the packer itself - `c` holds HOST pointers, the three outputs are host buffers
 * (pinned for the upload) with n, n and d3d_wire_size() entries; n_threads host threads. */
int64_t d3d_wire_size(const int32_t *type, int64_t n);
int d3d_pack_wire_host(const d3d_colliders *c, uint8_t *wire_type, int32_t *wire_off, double *wire,
                       int n_threads);

/* colliders.py:131,221,272,326,374,426,479,533,590,629 <collider>.support_function(d):
 * out[k,:] = support of collider idx[k] in direction dirs[k,:]. */
int d3d_support(const d3d_colliders *c, const int32_t *idx, const double *dirs, int64_t n,
                double *out, void *stream);

/* colliders.py:134,171,224,266,319,367,419,472,527,582 <collider>.center(): out[n,3] */
int d3d_center(const d3d_colliders *c, double *out, void *stream);

/* colliders.py:140,180,234,281,335,383,435,489,543,599,639 <collider>.aabb()
 * (containment.py:6-229): out[n,3,2] = [[xmin,xmax],[ymin,ymax],[zmin,zmax]]. */
int d3d_aabb(const d3d_colliders *c, double *out, void *stream);

/* Workspace sizes over n_pairs pairs.  d3d_gjk_distance needs ~333 B per pair (the pair
 * order by collider types plus the parked final simplex of every pair), d3d_gjk_intersection
 * 5 B per pair.  The larger size is accepted by both calls. */
size_t d3d_gjk_workspace_bytes(int64_t n_pairs);
size_t d3d_gjk_intersection_workspace_bytes(int64_t n_pairs);

/* gjk/_gjk_jolt.py:138-221 gjk_distance_jolt (= gjk.gjk, gjk/__init__.py:27) for
 * pairs[k] = (index of collider 1, index of collider 2):
 *   out_dist[k]      distance (MAX_FLOAT when status is CLIPPED)
 *   out_a/out_b[k,3] closest points on collider 1 / 2          (may be NULL)
 *   out_Y[k,4,3]     simplex of Minkowski-difference points    (may be NULL)
 *   out_npoints[k]   number of valid rows of out_Y             (may be NULL)
 *   out_iters[k]     GJK iterations (gjk_distance_jolt_iterations, :714-785; may be NULL)
 *   out_status[k]    D3D_NO_INTERSECTION / D3D_INTERSECTION / D3D_CLIPPED / ... (may be NULL) */
int d3d_gjk_distance(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs,
                     double tolerance, double max_distance_squared, double sanity_check,
                     double *out_dist, double *out_a, double *out_b, double *out_Y,
                     int32_t *out_npoints, int32_t *out_iters, int32_t *out_status,
                     void *workspace, size_t ws_bytes, void *stream);

/* gjk/_gjk_jolt.py:29-135 gjk_intersection_jolt (= gjk.gjk_intersection): out_hit[k] in {0,1} */
int d3d_gjk_intersection(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs,
                         double tolerance, uint8_t *out_hit, int32_t *out_iters,
                         int32_t *out_status, void *workspace, size_t ws_bytes, void *stream);

/* Opt-in fp32 arithmetic (BASELINE north_star): same arguments and buffer types (fp64 in HBM)
 * as d3d_gjk_distance / d3d_gjk_intersection, all arithmetic in single precision.  Not
 * bit-compatible with the reference; tolerance against the fp64 path (unit-scale shapes,
 * tests/test_gjk_gpu.py): 99.9 % of the distances within 1e-4, every result an upper bound at
 * most 0.2 off (rare early termination on polytope-vs-curved pairs, as in Jolt's own
 * single-precision GJK whose degeneracy thresholds this mode uses), intersection flags equal
 * where the fp64 distance exceeds 1e-3. */
int d3d_gjk_distance_f32(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs,
                         double tolerance, double max_distance_squared, double sanity_check,
                         double *out_dist, double *out_a, double *out_b, double *out_Y,
                         int32_t *out_npoints, int32_t *out_iters, int32_t *out_status,
                         void *workspace, size_t ws_bytes, void *stream);
int d3d_gjk_intersection_f32(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs,
                             double tolerance, uint8_t *out_hit, int32_t *out_iters,
                             int32_t *out_status, void *workspace, size_t ws_bytes, void *stream);

/* gjk/_gjk_libccd.py:14-91 gjk_intersection_libccd(collider1, collider2, max_iterations): the
 * reference's second, libccd-style boolean GJK (:112-266 simplex cases), which its tests use as
 * an independent cross-check of the Jolt variant (test/test_gjk.py:341-354).  out_hit[k] = 1 when
 * the pair intersects; out_iters (optional) = support evaluations.  `perm` (optional,
 * int32[n_pairs]): processing order, e.g. pairs grouped by collider types. */
int d3d_gjk_intersection_libccd(const d3d_colliders *c, const int32_t *pairs, const int32_t *perm,
                                int64_t n_pairs, int max_iterations, uint8_t *out_hit,
                                int32_t *out_iters, void *stream);

/* epa.py:9-78 epa(simplex, collider1, collider2, max_iter, max_loose_edges, max_faces, epsilon)
 * for pairs[k] with GJK simplex Y[k,4,3] (out_Y of d3d_gjk_distance).  The reference reads all
 * four rows of the simplex although GJK may have ended with fewer valid points (its result is
 * then undefined, rows of np.empty); npoints[k] (out_npoints of d3d_gjk_distance, may be NULL
 * = all 4) marks those pairs: they are skipped with status D3D_EPA_BAD_SIMPLEX, mtv 0.
 *   out_mtv[k,3]   minimum translation vector (depth = |mtv|, normal = mtv / |mtv|)
 *   out_success[k] 1 = converged before max_iter
 *   out_nfaces[k], out_iters[k] (may be NULL); out_faces[k,max_faces,4,3] (may be NULL)
 *   out_status[k]  D3D_INTERSECTION, or D3D_EPA_MAX_FACES where the reference raises
 *                  AssertionError (epa.py:128)
 * Limits: 4 <= max_faces <= 64, 1 <= max_loose_edges <= 32 (the reference's defaults).
 * For status D3D_EPA_MAX_FACES only the status is defined (mtv 0, success 0).
 * Two kernels with identical results: a thread-per-pair kernel (default limits, out_faces NULL,
 * n_pairs >= 20000) followed by the warp-per-pair kernel on what it hands over, or the warp
 * kernel alone; the environment variable D3D_EPA_KERNEL=warp|thread overrides the choice.
 * Workspace: 9 bytes per pair + 3.9 KB per thread of the thread kernel (<= 298 MB). */
size_t d3d_epa_workspace_bytes(int64_t n_pairs);
int d3d_epa(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs, const double *Y,
            const int32_t *npoints, int max_iter, int max_loose_edges, int max_faces, double epsilon, double *out_mtv,
            uint8_t *out_success, int32_t *out_nfaces, int32_t *out_iters, int32_t *out_status,
            double *out_faces, void *workspace, size_t ws_bytes, void *stream);

/* mpr.py:21-50 mpr_intersection (want_penetration = 0) and mpr.py:53-109 mpr_penetration
 * (want_penetration = 1) for pairs[k]:
 *   out_hit[k]     1 = intersecting
 *   out_depth[k], out_dir[k,3], out_pos[k,3]  penetration depth, direction (adding
 *                  depth * dir to collider 2 separates the pair) and contact position;
 *                  zero where out_hit is 0 (the reference returns None there)
 *   out_status[k]  D3D_NO_INTERSECTION / D3D_INTERSECTION / D3D_ITER_CAP (may be NULL)
 * `perm` (optional, int32[n_pairs]): processing order, e.g. pairs grouped by collider types. */
int d3d_mpr(const d3d_colliders *c, const int32_t *pairs, const int32_t *perm, int64_t n_pairs,
            double tolerance, int max_iterations, int want_penetration, uint8_t *out_hit,
            double *out_depth, double *out_dir, double *out_pos, int32_t *out_status,
            void *stream);

/* ---- broad phase ---------------------------------------------------------- */

/* aabb_tree.py:465-500 all_aabbs_overlap (brute force): appends every (i, j) with
 * aabb_overlap(aabbs1[i], aabbs2[j]) (closed intervals, :503-527) to out_pairs[cap,2];
 * *out_count (device, 64-bit) is exact even if it exceeds cap (then re-run with more room).
 * Order of the list is unspecified (the reference's is row-major). */
int d3d_aabb_overlap_brute(const double *aabb1, int64_t n1, const double *aabb2, int64_t n2,
                           int32_t *out_pairs, int64_t cap, unsigned long long *out_count,
                           void *stream);

/* Workspace (tree storage + sort scratch) for a BVH over n boxes. */
size_t d3d_bvh_workspace_bytes(int64_t n);

/* aabb_tree.py:31-101 AabbTree.insert_aabbs + :194-341 insert_aabbs / insert_leaf /
 * fix_upward_tree: builds the tree over aabb[n,3,2] into the caller's workspace (the
 * workspace IS the tree handle).  The tree is an LBVH, not the reference's insertion tree:
 * only the overlap sets are contractual. */
int d3d_bvh_build(const double *aabb, int64_t n, void *workspace, size_t ws_bytes, void *stream);

/* aabb_tree.py:161-181 AabbTree.overlaps_aabb (:381-403 query_overlap) for n_query boxes and
 * :121-159 overlaps_aabb_tree (:344-378 query_overlap_of_other_tree) when the query boxes are
 * the leaves of another tree: every overlapping (tree object index, query index) pair.
 *   d3d_bvh_overlap_count  pass 1: counts per query + exclusive scan into the query workspace;
 *                          *out_count (device, 64-bit) = exact number of pairs
 *   d3d_bvh_overlap_fill   pass 2: writes out_pairs[cap,2] (positions >= cap are dropped), pairs
 *                          grouped by query in processing order, leaves in depth-first order
 *   d3d_bvh_overlap_ordered  both passes in one call
 *   d3d_bvh_overlap        ONE traversal: hits are staged per warp and appended to out_pairs with
 *                          one reservation per 512 pairs; the same set of pairs in an order that
 *                          depends on the scheduling; *out_count is exact even when cap is too
 *                          small (the excess is dropped); query_ws may be NULL; here packet may
 *                          also be an explicit group width 2, 4, 8, 16, 32 (1 = the default, 8)
 * `order` (optional, int32[n_query]) = processing order of the queries (spatially sorted queries
 * traverse coherently).  packet = 1: the 32 queries of a warp traverse together (node fetched
 * once per warp; use with spatially sorted queries, best on dense scenes), packet = 0: one
 * independent traversal per thread.  Identical results.  *out_visits (optional) = node
 * records fetched (measurement). */
size_t d3d_bvh_query_workspace_bytes(int64_t n_query);
int d3d_bvh_overlap_count(const void *workspace, int64_t n, const double *query, const int32_t *order,
                          int64_t n_query, int packet, unsigned long long *out_count,
                          unsigned long long *out_visits, void *query_ws, size_t query_ws_size,
                          void *stream);
int d3d_bvh_overlap_fill(const void *workspace, int64_t n, const double *query, const int32_t *order,
                         int64_t n_query, int packet, int32_t *out_pairs, int64_t cap,
                         const void *query_ws, void *stream);
int d3d_bvh_overlap(const void *workspace, int64_t n, const double *query, const int32_t *order,
                    int64_t n_query, int packet, int32_t *out_pairs, int64_t cap,
                    unsigned long long *out_count, unsigned long long *out_visits, void *query_ws,
                    size_t query_ws_size, void *stream);
int d3d_bvh_overlap_ordered(const void *workspace, int64_t n, const double *query, const int32_t *order,
                            int64_t n_query, int packet, int32_t *out_pairs, int64_t cap,
                            unsigned long long *out_count, unsigned long long *out_visits,
                            void *query_ws, size_t query_ws_size, void *stream);

/* aabb_tree.py:121-159 overlaps_aabb_tree(self) / broad_phase.py:229-252 aabb_overlapping_with_self
 * without the redundancy: the tree against its own leaves, ONE traversal.  Every unordered
 * overlapping pair is appended once as (smaller, larger) object index; (i, i) is not reported.
 * A leaf only walks the part of the tree behind itself in Morton order, so nodes visited and
 * bytes written are half of d3d_bvh_overlap over the same boxes.  part / n_parts split the work
 * between GPUs that hold replicas of the tree: the leaves are dealt in blocks of 128 of the
 * Morton order, round-robin (block b goes to part b % n_parts), which balances pairs and
 * traversal cost; the parts' lists are disjoint and their union is the full set.  One GPU:
 * part = 0, n_parts = 1. */
int d3d_bvh_overlap_self(const void *workspace, int64_t n, int part, int n_parts, int packet,
                         int32_t *out_pairs, int64_t cap, unsigned long long *out_count,
                         unsigned long long *out_visits, void *stream);

/* d3d_bvh_overlap_self FUSED with the all-gather of the ranks' pair lists (the one collective of
 * the multi-GPU design: "all-gather variable-length contact and result lists"): the traversal
 * kernel itself stores every flushed chunk of pairs into the pair buffer of EVERY GPU through
 * peer pointers over NVLink / NVSwitch, so the transfer overlaps the walk chunk by chunk and no
 * separate collective runs afterwards.  peer_pairs: DEVICE array of n_parts pointers to the
 * int32[n_parts * segment_cap, 2] buffers of all ranks (a symmetric allocation; entry `part` is
 * this GPU's own); rank r's pairs land in rows [r * segment_cap, r * segment_cap + count_r) of
 * every buffer.  peer_counts: device array of n_parts pointers to uint64[n_parts] arrays; entry
 * [p][r] receives rank r's exact count (pairs beyond segment_cap are dropped).  local_count:
 * this rank's cursor.  The caller synchronises the ranks afterwards (a barrier on the stream,
 * e.g. torch symmetric memory's) before reading other ranks' segments. */
int d3d_bvh_overlap_self_gather(const void *workspace, int64_t n, int part, int n_parts,
                                int32_t *const *peer_pairs, int64_t segment_cap,
                                unsigned long long *const *peer_counts, unsigned long long *local_count,
                                unsigned long long *out_visits, void *stream);

/* Morton order of the tree's objects: out[j] = object index of sorted leaf j. */
int d3d_bvh_leaf_order(const void *workspace, int64_t n, int32_t *out, void *stream);

/* aabb_tree.py:183-191 AabbTree.get_root_aabb: out[3,2] */
int d3d_bvh_root_aabb(const void *workspace, int64_t n, double *out, void *stream);

/* ---- hydroelastic contact: the other consumer of the AABB broad phase -------- */

/* hydroelastic_contact/_mesh_processing.py:4-20 tetrahedral_mesh_aabbs:
 * points[n,4,3] -> out_aabb[n,3,2] (feeds d3d_bvh_build / d3d_bvh_overlap, which replace
 * rigid_body.aabbtree_.overlaps_aabb_tree / all_aabbs_overlap in
 * hydroelastic_contact/_interface.py:76-80). */
int d3d_tetra_aabb(const double *points, int64_t n, double *out_aabb, void *stream);

/* hydroelastic_contact/_barycentric_transform.py:4-9 barycentric_transforms:
 * points[n,4,3] -> out_X[n,4,4], the inverse of [[p0 p1 p2 p3], [1 1 1 1]] in closed form (the
 * reference calls numpy's pinv); row i = barycentric coordinate function of vertex i. */
int d3d_tetra_barycentric(const double *points, int64_t n, double *out_X, void *stream);

/* hydroelastic_contact/_tetrahedron_intersection.py:7-140 intersect_tetrahedron_pairs for the
 * candidate pairs[k] = (tetrahedron of mesh 1, tetrahedron of mesh 2), both meshes expressed in
 * one frame (points[n,4,3], vertex potentials eps[n,4]); X1 / X2 [n,4,4] are optional (NULL: the
 * barycentric transforms are computed per pair on the fly).
 *   out_hit[k]        1 = the tetrahedra intersect (contact polygon with >= 3 vertices)
 *   out_plane[k,4]    contact plane in Hesse normal form (:165-216)
 *   out_nverts[k], out_poly[k,max_vertices,3]  contact polygon, counter-clockwise (:377-423)
 *   out_status[k]     0, 1 = "same tetrahedron" branch (:132-134, 143-162), 2 = polygon had more
 *                     than max_vertices vertices (excess dropped); may be NULL
 * 3 <= max_vertices <= 24.  Tolerance contract: planes and polygon regions within 1e-9 of the
 * reference's (its pinv / atan2 are not reproduced bit for bit); where a face normal is parallel
 * to the contact plane normal the reference reads an uninitialised half-plane row
 * (:279-292 indexes by i, not hp_idx) - here the remaining half-planes are used. */
int d3d_tetra_intersect_pairs(const int32_t *pairs, int64_t n_pairs, const double *points1,
                              const double *eps1, const double *X1, const double *points2,
                              const double *eps2, const double *X2, double youngs_modulus1,
                              double youngs_modulus2, int max_vertices, uint8_t *out_hit,
                              double *out_plane, int32_t *out_nverts, double *out_poly,
                              int32_t *out_status, void *stream);

/* ---- robot self-collision (BASELINE config 4) ------------------------------ */

/* Batched forward kinematics; replaces the per-configuration pytransform3d calls
 * UrdfTransformManager.set_joint + get_transform(frame, "origin") made by
 * broad_phase.py:111,148.  The kinematic tree is flattened by
 * distance3d_b200.urdf.UrdfTransformManager.compile_kinematics: for frame k the pose is
 * prod_{s in [chain_off[k], chain_off[k+1])} chain_fixed[s] * joint(chain_joint[s], q)
 * (revolute: Rodrigues rotation about joint_axis, value clipped to joint_limits).
 * q[n_cfg, n_joints] -> out_pose[n_cfg, n_frames, 4, 4] (row-major). */
int d3d_fk_urdf(int n_frames, int n_joints, const double *joint_axis, const double *joint_limits,
                const int32_t *joint_type, const int32_t *chain_off, const double *chain_fixed,
                const int32_t *chain_joint, const double *q, int64_t n_cfg, double *out_pose,
                void *stream);

/* The same poses, bit for bit, with the common prefixes of the frames' chains evaluated once per
 * configuration (one thread per configuration instead of one per (configuration, frame)).  The
 * chain steps form a tree: node i = (node_parent[i] or -1 for the base, node_fixed[i], node_joint[i]),
 * parents before children; node_keep[i] = slot (0 .. n_keep-1, n_keep <= 16) of a node that has
 * children, -1 otherwise; frames node_out[node_out_off[i] .. node_out_off[i+1]) end at node i.
 * compile_kinematics emits these arrays next to the flat chains. */
int d3d_fk_urdf_tree(int n_frames, int n_joints, const double *joint_axis, const double *joint_limits,
                     const int32_t *joint_type, int n_nodes, int n_keep, const int32_t *node_parent,
                     const double *node_fixed, const int32_t *node_joint, const int32_t *node_keep,
                     const int32_t *node_out_off, const int32_t *node_out, const double *q, int64_t n_cfg,
                     double *out_pose, void *stream);

/* self_collision.py:22-36: candidates of a collider = AABB overlaps minus white-list.  The
 * non-white-listed (frame a, frame b) combinations are a fixed `pattern[n_pattern,2]`; for
 * each of n_groups groups of group_size consecutive boxes the overlapping candidate pairs
 * (global box indices) are appended to out_pairs[cap,2]; *out_count is exact. */
int d3d_filter_pairs(const double *aabb, int64_t n_groups, int group_size, const int32_t *pattern,
                     int n_pattern, int32_t *out_pairs, int64_t cap, unsigned long long *out_count,
                     void *stream);

/* d3d_aabb + d3d_filter_pairs in one pass (group_size <= 128): the boxes of a group are computed,
 * stored to out_aabb[n,3,2] (may be NULL when nobody reads them afterwards) and tested from shared
 * memory.  Same pairs as the two calls, in a different order. */
int d3d_aabb_filter_pairs(const d3d_colliders *c, int64_t n_groups, int group_size, const int32_t *pattern,
                          int n_pattern, double *out_aabb, int32_t *out_pairs, int64_t cap,
                          unsigned long long *out_count, void *stream);

/* self_collision.py:31-35: mask[pairs[t,0]] = mask[pairs[t,1]] = 1 for every t < min(*count, cap)
 * with hit[t] != 0. */
int d3d_scatter_hits(const int32_t *pairs, const uint8_t *hit, const unsigned long long *count,
                     int64_t cap, uint8_t *mask, void *stream);

/* self_collision.py:22-36 with the reference's candidate ORDER, for white-lists that are not
 * symmetric (urdf_utils.py:79-81): per group the reference's incremental AABB tree
 * (aabb_tree.py:194-341, one insert per collider as in broad_phase.py:144-151) is rebuilt
 * over the group's boxes and the detect loop is replayed - frames in order, candidates in the
 * tree's depth-first query order (aabb_tree.py:381-403) minus whitelist[frame], the first
 * intersecting candidate flags both frames.  pairs/hit/count: output of d3d_filter_pairs and
 * d3d_gjk_intersection over every pair that is not white-listed in BOTH directions;
 * whitelist[group_size]: bit j of entry i = frame j is white-listed for frame i;
 * hit_bits[n_groups*group_size]: scratch; mask[n_groups*group_size]: output.
 * group_size <= 64. */
int d3d_detect_ordered(const double *aabb, int64_t n_groups, int group_size, const int32_t *pairs,
                       const uint8_t *hit, const unsigned long long *count, int64_t cap,
                       const unsigned long long *whitelist, unsigned long long *hit_bits,
                       uint8_t *mask, void *stream);

#ifdef __cplusplus
}
#endif
#endif
