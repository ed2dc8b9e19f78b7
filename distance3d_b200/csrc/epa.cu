// Expanding polytope algorithm, one WARP per intersecting pair.
//
// Replaces distance3d/epa.py:9-202 (epa, Polytope, LooseEdges), which is plain
// interpreted Python in the reference.  The polytope (up to max_faces faces = 3
// vertices + unit normal each) and the loose-edge list live in shared memory in
// structure-of-arrays form ([12][max_faces], conflict free for lane-per-face access); the
// loose-edge list is held in registers, entry e in lane e
// (max_faces <= 64, max_loose_edges <= 32: one or two lanes-worth.)  Per iteration the warp
//   A  finds the face closest to the origin (lane-strided scan + shuffle arg-min,
//      lowest index wins ties like np.argmin),
//   B  evaluates the support point in the face normal (hull vertex scans are
//      cooperative, analytic supports replicated),
//   C  tests convergence,
//   D  marks the faces that see the new point (lane per face), replays the
//      reference's swap-with-last removal loop on a slot permutation, and maintains
//      the loose-edge list with a lane-parallel search per edge (first match wins; a matched
//      entry is overwritten by the last one with six shuffles),
//   E  builds the new faces lane-per-edge and compacts the valid ones in order.
// The reference's quirks are reproduced on purpose: the
// "swap" in fix_ccw_normal_direction only copies v1 over v0, degenerate new faces
// are skipped but leave their data behind, the max_faces assertion becomes status
// D3D_EPA_MAX_FACES, and after max_iter iterations the result is read from the slot
// that held the last closest face as it looks THEN.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "d3d_common.cuh"
#include "d3d_support.cuh"

namespace {

struct EpaParams {
    int max_iter, max_loose_edges, max_faces;
    double epsilon;
    double eps_sq_thr;   // smallest t with sqrt(t) >= epsilon:  sqrt(s) < epsilon  <=>  s < t
    double half_sq_thr;  // same for 0.5
    const double *Y;
    const int32_t *npoints;
    double *out_mtv;
    uint8_t *out_success;
    int32_t *out_nfaces;
    int32_t *out_iters;
    int32_t *out_status;
    double *out_faces;
    int *counter;
    const int *perm;   // processing order: pairs grouped by (typeA, typeB)
    const int *n_dev;  // non-null: the length of `perm` is read from the device (fallback pass)
};

// ---------------------------------------------------------------------------
// Processing order.  EPA pairs arrive in candidate order; with 32 warps per SM each inside a
// different arm of the ten-way support switch the instruction cache thrashes
// (profiles/r01_ncu_k_epa_v2_occupancy7.txt: 5.4 stall cycles per issue waiting for
// instructions).  A counting sort by (typeA, typeB) - the same idea as in gjk.cu - makes
// co-resident warps run the same support code; pairs without a full simplex go last.
#define EPA_NBINS (D3D_NUM_TYPES * D3D_NUM_TYPES + 1)
struct EpaOrder {
    int *hist;     // [128]
    int *cursor;   // [128]
    uint8_t *keys; // [P]
    int *perm;     // [P]
};

__global__ void k_epa_keys(d3d_colliders c, const int32_t *__restrict__ pairs, const int32_t *npoints,
                           int64_t n, EpaOrder w) {
    __shared__ int sh[EPA_NBINS];
    for (int i = threadIdx.x; i < EPA_NBINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        int2 pr = __ldg(reinterpret_cast<const int2 *>(pairs) + k);
        int key = __ldg(c.type + pr.x) * D3D_NUM_TYPES + __ldg(c.type + pr.y);
        if (npoints && __ldg(npoints + k) != 4) key = EPA_NBINS - 1;
        w.keys[k] = (uint8_t)key;
        atomicAdd(&sh[key], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < EPA_NBINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&w.hist[i], sh[i]);
}

// Exclusive scan of the bins; *wide_types = 1 when a pair has a collider that is neither an analytic
// primitive nor a hull (the thread kernel then runs its all-types instance, see k_epa_thread).
__global__ void k_epa_scan(EpaOrder w, int *wide_types) {
    if (threadIdx.x == 0) {
        int acc = 0, wide = 0;
        const int lean = D3D_PRIMITIVE_MASK | (1 << D3D_HULL);
        for (int i = 0; i < EPA_NBINS; ++i) {
            w.cursor[i] = acc;
            acc += w.hist[i];
            if (i < EPA_NBINS - 1 && w.hist[i] &&
                !(((lean >> (i / D3D_NUM_TYPES)) & 1) && ((lean >> (i % D3D_NUM_TYPES)) & 1)))
                wide = 1;
        }
        *wide_types = wide;
    }
}

__global__ void __launch_bounds__(256) k_epa_scatter(int64_t n, EpaOrder w) {
    __shared__ int hist[EPA_NBINS];
    const int64_t tile = 256 * 8;
    for (int64_t t0 = blockIdx.x * tile; t0 < n; t0 += (int64_t)gridDim.x * tile) {
        for (int i = threadIdx.x; i < EPA_NBINS; i += 256) hist[i] = 0;
        __syncthreads();
        int key[8], rank[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int64_t k = t0 + j * 256 + threadIdx.x;
            key[j] = k < n ? w.keys[k] : -1;
            if (key[j] >= 0) rank[j] = atomicAdd(&hist[key[j]], 1);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < EPA_NBINS; i += 256) {
            int h = hist[i];
            hist[i] = h ? atomicAdd(&w.cursor[i], h) : 0;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (key[j] >= 0) w.perm[hist[key[j]] + rank[j]] = (int)(t0 + j * 256 + threadIdx.x);
        __syncthreads();
    }
}

inline size_t epa_au(size_t x) { return (x + 255) / 256 * 256; }
// counters + histogram | keys | order | fallback list | thread-kernel state
inline size_t epa_ws_base(int64_t n) { return 2048 + epa_au((size_t)n) + 2 * epa_au((size_t)n * 4); }
inline size_t epa_ws_bytes(int64_t n);

#ifndef EPA_WARPS
#define EPA_WARPS 4
#endif

// MF > 0: max_faces known at compile time (the reference's default 64: constant offsets in
// every face access, 10 % of the kernel's instructions were address arithmetic); MF = 0: run time.
template <int MF>
struct WarpMem {
    double *faces;  // [12][max_faces]: v0 xyz, v1 xyz, v2 xyz, n xyz
    int *perm;      // [max_faces]
    int mf_rt;
    D3D_DEV int mfv() const { return MF ? MF : mf_rt; }
    D3D_DEV v3 fget(int i, int which) const {
        const int mf = mfv();
        return V3(faces[(3 * which) * mf + i], faces[(3 * which + 1) * mf + i], faces[(3 * which + 2) * mf + i]);
    }
    D3D_DEV void fset(int i, int which, v3 v) const {
        const int mf = mfv();
        faces[(3 * which) * mf + i] = v.x; faces[(3 * which + 1) * mf + i] = v.y; faces[(3 * which + 2) * mf + i] = v.z;
    }
};

// epa.py:99-102 compute_normal
D3D_DEV v3 face_normal(v3 v0, v3 v1, v3 v2) { return normalized(cross(v1 - v0, v2 - v0)); }

#ifndef EPA_BLOCKS_PER_SM
#define EPA_BLOCKS_PER_SM 6  // 80 registers, 24 warps per SM; r02 sweep (scripts/gpu_sessions/r02_run2.sh) C3 / C5 ms at 4,5,6,8: 7.3 7.7 7.2 7.8 / 493 491 481 494
#endif
// -DEPA_PROFILE (development builds only, scripts/build_variant_epa.sh): cycles per phase of the
// iteration and a few event counts, summed over all warps; read with d3d_debug_epa_profile.
#ifdef EPA_PROFILE
__device__ unsigned long long g_epa_prof[16];
#define EPA_PROF_DECL long long prof_t = clock64(); unsigned long long prof[16] = {0}
#define EPA_PROF(slot) { long long t_ = clock64(); prof[slot] += (unsigned long long)(t_ - prof_t); prof_t = t_; }
#define EPA_COUNT(slot, v) prof[slot] += (unsigned long long)(v)
#define EPA_PROF_FLUSH if (lane == 0) { for (int q_ = 0; q_ < 16; ++q_) if (prof[q_]) atomicAdd(&g_epa_prof[q_], prof[q_]); }
#else
#define EPA_PROF_DECL
#define EPA_PROF(slot)
#define EPA_COUNT(slot, v)
#define EPA_PROF_FLUSH
#endif

template <int MF>
__global__ void __launch_bounds__(EPA_WARPS * 32, EPA_BLOCKS_PER_SM)
k_epa(d3d_colliders c, const int32_t *__restrict__ pairs, int64_t n_pairs, EpaParams prm) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu, lt = (1u << lane) - 1;
    const int mf = MF ? MF : prm.max_faces, ml = prm.max_loose_edges;
    size_t per_warp = (size_t)12 * mf + (mf + 1) / 2 + 2 + 2 * D3D_COLLIDER_FIELDS;  // doubles
    WarpMem<MF> W;
    W.faces = smem + wid * per_warp;
    W.perm = reinterpret_cast<int *>(W.faces + 12 * mf);
    // The two collider records live in shared memory, one copy per warp.  Kept per lane they
    // went to local memory behind the out-of-line support switch (300 B x 32 lanes x 32 warps
    // per SM, more than the L1: profiles/r02_ncu_k_epa_v3_binned.txt shows 2e8 local
    // accesses, a 45 % L1 hit rate and 5.9 stall cycles per issue on the scoreboard).
    double *recA = W.faces + 12 * mf + (mf + 1) / 2 + 2, *recB = recA + D3D_COLLIDER_FIELDS;
    W.mf_rt = mf;
    const double eps = prm.epsilon;

    const int64_t n_items = prm.n_dev ? (int64_t)__ldg(prm.n_dev) : n_pairs;
    for (;;) {
        int k = 0;
        if (lane == 0) {
            k = atomicAdd(prm.counter, 1);
            k = k < n_items ? __ldg(prm.perm + k) : -1;
        }
        k = __shfl_sync(FULL, k, 0);
        if (k < 0) break;
        if (prm.npoints && __ldg(prm.npoints + k) != 4) {  // undefined in the reference (np.empty rows)
            if (lane == 0) {
                st3(prm.out_mtv + 3 * (int64_t)k, V3(0.0, 0.0, 0.0));
                prm.out_success[k] = 0;
                if (prm.out_nfaces) prm.out_nfaces[k] = 0;
                if (prm.out_iters) prm.out_iters[k] = 0;
                if (prm.out_status) prm.out_status[k] = D3D_EPA_BAD_SIMPLEX;
            }
            continue;
        }
        EPA_PROF_DECL;
        int2 pr = __ldg(reinterpret_cast<const int2 *>(pairs) + k);
        __syncwarp();  // the previous pair's records are no longer read
        ColliderSmem<1> A = stage_collider_warp(c, pr.x, recA, lane), B = stage_collider_warp(c, pr.y, recB, lane);

        // epa.py:83-97: zero-initialised polytope, faces ABC, ACD, ADB, BDC
#pragma unroll 1
        for (int i = lane; i < 12 * mf; i += 32) W.faces[i] = 0.0;
        __syncwarp();
        if (lane < 4) {
            const double *Y = prm.Y + 12 * (int64_t)k;
            const int ia[4] = {0, 0, 0, 1}, ib[4] = {1, 2, 3, 3}, ic[4] = {2, 3, 1, 2};
            v3 v0 = ld3(Y + 3 * ia[lane]), v1 = ld3(Y + 3 * ib[lane]), v2 = ld3(Y + 3 * ic[lane]);
            W.fset(lane, 0, v0); W.fset(lane, 1, v1); W.fset(lane, 2, v2);
            W.fset(lane, 3, face_normal(v0, v1, v2));
        }
        __syncwarp();
        int n_faces = 4, closest = 0, it = 0, status = D3D_INTERSECTION;
        bool done = false, success = false;
        v3 mtv = V3(0.0, 0.0, 0.0);
        EPA_PROF(0);

        for (it = 0; it < prm.max_iter; ++it) {
            // ---- A: closest face, first arg-min of sum(v0 * n) (epa.py:104-109)
            double best = 0.0;
            int bi = 0x7fffffff;
#pragma unroll 1
            for (int i = lane; i < n_faces; i += 32) {
                double d = dot_plain(W.fget(i, 0), W.fget(i, 3));
                if (bi == 0x7fffffff || d < best) { best = d; bi = i; }
            }
            closest = warp_first_extreme<false>(best, bi, bi != 0x7fffffff);
            double min_dist = __shfl_sync(FULL, best, closest & 31);  // the owner's local best IS slot `closest`
            EPA_PROF(1);
            // ---- B: support point of A - B in the face normal (epa.py:62-65)
            v3 sd = W.fget(closest, 3);
            v3 new_point = support_call<32, 1, D3D_ALL_TYPES_MASK>(A.type, A.nv, A.V, recA, c.graph, sd.x, sd.y, sd.z, lane) -
                           support_call<32, 1, D3D_ALL_TYPES_MASK>(B.type, B.nv, B.V, recB, c.graph, -sd.x, -sd.y, -sd.z, lane);
            __syncwarp();  // MeshGraph: lane 0 stored the vertex the climb ended on
            EPA_PROF(2);
            // ---- C: convergence (epa.py:67-70)
            double proj = dot_blas(new_point, sd);
            if (proj - min_dist < eps) {
                mtv = sd * proj;
                success = true; done = true; ++it;
                break;
            }
            // ---- D: faces that see the new point (epa.py:122-124, 157-165)
            unsigned vis_lo = 0, vis_hi = 0;
            {
                bool v0 = false, v1 = false;
                if (lane < n_faces) v0 = dot_blas(W.fget(lane, 3), new_point - W.fget(lane, 0)) > eps;
                if (lane + 32 < n_faces)
                    v1 = dot_blas(W.fget(lane + 32, 3), new_point - W.fget(lane + 32, 0)) > eps;
                vis_lo = __ballot_sync(FULL, v0);
                vis_hi = __ballot_sync(FULL, v1);
            }
            // replay of the reference's swap-with-last loop on slot indices:
            // perm[s] = original slot of the face that ends up in slot s
#pragma unroll 1
            for (int i = lane; i < n_faces; i += 32) W.perm[i] = i;
            __syncwarp();
            // The loose-edge list lives in registers, entry e in lane e (max_loose_edges <= 32):
            // no shared-memory traffic and no warp barriers inside the edge loop.
            int n_loose = 0;
            v3 la = V3(0.0, 0.0, 0.0), lb = la;
            int nf = n_faces;
            EPA_PROF(3);
            EPA_COUNT(8, 1); EPA_COUNT(9, n_faces); EPA_COUNT(10, __popc(vis_lo) + __popc(vis_hi));
            {
                // visibility by SLOT: perm is the identity here, and removing slot i moves the
                // face (and its bit) of slot nf-1 into slot i.  The scan jumps from one visible
                // slot to the next instead of stepping over every face.
                unsigned long long vis_slot = ((unsigned long long)vis_hi << 32) | vis_lo;
                int i = 0;
                for (;;) {  // uniform: every lane replays the same integer bookkeeping
                    unsigned long long rest = (vis_slot >> i) << i;
                    if (nf < 64) rest &= (1ull << nf) - 1ull;
                    if (rest == 0ull) break;
                    i = __ffsll((long long)rest) - 1;
                    int f = W.perm[i];
                    // epa.py:167-187: edges of the removed face against the loose-edge list
                    v3 fv0 = W.fget(f, 0), fv1 = W.fget(f, 1), fv2 = W.fget(f, 2);  // broadcast reads
                    // one edge against the list; false when the list is full (epa.py:193-198, the
                    // caller then drops the remaining edges of this face).  Three inlined copies:
                    // selecting the end points by a loop index cost 9 % of the kernel's instructions.
                    auto edge = [&](v3 e0, v3 e1) -> bool {
                        // np.linalg.norm(x) < eps without the square root (exactly equivalent)
                        v3 d0 = lb - e0, d1 = la - e1;
                        bool match = lane < n_loose && dot_blas(d0, d0) < prm.eps_sq_thr &&
                                     dot_blas(d1, d1) < prm.eps_sq_thr;
                        unsigned mm = __ballot_sync(FULL, match);
                        if (mm) {  // first matching edge wins; overwrite_edge_with_last_edge (epa.py:200-202)
                            int found = __ffs(mm) - 1, last = n_loose - 1;
                            v3 ta = V3(__shfl_sync(FULL, la.x, last), __shfl_sync(FULL, la.y, last),
                                       __shfl_sync(FULL, la.z, last));
                            v3 tb = V3(__shfl_sync(FULL, lb.x, last), __shfl_sync(FULL, lb.y, last),
                                       __shfl_sync(FULL, lb.z, last));
                            if (lane == found) { la = ta; lb = tb; }
                            --n_loose;
                        } else {  // add_edge_to_list (epa.py:193-198)
                            if (n_loose >= ml) return false;
                            if (lane == n_loose) { la = e0; lb = e1; }
                            ++n_loose;
                        }
                        return true;
                    };
                    if (edge(fv0, fv1) && edge(fv1, fv2)) edge(fv2, fv0);
                    // remove_face (epa.py:118-120): slot i takes the last face, re-test slot i.
                    // Every lane stores the same value; the barriers only order the uniform
                    // reads and writes of the other lanes (racecheck-clean).
                    int moved = W.perm[nf - 1];
                    __syncwarp();
                    W.perm[i] = moved;
                    __syncwarp();
                    unsigned long long last_bit = (vis_slot >> (nf - 1)) & 1ull;
                    vis_slot = (vis_slot & ~(1ull << i) & ~(1ull << (nf - 1))) | (i < nf - 1 ? (last_bit << i) : 0ull);
                    --nf;
                }
            }
            EPA_PROF(4);
            EPA_COUNT(11, n_faces - nf); EPA_COUNT(12, n_loose);
            // apply the permutation: slot s <- original slot perm[s] (reads before writes)
#pragma unroll 1
            for (int base = 0; base < n_faces; base += 32) {
                int s = base + lane;
                int src = s < n_faces ? W.perm[s] : s;
                v3 a0, a1, a2, a3;
                bool mv = s < n_faces && src != s;
                if (mv) { a0 = W.fget(src, 0); a1 = W.fget(src, 1); a2 = W.fget(src, 2); a3 = W.fget(src, 3); }
                __syncwarp();
                if (mv) { W.fset(s, 0, a0); W.fset(s, 1, a1); W.fset(s, 2, a2); W.fset(s, 3, a3); }
                __syncwarp();
            }
            n_faces = nf;
            EPA_PROF(5);
            // ---- E: one new face per loose edge (epa.py:126-146)
            bool overflow = false;
            if (n_loose > 0) {
                int e = lane;
                bool have = e < n_loose;
                v3 v0 = V3(0, 0, 0), v1 = v0, nrm = v0;
                bool valid = false;
                if (have) {
                    v0 = la; v1 = lb;
                    nrm = face_normal(v0, v1, new_point);
                    valid = !(dot_blas(nrm, nrm) < prm.half_sq_thr);
                }
                unsigned vm = __ballot_sync(FULL, valid);
                unsigned hm = __ballot_sync(FULL, have);
                int pos = n_faces + __popc(vm & lt);
                // assert self.n_faces < self.max_faces, evaluated before every edge
                overflow = __ballot_sync(FULL, have && pos >= mf) != 0;
                v3 w0 = v0, wn = nrm;
                if (overflow) valid = false;
                if (valid && dot_blas(v0, nrm) + 1e-6 < 0.0) { w0 = v1; wn = -nrm; }  // epa.py:139-146
                if (valid) { W.fset(pos, 0, w0); W.fset(pos, 1, v1); W.fset(pos, 2, new_point); W.fset(pos, 3, wn); }
                int n_new = __popc(vm);
                // a skipped degenerate face behind the last valid one leaves its data in the next slot
                int last_have = 31 - __clz(hm);
                bool trailing = !((vm >> last_have) & 1u);
                if (!overflow && trailing && lane == last_have && n_faces + n_new < mf) {
                    W.fset(n_faces + n_new, 0, v0); W.fset(n_faces + n_new, 1, v1);
                    W.fset(n_faces + n_new, 2, new_point); W.fset(n_faces + n_new, 3, nrm);
                }
                n_faces += n_new;
                __syncwarp();
            }
            EPA_PROF(6);
            if (overflow) { status = D3D_EPA_MAX_FACES; done = true; ++it; break; }
            __syncwarp();
        }
        if (!done) {  // epa.py:76-78
            v3 n = W.fget(closest, 3);
            mtv = n * dot_blas(W.fget(closest, 0), n);
        }
        if (lane == 0) {
            st3(prm.out_mtv + 3 * (int64_t)k, mtv);
            if (c.mesh_last) {  // mesh.py:85
                if (A.type == D3D_MESH) c.mesh_last[pr.x] = A.mesh_cur();
                if (B.type == D3D_MESH) c.mesh_last[pr.y] = B.mesh_cur();
            }
            prm.out_success[k] = success ? 1 : 0;
            if (prm.out_nfaces) prm.out_nfaces[k] = n_faces;
            if (prm.out_iters) prm.out_iters[k] = it;
            if (prm.out_status) prm.out_status[k] = status;
        }
        if (prm.out_faces) {
            double *o = prm.out_faces + (int64_t)k * mf * 12;
#pragma unroll 1
            for (int i = lane; i < mf; i += 32)
#pragma unroll 1
                for (int w = 0; w < 4; ++w) st3(o + 12 * i + 3 * w, W.fget(i, w));
        }
        EPA_PROF(7);
        EPA_PROF_FLUSH;
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------
// Thread-per-pair EPA (the default limits max_faces = 64, max_loose_edges = 32, max_iter <= 64).
//
// The warp kernel above spends most of its issue slots on work that one lane could do: the two
// support calls, the convergence test and the edge bookkeeping are replicated on 32 lanes
// (profiles/r02_epa_warp_kernel_phases.txt: ~2000 warp instructions per expanding iteration).
// Here a thread owns a pair and replays epa.py:9-202 sequentially, 32 pairs per warp instruction:
//   * polytope vertices get small integer ids (table of <= 4 + max_iter points); a face is three
//     ids, its normal and its cached distance np.sum(v0 * n) (the value find_face_closest_to_origin
//     recomputes every iteration from the same two vectors); a loose edge is two ids.
//   * the reference matches edges by COORDINATES with a tolerance (epa.py:189-191).  A new vertex
//     is compared with every vertex of the table: bit-equal -> it re-uses that id; closer than
//     epsilon but not equal -> the two ids are remembered as a "near pair" (up to four per
//     polytope) which the edge match treats as equal, exactly like the reference's distance test.
//     Without near pairs - nearly always - "ids equal" is the whole test.
//   * visibility is decided by dot(n, p) - dist with an error bound, the reference's expression
//     (which needs the face's first vertex) only inside the rounding band; the closest face of the
//     next iteration is collected on the way (see the comments in the loop).
//   * state lives in global memory, element e of thread t at [e * T + t]: the threads of a warp
//     walk their face lists in step, so the accesses coalesce (256-byte rows); the kernel is bound
//     by the DRAM traffic of that state (profiles/r02_ncu_k_epa_thread_v4.txt,
//     r02_epa_thread_kernel_phases.txt).  Collider records and the loose-edge list are in shared
//     memory.
//   * pairs the thread kernel does not finish (MeshGraph colliders, hulls above
//     EPAT_MAX_VERTICES, a fifth near pair, max_iter reached - the reference then reads a face
//     slot as it looks at that time) go to a list that the warp kernel processes afterwards from
//     scratch.
// One trip of the kernel's loop is one EPA iteration of every lane that owns a pair; the loop is
// warp-uniform and a warp takes its 32 pairs together (EPAT_REFILL_MIN), see below.
#define EPAT_THREADS 128
#ifndef EPAT_BLOCKS_PER_SM
#define EPAT_BLOCKS_PER_SM 4
#endif
#define EPAT_MAX_BLOCKS (148 * EPAT_BLOCKS_PER_SM)
#define EPAT_MAXV 68  // 4 + max_iter
#define EPAT_MF 64
#define EPAT_ML 32
#define EPAT_TM (D3D_ALL_TYPES_MASK & ~(1 << D3D_MESH))
#define EPAT_LEAN_TM (D3D_PRIMITIVE_MASK | (1 << D3D_HULL))
#ifndef EPAT_MAX_VERTICES
// Hulls with more vertices go to the warp kernel (cooperative vertex scan).  Measured on C3 (hulls
// of 64-256 vertices, 4 Mi pairs): warp kernel 243 ms; thread kernel with a serial scan per thread
// 266 ms; thread kernel whose warp scans the wide hulls of its 32 lanes one after the other 268 ms
// (64 dependent scans per warp and iteration) - both rejected.
#define EPAT_MAX_VERTICES 32
#endif
#ifndef EPAT_VB
#define EPAT_VB 4
#endif
#ifndef EPAT_QB
#define EPAT_QB 4
#endif
#ifndef EPAT_MIN_PAIRS
#define EPAT_MIN_PAIRS 20000
#endif
#define EPAT_STATE_BYTES ((size_t)(EPAT_MAXV * 24 + EPAT_MF * 24 + EPAT_MF * 8 + EPAT_MF * 4))
#define EPAT_SMEM_BYTES ((size_t)EPAT_THREADS * (2 * D3D_COLLIDER_FIELDS * 8 + EPAT_ML * 2))

// -DEPA_PROFILE: cycles per phase of the thread kernel's iteration loop (lane 0 of every warp;
// slots 0-6 = refill, tie scan, support, vertex id, visibility, removal, new faces; 8 = lanes
// owning a pair summed over the trips, 9 = trips); scripts/epa_thread_profile.py
#ifdef EPA_PROFILE
#define EPAT_PROF_DECL long long prof_t = clock64(); unsigned long long prof[16] = {0}
#define EPAT_PROF(slot) { long long t_ = clock64(); prof[slot] += (unsigned long long)(t_ - prof_t); prof_t = t_; }
#define EPAT_COUNT(slot, v) prof[slot] += (unsigned long long)(v)
#define EPAT_PROF_FLUSH if ((threadIdx.x & 31) == 0) { for (int q_ = 0; q_ < 16; ++q_) if (prof[q_]) atomicAdd(&g_epa_prof[q_], prof[q_]); }
#else
#define EPAT_PROF_DECL
#define EPAT_PROF(slot)
#define EPAT_COUNT(slot, v)
#define EPAT_PROF_FLUSH
#endif

struct EpaThreadState {
    double *vtx;     // [EPAT_MAXV * 3][T]
    double *fnrm;    // [EPAT_MF * 3][T]
    double *fdist;   // [EPAT_MF][T]
    uint32_t *fids;  // [EPAT_MF][T]  id0 | id1 << 8 | id2 << 16
    int64_t T;
    int *fb_count;
    int *fb_list;
    const int *wide_types;  // written by k_epa_scan
};

static __device__ __noinline__ v3 face_normal_call(v3 v0, v3 v1, v3 v2) { return face_normal(v0, v1, v2); }

// Two instances by the types the support switch compiles in (the kernel stalls on instruction
// fetch: C5 mix 37.6 -> 35.9 ms with the lean one): TM = primitives + hulls runs when the batch has
// no other type (flag written by k_epa_scan), TM = EPAT_TM otherwise; the other launch exits at once.
template <int TM>
__global__ void __launch_bounds__(EPAT_THREADS, EPAT_BLOCKS_PER_SM)
k_epa_thread(d3d_colliders c, const int32_t *__restrict__ pairs, int64_t n_pairs, EpaParams prm, EpaThreadState S) {
    extern __shared__ double smem[];
    if ((*S.wide_types != 0) != (TM == EPAT_TM)) return;
    const int tid = threadIdx.x;
    // element e of this thread's state sits at [e * T]; rows x threads < 2^31, so the index
    // arithmetic is 32-bit (one IMAD.WIDE per access)
    const unsigned T = (unsigned)S.T;
    double *recA = smem + tid, *recB = recA + D3D_COLLIDER_FIELDS * EPAT_THREADS;
    // loose edges: 16 bits each (la | lb << 8), four to a 64-bit word, word w of this thread at
    // [w * EPAT_THREADS] (conflict free); the search compares four entries per load
    unsigned long long *loose64 = reinterpret_cast<unsigned long long *>(smem + 2 * D3D_COLLIDER_FIELDS * EPAT_THREADS) + tid;
    uint16_t *loose16 = reinterpret_cast<uint16_t *>(loose64);
#define LOOSE(e) loose16[(((e) >> 2) * EPAT_THREADS) * 4 + ((e) & 3)]
    double *vtx = S.vtx + blockIdx.x * (int64_t)EPAT_THREADS + tid;
    double *fnrm = S.fnrm + blockIdx.x * (int64_t)EPAT_THREADS + tid;
    double *fdist = S.fdist + blockIdx.x * (int64_t)EPAT_THREADS + tid;
    uint32_t *fids = S.fids + blockIdx.x * (int64_t)EPAT_THREADS + tid;
    const double eps = prm.epsilon;
#define VT(j) V3(vtx[(unsigned)(3 * (j)) * T], vtx[(unsigned)(3 * (j) + 1) * T], vtx[(unsigned)(3 * (j) + 2) * T])
#define FN(i) V3(fnrm[(unsigned)(3 * (i)) * T], fnrm[(unsigned)(3 * (i) + 1) * T], fnrm[(unsigned)(3 * (i) + 2) * T])
#define SET_FACE(i, ids_, n_, d_)                                                        \
    {                                                                                    \
        fids[(unsigned)(i) * T] = (ids_);                                                          \
        fnrm[(unsigned)(3 * (i)) * T] = (n_).x; fnrm[(unsigned)(3 * (i) + 1) * T] = (n_).y; fnrm[(unsigned)(3 * (i) + 2) * T] = (n_).z; \
        fdist[(unsigned)(i) * T] = (d_);                                                           \
    }
    int k = -1, n_faces = 0, it = 0, nv = 0, closest = 0;
    double vmax = 0.0;  // largest |coordinate| in the vertex table
    double min_dist = 0.0;
    bool need_scan = true;
    ColliderSmem<EPAT_THREADS> A, B;
    A.base = recA; B.base = recB; A.gpool = B.gpool = c.graph;
    A.type = B.type = 0; A.nv = B.nv = 0; A.V = B.V = nullptr;
    const unsigned FULL = 0xffffffffu;
    bool exhausted = false;

    // Near pairs: two DIFFERENT table entries closer than epsilon (the reference's edge test would
    // call them equal).  Up to four pairs (a | b << 8) are remembered per polytope; with none -
    // nearly always - "ids equal" is the whole test.
    unsigned long long near_pairs = 0ull, near_ids = 0ull;  // near_ids: bit i = vertex i is in a near pair
    int n_near = 0;
    auto same_vertex = [&](uint32_t x, uint32_t y) -> bool {
        if (x == y) return true;
        const unsigned long long a = x | (y << 8), b = y | (x << 8);
        bool hit = false;
        for (int q = 0; q < n_near; ++q) {
            const unsigned long long e = (near_pairs >> (16 * q)) & 0xffffull;
            hit |= e == a || e == b;
        }
        return hit;
    };
    // id of point p in the vertex table: an existing bit-equal vertex or a new entry; -1 when the
    // near-pair list is full (-> warp kernel)
    auto vertex_id = [&](v3 p) -> int {
        int same_as = -1, n_close = 0;
        unsigned close_ids = 0u;  // up to four table entries within epsilon of p
        for (int j0 = 0; j0 < nv; j0 += EPAT_QB) {  // EPAT_QB table entries per trip, loads in flight together
            v3 q[EPAT_QB];
#pragma unroll
            for (int u = 0; u < EPAT_QB; ++u) q[u] = VT(min(j0 + u, EPAT_MAXV - 1));
#pragma unroll
            for (int u = 0; u < EPAT_QB; ++u) {
                v3 d = q[u] - p;
                if (j0 + u < nv && dot_blas(d, d) < prm.eps_sq_thr) {
                    bool same = __double_as_longlong(q[u].x) == __double_as_longlong(p.x) &&
                                __double_as_longlong(q[u].y) == __double_as_longlong(p.y) &&
                                __double_as_longlong(q[u].z) == __double_as_longlong(p.z);
                    if (same && same_as < 0) same_as = j0 + u;
                    if (!same) {
                        if (n_close < 4) close_ids |= (unsigned)(j0 + u) << (8 * n_close);
                        ++n_close;
                    }
                }
            }
        }
        if (same_as >= 0) return same_as;
        if (n_near + n_close > 4) return -1;
        for (int q = 0; q < n_close; ++q) {
            const unsigned other = (close_ids >> (8 * q)) & 0xffu;
            near_pairs |= (unsigned long long)(other | ((unsigned)nv << 8)) << (16 * n_near);
            near_ids |= (1ull << min(other, 63u)) | (1ull << min(nv, 63));  // ids 63.. share the last bit
            ++n_near;
        }
        vtx[(unsigned)(3 * nv) * T] = p.x; vtx[(unsigned)(3 * nv + 1) * T] = p.y; vtx[(unsigned)(3 * nv + 2) * T] = p.z;
        vmax = fmax(vmax, fmax(fabs(p.x), fmax(fabs(p.y), fabs(p.z))));
        return nv++;
    };
    auto fall_back = [&]() { S.fb_list[atomicAdd(S.fb_count, 1)] = k; k = -1; };
    auto finish = [&](v3 mtv, bool success, int status) {
        st3(prm.out_mtv + 3 * (int64_t)k, mtv);
        prm.out_success[k] = success ? 1 : 0;
        if (prm.out_nfaces) prm.out_nfaces[k] = n_faces;
        if (prm.out_iters) prm.out_iters[k] = it + 1;
        if (prm.out_status) prm.out_status[k] = status;
        k = -1;
    };

    // Every trip of this loop is one EPA iteration of every lane that owns a pair.  The loop is
    // warp-uniform (ballots at its head, no lane leaves early) so the lanes re-converge on each
    // phase; idle lanes are refilled once EPAT_REFILL_MIN of them wait (32: a warp takes 32 pairs
    // and runs them to the end together - lanes that started together have polytopes of similar
    // size, which keeps the face loops of a warp the same length; C5 mix, ms at 8 / 16 / 24 / 32:
    // 53.0 45.3 41.1 39.0).
#ifndef EPAT_REFILL_MIN
#define EPAT_REFILL_MIN 32
#endif
    EPAT_PROF_DECL;
    for (;;) {
        unsigned run_mask = __ballot_sync(FULL, k >= 0);
        if (32 - __popc(run_mask) >= EPAT_REFILL_MIN || run_mask == 0) {
            if (k < 0 && !exhausted) do {  // next pair (epa.py:83-97)
                int q = atomicAdd(prm.counter, 1);
                if (q >= n_pairs) { exhausted = true; break; }
                k = __ldg(prm.perm + q);
                n_faces = 0; it = -1; nv = 0; vmax = 0.0; near_pairs = 0ull; near_ids = 0ull; n_near = 0;
                if (prm.npoints && __ldg(prm.npoints + k) != 4) {  // undefined in the reference (np.empty rows)
                    finish(V3(0.0, 0.0, 0.0), false, D3D_EPA_BAD_SIMPLEX);
                    break;
                }
                it = 0;
                int2 pr = __ldg(reinterpret_cast<const int2 *>(pairs) + k);
                A = stage_collider<EPAT_THREADS>(c, pr.x, recA);
                B = stage_collider<EPAT_THREADS>(c, pr.y, recB);
                if (A.type == D3D_MESH || B.type == D3D_MESH || A.nv > EPAT_MAX_VERTICES || B.nv > EPAT_MAX_VERTICES) {
                    fall_back();
                    break;
                }
                const double *Y = prm.Y + 12 * (int64_t)k;
                int id[4];
                bool near = false;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    id[i] = vertex_id(ld3(Y + 3 * i));
                    near |= id[i] < 0;
                }
                if (near) { fall_back(); break; }
#pragma unroll 1
                for (int f = 0; f < 4; ++f) {  // faces ABC, ACD, ADB, BDC
                    const int a = id[(0x1000 >> (4 * f)) & 3], b = id[(0x3321 >> (4 * f)) & 3], cc = id[(0x2132 >> (4 * f)) & 3];
                    v3 v0 = VT(a);
                    v3 n = face_normal_call(v0, VT(b), VT(cc));
                    const double d = dot_plain(v0, n);
                    SET_FACE(f, (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)cc << 16), n, d);
                    if (f == 0 || d < min_dist) { min_dist = d; closest = f; }  // first arg-min, in slot order
                }
                n_faces = 4;
                need_scan = false;
            } while (0);
            run_mask = __ballot_sync(FULL, k >= 0);
            if (run_mask == 0) {
                if (__all_sync(FULL, exhausted)) break;
                continue;
            }
        }
        EPAT_PROF(0);
        EPAT_COUNT(8, __popc(run_mask)); EPAT_COUNT(9, 1);
        bool go = k >= 0;
        // ---- closest face, first arg-min (epa.py:104-109).  The minimum over the faces that
        // survive an iteration is collected while their visibility is tested, and the new faces
        // are compared as they are made; only an exact tie between two survivors (whose slot
        // order changes with the removals) needs the scan over all faces.
        if (go && need_scan) {
            min_dist = fdist[0];
            closest = 0;
#pragma unroll 4
            for (int i = 1; i < n_faces; ++i) {
                double d = fdist[(unsigned)i * T];
                if (d < min_dist) { min_dist = d; closest = i; }
            }
        }
        __syncwarp();
        EPAT_PROF(1);
        // ---- support point of A - B in the face normal (epa.py:62-65), convergence (epa.py:67-70)
        v3 p = V3(0.0, 0.0, 0.0);
        int pid = 0;
        if (go) {
            v3 sd = FN(closest);
            p = support_call<1, EPAT_THREADS, TM>(A.type, A.nv, A.V, recA, c.graph, sd.x, sd.y, sd.z, 0) -
                support_call<1, EPAT_THREADS, TM>(B.type, B.nv, B.V, recB, c.graph, -sd.x, -sd.y, -sd.z, 0);
            double proj = dot_blas(p, sd);
            if (proj - min_dist < eps) {
                finish(sd * proj, true, D3D_INTERSECTION);
                go = false;
            }
        }
        __syncwarp();
        EPAT_PROF(2);
        if (go) {
            pid = vertex_id(p);
            if (pid < 0) { fall_back(); go = false; }
        }
        __syncwarp();
        EPAT_PROF(3);
        // ---- faces that see the new point (epa.py:122-124).  dot(n, p) - dist differs from the
        // reference's dot(n, p - v0) by rounding only; the reference's expression, which needs the
        // face's first vertex, is evaluated inside that band.  With u = 2^-53, M = largest
        // |coordinate| in the vertex table (p included), |n| = 1 so sum |n_i| <= sqrt(3), and
        // R = n . (p - v0) in exact arithmetic:
        //   reference: d = fl(p - v0) (u per component), three-term fma chain (3 u)
        //              -> |ref - R| <= 4.1 u sqrt(3) 2M = 14.2 u M
        //   here:      |fl(n . p) - n . p| <= 3 u sqrt(3) M, the same for dist = fl(v0 . n),
        //              the subtraction u |a - dist| <= 3.5 u M          -> |sgn - R| <= 13.9 u M
        // so |sgn - ref| < 29 u M = 3.2e-15 M; the band is 1e-13 M (30 x that).
        unsigned long long vis = 0ull;
        double smin = D3D_MAX_FLOAT;  // minimum of dist over the survivors, its slot, exact tie seen
        int smin_slot = 0;
        bool tie = false;
        if (go) {
            const double margin = 1e-13 * vmax;
            unsigned long long amb = 0ull;
            // EPAT_VB faces per trip, their 4 x EPAT_VB loads in flight together: a dependent round
            // of state loads costs ~2000 cycles (profiles/r02_epa_thread_kernel_phases.txt)
            for (int i0 = 0; i0 < n_faces; i0 += EPAT_VB) {
                v3 n[EPAT_VB];
                double dd[EPAT_VB];
#pragma unroll
                for (int u = 0; u < EPAT_VB; ++u) {
                    const int i = min(i0 + u, EPAT_MF - 1);
                    n[u] = FN(i);
                    dd[u] = fdist[(unsigned)i * T];
                }
#pragma unroll
                for (int u = 0; u < EPAT_VB; ++u) {
                    const int i = i0 + u;
                    if (i < n_faces) {
                        double sgn = dot_blas(n[u], p) - dd[u];
                        if (sgn > eps + margin) vis |= 1ull << i;
                        else if (sgn >= eps - margin) amb |= 1ull << i;
                        else if (dd[u] < smin) { smin = dd[u]; smin_slot = i; tie = false; }
                        else if (dd[u] == smin) tie = true;
                    }
                }
            }
            while (amb) {  // rare: inside the rounding band, the reference's own expression decides
                const int i = __ffsll((long long)amb) - 1;
                amb &= amb - 1ull;
                const double d = fdist[(unsigned)i * T];
                if (dot_blas(FN(i), p - VT(fids[(unsigned)i * T] & 0xffu)) > eps) vis |= 1ull << i;
                else if (d < smin) { smin = d; smin_slot = i; tie = false; }
                else if (d == smin) tie = true;
            }
        }
        __syncwarp();
        EPAT_PROF(4);
        // ---- remove them, loose edges (epa.py:157-202): removing slot i moves the last face (and
        // its visibility bit) into slot i
        int n_loose = 0;
        if (go) {
            int i = 0;
            for (;;) {
                unsigned long long rest = (vis >> i) << i;
                if (n_faces < 64) rest &= (1ull << n_faces) - 1ull;
                if (rest == 0ull) break;
                i = __ffsll((long long)rest) - 1;
                const uint32_t ids = fids[(unsigned)i * T];
                uint32_t ring = ids | (ids << 24);  // id0 id1 id2 id0
#pragma unroll 1
                for (int j = 0; j < 3; ++j, ring >>= 8) {
                    // entry = la | lb << 8; it matches when lb == e0 and la == e1 (epa.py:189-191)
                    const uint32_t e0 = ring & 0xffu, e1 = (ring >> 8) & 0xffu;
                    const uint16_t want = (uint16_t)(e1 | (e0 << 8));
                    int e = 0;
                    if ((((near_ids >> min(e0, 63u)) | (near_ids >> min(e1, 63u))) & 1ull) == 0ull) {
                        // neither end of this edge has a near twin: equal ids is the whole test
                        const unsigned long long pat = want * 0x0001000100010001ull;
                        e = n_loose;
                        for (int w = 0; 4 * w < n_loose; ++w) {
                            const unsigned long long x = loose64[w * EPAT_THREADS] ^ pat;
                            // lowest flagged field = first 16-bit field of x that is zero
                            const unsigned long long z = (x - 0x0001000100010001ull) & ~x & 0x8000800080008000ull;
                            if (z) {
                                const int hit = 4 * w + ((__ffsll((long long)z) - 1) >> 4);
                                if (hit < n_loose) e = hit;  // else: a stale entry behind the end of the list
                                break;
                            }
                        }
                    } else {
                        for (; e < n_loose; ++e) {
                            const uint32_t le = LOOSE(e);
                            if (same_vertex(le >> 8, e0) && same_vertex(le & 0xffu, e1)) break;
                        }
                    }
                    if (e < n_loose) {  // overwrite_edge_with_last_edge (epa.py:200-202)
                        --n_loose;
                        LOOSE(e) = LOOSE(n_loose);
                    } else {            // add_edge_to_list (epa.py:193-198)
                        if (n_loose >= EPAT_ML) break;
                        LOOSE(n_loose) = (uint16_t)(e0 | (e1 << 8));
                        ++n_loose;
                    }
                }
                const int last = n_faces - 1;  // remove_face (epa.py:118-120)
                const unsigned long long last_bit = (vis >> last) & 1ull;
                if (i != last) {
                    v3 n = FN(last);
                    SET_FACE(i, fids[(unsigned)last * T], n, fdist[(unsigned)last * T]);
                    if (smin_slot == last) smin_slot = i;
                }
                vis = (vis & ~(1ull << i) & ~(1ull << last)) | (i < last ? (last_bit << i) : 0ull);
                n_faces = last;
            }
        }
        __syncwarp();
        EPAT_PROF(5);
        // ---- one new face per loose edge (epa.py:126-146)
        if (go) {
            bool overflow = false;
#pragma unroll 1
            for (int e = 0; e < n_loose; ++e) {
                if (n_faces >= EPAT_MF) { overflow = true; break; }  // assert self.n_faces < self.max_faces
                const uint32_t ed = LOOSE(e);
                uint32_t id0 = ed & 0xffu;
                const uint32_t id1 = ed >> 8;
                v3 v0 = VT(id0), v1 = VT(id1);
                v3 n = face_normal_call(v0, v1, p);
                if (dot_blas(n, n) < prm.half_sq_thr) continue;
                if (dot_blas(v0, n) + 1e-6 < 0.0) { id0 = id1; v0 = v1; n = -n; }  // epa.py:139-146
                const double d = dot_plain(v0, n);
                SET_FACE(n_faces, id0 | (id1 << 8) | ((uint32_t)pid << 16), n, d);
                if (d < smin) { smin = d; smin_slot = n_faces; }  // survivors sit in lower slots: they win ties
                ++n_faces;
            }
            min_dist = smin; closest = smin_slot; need_scan = tie;
            if (overflow) finish(V3(0.0, 0.0, 0.0), false, D3D_EPA_MAX_FACES);
            else if (++it >= prm.max_iter) fall_back();  // epa.py:76-78 reads the last closest slot as it is now
        }
        EPAT_PROF(6);
    }
    EPAT_PROF_FLUSH;
#undef VT
#undef FN
#undef SET_FACE
#undef LOOSE
}

inline int64_t epat_threads_for(int64_t n) {
    int64_t blocks = d3d_min64((n + EPAT_THREADS - 1) / EPAT_THREADS, EPAT_MAX_BLOCKS);
    return blocks * EPAT_THREADS;
}
inline size_t epa_ws_bytes(int64_t n) { return epa_ws_base(n) + epa_au((size_t)epat_threads_for(n) * EPAT_STATE_BYTES); }

// smallest double t with sqrt(t) >= x (sqrt is correctly rounded and monotone on host and device)
double sqrt_threshold(double x) {
    double t = x * x;
    while (t > 0.0 && sqrt(nextafter(t, 0.0)) >= x) t = nextafter(t, 0.0);
    while (sqrt(t) < x) t = nextafter(t, 1e308);
    return t;
}

}  // namespace

extern "C" {

#ifdef EPA_PROFILE
int d3d_debug_epa_profile(unsigned long long *out16, int reset) {
    D3D_CUDA_CHECK(cudaDeviceSynchronize());
    D3D_CUDA_CHECK(cudaMemcpyFromSymbol(out16, g_epa_prof, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        D3D_CUDA_CHECK(cudaMemcpyToSymbol(g_epa_prof, z, sizeof(z)));
    }
    return 0;
}
#endif

size_t d3d_epa_workspace_bytes(int64_t n_pairs) { return epa_ws_bytes(n_pairs); }

int d3d_epa(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs, const double *Y,
            const int32_t *npoints, int max_iter, int max_loose_edges, int max_faces, double epsilon, double *out_mtv,
            uint8_t *out_success, int32_t *out_nfaces, int32_t *out_iters, int32_t *out_status,
            double *out_faces, void *workspace, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pairs == 0) return 0;
    if (!c || !pairs || !Y || !out_mtv || !out_success || !workspace)
        return d3d_set_error("d3d_epa: null argument");
    if (n_pairs > 0x7fffffff) return d3d_set_error("d3d_epa: more than 2^31-1 pairs in one call");
    if (ws_bytes < epa_ws_bytes(n_pairs)) return d3d_set_error("d3d_epa: workspace too small");
    if (max_faces < 4 || max_faces > 64 || max_loose_edges < 1 || max_loose_edges > 32)
        return d3d_set_error("d3d_epa: max_faces must be in [4, 64] and max_loose_edges in [1, 32]");
    EpaParams prm;
    prm.max_iter = max_iter; prm.max_loose_edges = max_loose_edges; prm.max_faces = max_faces;
    prm.epsilon = epsilon; prm.eps_sq_thr = sqrt_threshold(epsilon); prm.half_sq_thr = sqrt_threshold(0.5);
    prm.Y = Y; prm.npoints = npoints; prm.out_mtv = out_mtv; prm.out_success = out_success;
    prm.out_nfaces = out_nfaces; prm.out_iters = out_iters; prm.out_status = out_status;
    prm.out_faces = out_faces; prm.counter = reinterpret_cast<int *>(workspace);
    D3D_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 2048, stream));
    EpaOrder ord;
    {
        char *p = reinterpret_cast<char *>(workspace);
        ord.hist = reinterpret_cast<int *>(p + 512);
        ord.cursor = reinterpret_cast<int *>(p + 1280);
        ord.keys = reinterpret_cast<uint8_t *>(p + 2048);
        ord.perm = reinterpret_cast<int *>(p + 2048 + epa_au((size_t)n_pairs));
    }
    prm.perm = ord.perm;
    prm.n_dev = nullptr;
    const int sms = d3d_sm_count();
    {
        int kb = (int)d3d_min64((n_pairs + 255) / 256, (int64_t)sms * 8);
        k_epa_keys<<<kb, 256, 0, stream>>>(*c, pairs, npoints, n_pairs, ord);
        k_epa_scan<<<1, 32, 0, stream>>>(ord, reinterpret_cast<int *>(workspace) + 6);
        k_epa_scatter<<<(int)d3d_min64((n_pairs + 2047) / 2048, (int64_t)sms * 8), 256, 0, stream>>>(n_pairs, ord);
    }
    // Thread-per-pair kernel first (reference default limits, no polytope output); what it hands
    // back is finished by the warp kernel.  A thread works through its pair alone, so the kernel
    // needs many pairs per SM to pay off (measured cross-over ~2e4 pairs on B200); smaller
    // batches go to the warp kernel directly.  D3D_EPA_KERNEL=warp / thread overrides the choice
    // (tests compare the two kernels).
    const char *mode = getenv("D3D_EPA_KERNEL");
    const bool warp_only = mode ? strcmp(mode, "warp") == 0 : n_pairs < EPAT_MIN_PAIRS;
    if (mode && strcmp(mode, "warp") != 0 && strcmp(mode, "thread") != 0)
        return d3d_set_error("d3d_epa: D3D_EPA_KERNEL must be 'warp' or 'thread'");
    if (!warp_only && max_faces == EPAT_MF && max_loose_edges == EPAT_ML && max_iter <= EPAT_MAXV - 4 && !out_faces) {
        char *p = reinterpret_cast<char *>(workspace);
        EpaThreadState S;
        S.T = epat_threads_for(n_pairs);  // <= 148 * 4 * 128 threads, 204 rows each: 32-bit element indices
        S.fb_count = reinterpret_cast<int *>(p + 8);
        S.fb_list = reinterpret_cast<int *>(p + 2048 + epa_au((size_t)n_pairs) + epa_au((size_t)n_pairs * 4));
        char *st = p + epa_ws_base(n_pairs);
        S.vtx = reinterpret_cast<double *>(st);
        S.fnrm = S.vtx + (size_t)EPAT_MAXV * 3 * S.T;
        S.fdist = S.fnrm + (size_t)EPAT_MF * 3 * S.T;
        S.fids = reinterpret_cast<uint32_t *>(S.fdist + (size_t)EPAT_MF * S.T);
        S.wide_types = reinterpret_cast<int *>(p) + 6;
        D3D_CUDA_CHECK(cudaFuncSetAttribute(k_epa_thread<EPAT_LEAN_TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EPAT_SMEM_BYTES));
        D3D_CUDA_CHECK(cudaFuncSetAttribute(k_epa_thread<EPAT_TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EPAT_SMEM_BYTES));
        k_epa_thread<EPAT_LEAN_TM><<<(int)(S.T / EPAT_THREADS), EPAT_THREADS, EPAT_SMEM_BYTES, stream>>>(*c, pairs, n_pairs, prm, S);
        k_epa_thread<EPAT_TM><<<(int)(S.T / EPAT_THREADS), EPAT_THREADS, EPAT_SMEM_BYTES, stream>>>(*c, pairs, n_pairs, prm, S);
        D3D_CUDA_CHECK(cudaGetLastError());
        prm.perm = S.fb_list;
        prm.n_dev = S.fb_count;
        prm.counter = reinterpret_cast<int *>(p + 16);
    }
    size_t per_warp = (size_t)12 * max_faces + (max_faces + 1) / 2 + 2 + 2 * D3D_COLLIDER_FIELDS;
    size_t smem = per_warp * EPA_WARPS * sizeof(double);
    int blocks = (int)d3d_min64((n_pairs + EPA_WARPS - 1) / EPA_WARPS, (int64_t)sms * (EPA_BLOCKS_PER_SM + 2));
    if (max_faces == 64) {
        D3D_CUDA_CHECK(cudaFuncSetAttribute(k_epa<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_epa<64><<<blocks, EPA_WARPS * 32, smem, stream>>>(*c, pairs, n_pairs, prm);
    } else {
        D3D_CUDA_CHECK(cudaFuncSetAttribute(k_epa<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_epa<0><<<blocks, EPA_WARPS * 32, smem, stream>>>(*c, pairs, n_pairs, prm);
    }
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
