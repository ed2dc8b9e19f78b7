// Batched Jolt-variant GJK (distance + closest points, and boolean intersection).
//
// Replaces the Python driver loops gjk_distance_jolt / gjk_intersection_jolt and
// their numba step functions (distance3d/gjk/_gjk_jolt.py:29-288) for MANY pairs
// per call.  Pipeline, all on the caller's stream, no host synchronisation:
//
//   k_pair_keys   key = (typeA, typeB) bin, or the "wide" bin when a collider has
//                 more than D3D_THREAD_HULL_MAX vertices; warp-aggregated histogram
//   k_bin_scan    exclusive scan of the 101 bins
//   k_bin_scatter counting-sort permutation (warp-aggregated cursors)
//   k_gjk<MODE,1>  persistent, one THREAD per pair over the sorted order, lanes
//                 refill from a warp-private chunk so a warp never idles on its
//                 slowest pair; analytic supports in registers, simplex in
//                 shared memory ([slot][component][thread], conflict free)
//   k_gjk<MODE,32> one WARP per pair for wide hulls: vertex max-dot by strided
//                 scan + __shfl_xor reduction (lowest index wins ties), the
//                 simplex solve is replicated on all lanes
#include <stdlib.h>

#include "d3d_common.cuh"
#include "d3d_simplex.cuh"
#include "d3d_support.cuh"

// hulls / meshes with more vertices than this go to the warp-per-pair kernel (measured on
// B200: thread-per-pair wins up to ~64 vertices, 103 vs 47 Mpairs/s at 17-32 vertices; the
// warp kernel wins 3x at 64-256 vertices)
// (runtime override for tuning: environment variable D3D_THREAD_HULL_MAX)
#define D3D_THREAD_HULL_MAX_DEFAULT 64
// Instances of the thread kernel by the types their support switch compiles in (the kernel is
// instruction-fetch bound, unused cases cost run time): analytic primitives, + ConvexHullVertices
// (the mixed-shape pipeline: C5 GJK stage 56.0 -> 51.1 ms against running its hull pairs on the
// all-types instance), everything.
#define GJK_CONVEX_MASK (D3D_PRIMITIVE_MASK | (1 << D3D_HULL))
#define GJK_INSTANCE_OF(TM) ((TM) == D3D_PRIMITIVE_MASK ? 0 : ((TM) == GJK_CONVEX_MASK ? 1 : 2))
#define D3D_NBINS (D3D_NUM_TYPES * D3D_NUM_TYPES + 1)
#define D3D_WIDE_BIN (D3D_NUM_TYPES * D3D_NUM_TYPES)

namespace {

struct GjkWorkspace {
    int *counters;  // sorted order = [primitive bins | primitive + hull bins | other thread bins | wide bin]
                    // [0] [7] [5] next index of thread instance 0 / 1 / 2     [1] next index, warp kernel
                    // [4] [6] [2] end of the range of instance 0 / 1 / 2 (2 = end of the thread ranges)
                    // [3] total
    int *hist;      // [128]
    int *cursor;    // [128]
    uint8_t *keys;  // [P]
    int *perm;      // [P]
    int64_t n_pairs;
    real *fin;      // [P][GJK_FIN_FIELDS] final state of the thread-kernel pairs (distance mode)
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Distance mode: the thread kernel parks the final simplex of every pair in the workspace
// (41 values per pair) and k_gjk_finish turns it into closest
// points afterwards; see k_gjk_finish for why.
#define GJK_FIN_FIELDS 41
inline size_t gjk_ws_bytes(int64_t n, bool distance) {
    return 4096 + align_up((size_t)n, 256) + align_up((size_t)n * 4, 256) +
           (distance ? (size_t)n * GJK_FIN_FIELDS * sizeof(double) : 0);
}

inline GjkWorkspace carve(void *ws, int64_t n) {
    GjkWorkspace w;
    char *p = reinterpret_cast<char *>(ws);
    w.counters = reinterpret_cast<int *>(p);
    w.hist = reinterpret_cast<int *>(p + 1024);
    w.cursor = reinterpret_cast<int *>(p + 2048);
    w.keys = reinterpret_cast<uint8_t *>(p + 4096);
    w.perm = reinterpret_cast<int *>(p + 4096 + align_up((size_t)n, 256));
    w.n_pairs = n;
    w.fin = reinterpret_cast<real *>(p + 4096 + align_up((size_t)n, 256) + align_up((size_t)n * 4, 256));
    return w;
}

__device__ __forceinline__ bool is_wide(const d3d_colliders &c, int i, int hull_max) {
    int t = __ldg(c.type + i);
    return (t == D3D_HULL || t == D3D_MESH) && __ldg(c.vert_len + i) > hull_max;
}

__global__ void k_pair_keys(d3d_colliders c, const int32_t *__restrict__ pairs, int64_t n,
                            GjkWorkspace w, int hull_max) {
    __shared__ int sh[D3D_NBINS];
    for (int i = threadIdx.x; i < D3D_NBINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        int2 pr = __ldg(reinterpret_cast<const int2 *>(pairs) + k);
        int key;
        if (is_wide(c, pr.x, hull_max) || is_wide(c, pr.y, hull_max)) key = D3D_WIDE_BIN;
        else key = __ldg(c.type + pr.x) * D3D_NUM_TYPES + __ldg(c.type + pr.y);
        w.keys[k] = (uint8_t)key;
        atomicAdd(&sh[key], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D3D_NBINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&w.hist[i], sh[i]);
}

__global__ void k_bin_scan(GjkWorkspace w, int64_t n, int split_min) {
    __shared__ int hist[D3D_NBINS];
    for (int i = threadIdx.x; i < D3D_NBINS; i += blockDim.x) hist[i] = w.hist[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        // Class of a bin = the smallest instance whose switch covers both types.  Each class
        // present starts as its own group (= one launch of its instance); a group below
        // split_min pairs is merged with the group above it (or, at the top, below it) and the
        // merged group runs on the larger instance: for small mixed batches the tail of a
        // second launch costs more than the leaner instance saves (1 Mi pairs with 30 %
        // hulls: 1.73e8 pairs/s split, 1.90e8 unsplit).
        int cls[D3D_WIDE_BIN];
        long long n_cls[3] = {0, 0, 0};
        for (int i = 0; i < D3D_WIDE_BIN; ++i) {
            const int ta = i / D3D_NUM_TYPES, tb = i % D3D_NUM_TYPES;
            const bool prim = ((D3D_PRIMITIVE_MASK >> ta) & 1) && ((D3D_PRIMITIVE_MASK >> tb) & 1);
            const bool convex = ((GJK_CONVEX_MASK >> ta) & 1) && ((GJK_CONVEX_MASK >> tb) & 1);
            cls[i] = prim ? 0 : (convex ? 1 : 2);
            n_cls[cls[i]] += hist[i];
        }
        int group_of[3] = {0, 1, 2};  // instance that runs class c
        for (int round = 0; round < 2; ++round) {
            long long size[3] = {0, 0, 0};
            for (int c = 0; c < 3; ++c) size[group_of[c]] += n_cls[c];
            int groups = (size[0] > 0) + (size[1] > 0) + (size[2] > 0);
            if (groups < 2) break;
            for (int g = 0; g < 3; ++g) {
                if (size[g] == 0 || size[g] >= split_min) continue;
                int up = -1, down = -1;
                for (int h = g + 1; h < 3; ++h) if (size[h] > 0) { up = h; break; }
                for (int h = g - 1; h >= 0; --h) if (size[h] > 0) { down = h; break; }
                // merging down moves the lower group UP into this instance (it covers more types)
                for (int c = 0; c < 3; ++c) {
                    if (up >= 0 && group_of[c] == g) group_of[c] = up;
                    else if (up < 0 && down >= 0 && group_of[c] == down) group_of[c] = g;
                }
                break;
            }
        }
        int acc = 0;
        for (int g = 0; g < 3; ++g) {
            for (int i = 0; i < D3D_WIDE_BIN; ++i) {
                if (group_of[cls[i]] != g) continue;
                w.cursor[i] = acc;
                acc += hist[i];
            }
            w.counters[g == 0 ? 4 : (g == 1 ? 6 : 2)] = acc;
        }
        w.cursor[D3D_WIDE_BIN] = acc;
        acc += hist[D3D_WIDE_BIN];
        w.counters[0] = 0;
        w.counters[1] = 0;
        w.counters[5] = 0;
        w.counters[7] = 0;
        w.counters[3] = acc;
    }
}

// Counting-sort scatter.  A block ranks a tile of keys in a shared-memory histogram and
// reserves one range per (tile, bin) with a single global atomic: the batch has a few dozen
// distinct keys, so per-warp reservations serialise on the same few addresses (4 Mi pairs:
// 0.49 ms with one atomic per warp and key, profiles/r01_launches_final.csv).
#define BIN_TILE_ITEMS 8
__global__ void __launch_bounds__(256) k_bin_scatter(int64_t n, GjkWorkspace w) {
    __shared__ int hist[D3D_NBINS];
    const int64_t tile = 256 * BIN_TILE_ITEMS;
    for (int64_t t0 = blockIdx.x * tile; t0 < n; t0 += (int64_t)gridDim.x * tile) {
        for (int i = threadIdx.x; i < D3D_NBINS; i += 256) hist[i] = 0;
        __syncthreads();
        int key[BIN_TILE_ITEMS], rank[BIN_TILE_ITEMS];
#pragma unroll
        for (int j = 0; j < BIN_TILE_ITEMS; ++j) {
            int64_t k = t0 + j * 256 + threadIdx.x;
            key[j] = k < n ? w.keys[k] : -1;
            if (key[j] >= 0) rank[j] = atomicAdd(&hist[key[j]], 1);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < D3D_NBINS; i += 256) {
            int h = hist[i];
            hist[i] = h ? atomicAdd(&w.cursor[i], h) : 0;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < BIN_TILE_ITEMS; ++j)
            if (key[j] >= 0) w.perm[hist[key[j]] + rank[j]] = (int)(t0 + j * 256 + threadIdx.x);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Per-pair state in shared memory.  Field f of the owning thread (thread kernel:
// STRIDE = block size) or of the owning warp (warp kernel: STRIDE = 1) lives at
// base[f * STRIDE]:
//   [ 0,16)  collider A record (d3d_support.cuh ColliderSmem)
//   [16,32)  collider B record
//   [32,44)  Y  simplex points of A - B   (slot s, component c at 32 + 3 s + c)
//   [44,56)  P  support points on A
//   [56,68)  Q  support points on B
#define GJK_FIELDS 68       // warp kernel: everything in shared memory
// GJK_PQ_GLOBAL = 1 (thread kernel, distance mode): P and Q are written once per iteration
// and read only at the very end, so they do not live in shared memory at all: the new support
// points go straight into the pair's parked record in the workspace (the record k_gjk_finish
// reads anyway; it stays in L2 while the pair is in flight).  44 instead of 68 values per
// thread.  0 = P / Q in shared memory.
#ifndef GJK_PQ_GLOBAL
#define GJK_PQ_GLOBAL 1
#endif
#define GJK_FIELDS_THREAD (GJK_PQ_GLOBAL ? 44 : 68)
#define GJK_OFF_B 16
#define GJK_OFF_Y 32
#define GJK_OFF_P 44
#define GJK_OFF_Q 56

// Simplex storage.  Points never move: slot i of the simplex (the reference's row i of Y / P / Q)
// lives in PHYSICAL slot (perm >> 2 i) & 3, and update_simplex_y / update_simplex_ypq
// (_gjk_jolt.py:643-664, "keep the rows selected by the mask, in order") only rewrites the
// 8-bit permutation instead of copying up to nine vectors through shared memory.
template <int STRIDE>
struct Simplex {
    real *base;
    real *pq;  // thread kernel with GJK_PQ_GLOBAL: P then Q of this pair in its parked record (stride 1)
    int perm;
    D3D_DEV int phys(int i) const { return (perm >> (2 * i)) & 3; }
    D3D_DEV real &at_phys(int off, int ps, int c) const {
        if (GJK_PQ_GLOBAL && STRIDE != 1 && off >= GJK_OFF_P) return pq[(off - GJK_OFF_P) + 3 * ps + c];
        return base[(off + 3 * ps + c) * STRIDE];
    }
    D3D_DEV real &at(int off, int s, int c) const { return at_phys(off, phys(s), c); }
    D3D_DEV v3 get(int off, int s) const {
        const int ps = phys(s);
        return V3(at_phys(off, ps, 0), at_phys(off, ps, 1), at_phys(off, ps, 2));
    }
    D3D_DEV void set(int off, int s, v3 v) const {
        const int ps = phys(s);
        at_phys(off, ps, 0) = v.x; at_phys(off, ps, 1) = v.y; at_phys(off, ps, 2) = v.z;
    }
    // physical slot for a new point behind the n live ones; the permutation entry n is set to it
    D3D_DEV void claim(int n) {
        int used = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (i < n) used |= 1 << phys(i);
        const int free_slot = __ffs(~used) - 1;
        perm = (perm & ~(3 << (2 * n))) | (free_slot << (2 * n));
    }
    // keep the slots selected by `mask` (bit i = slot i) in order; returns the new count
    D3D_DEV int compact(int n, int mask) {
        int np = 0, nn = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < n && (mask & (1 << i))) { np |= phys(i) << (2 * nn); ++nn; }
        perm = np;
        return nn;
    }
};

struct GjkParams {
    real tolerance_sq;
    real max_distance_squared;
    real sanity_check;
    double *out_dist;
    double *out_a;
    double *out_b;
    double *out_Y;
    int32_t *out_npoints;
    int32_t *out_iters;
    int32_t *out_status;
    uint8_t *out_hit;
    const int32_t *graph;  // d3d_colliders::graph (MeshGraph adjacency pool)
    int32_t *mesh_last;    // d3d_colliders::mesh_last or nullptr
};

template <int STRIDE>
struct PairState {
    ColliderSmem<STRIDE> A, B;
    v3 sd;
    real v_len_sq, prev_v_len_sq;
    int n_points, iters, k, state;
};

template <int STRIDE>
D3D_DEV void stage_pair(PairState<STRIDE> &s, const d3d_colliders &c, int2 pr, real *base) {
    s.A = stage_collider<STRIDE>(c, pr.x, base);
    s.B = stage_collider<STRIDE>(c, pr.y, base + GJK_OFF_B * STRIDE);
}
// warp-shared record: the lanes share the stores (racecheck-clean)
template <>
D3D_DEV void stage_pair<1>(PairState<1> &s, const d3d_colliders &c, int2 pr, real *base) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    s.A = stage_collider_warp(c, pr.x, base, lane);
    s.B = stage_collider_warp(c, pr.y, base + GJK_OFF_B, lane);
}

template <int STRIDE>
D3D_DEV void init_pair(PairState<STRIDE> &s, Simplex<STRIDE> &S, const d3d_colliders &c, const int32_t *pairs,
                       int k, real *base) {
    S.perm = 0;
    int2 pr = __ldg(reinterpret_cast<const int2 *>(pairs) + k);
    stage_pair(s, c, pr, base);
    s.sd = V3(R(1.0), R(0.0), R(0.0));
    s.v_len_sq = R(1.0);  // np.dot(sd, sd), _gjk_jolt.py:197
    s.prev_v_len_sq = D3D_MAX_FLOAT;
    s.n_points = 0;
    s.iters = 0;
    s.k = k;
    s.state = D3D_UNKNOWN;
}

// max(|Y_i|^2) over the slots selected by mask (_gjk_jolt.py:634-640)
D3D_DEV real max_y_len_sq(v3 y0, v3 y1, v3 y2, v3 y3, int mask) {
    real m = (mask & 1) ? dot_blas(y0, y0) : -R(1.0);
    if (mask & 2) m = fmax(m, dot_blas(y1, y1));
    if (mask & 4) m = fmax(m, dot_blas(y2, y2));
    if (mask & 8) m = fmax(m, dot_blas(y3, y3));
    return m;
}

// ---------------------------------------------------------------------------
// One GJK iteration = gjk_pre (supports, early exits, add the point) -> closest point of the
// simplex to the origin -> gjk_post (bookkeeping of _distance_loop / _intersection_loop).

// First half of _distance_loop (MODE 0, _gjk_jolt.py:228-242) / _intersection_loop (MODE 1,
// :86-97).  Returns true when the simplex solve is needed.
template <int MODE, int G, int STRIDE, int TM>
D3D_DEV bool gjk_pre(PairState<STRIDE> &s, Simplex<STRIDE> &S, const GjkParams &prm, int lane) {
    if (s.iters >= D3D_GJK_ITER_CAP) { s.state = D3D_ITER_CAP; return false; }
    ++s.iters;
    v3 p = support_call<G, STRIDE, TM>(s.A.type, s.A.nv, s.A.V, s.A.base, prm.graph, s.sd.x, s.sd.y, s.sd.z, lane);
    v3 q = support_call<G, STRIDE, TM>(s.B.type, s.B.nv, s.B.V, s.B.base, prm.graph, -s.sd.x, -s.sd.y, -s.sd.z, lane);
    v3 w = p - q;
    real dot = dot_blas(s.sd, w);
    if (MODE == 0) {
        if (dot < R(0.0) && dot * dot > s.v_len_sq * prm.max_distance_squared) {
            s.state = D3D_CLIPPED;
            return false;
        }
    } else {
        if (dot < -D3D_EPS) { s.state = D3D_NO_INTERSECTION; return false; }
    }
    if (STRIDE == 1) __syncwarp();
    S.claim(s.n_points);
    S.set(GJK_OFF_Y, s.n_points, w);
    if (MODE == 0) { S.set(GJK_OFF_P, s.n_points, p); S.set(GJK_OFF_Q, s.n_points, q); }
    if (STRIDE == 1) __syncwarp();
    ++s.n_points;
    return true;
}

// Second half (_gjk_jolt.py:244-288 / :100-135) given the solver's answer.
template <int MODE, int STRIDE>
D3D_DEV void gjk_post(PairState<STRIDE> &s, Simplex<STRIDE> &S, const GjkParams &prm, bool ok,
                      v3 v_new, real v_len_sq_new, int simplex) {
    v3 y0 = S.get(GJK_OFF_Y, 0), y1 = S.get(GJK_OFF_Y, 1), y2 = S.get(GJK_OFF_Y, 2),
       y3 = S.get(GJK_OFF_Y, 3);
    if (ok) {
        s.sd = v_new;
        s.v_len_sq = v_len_sq_new;
    } else {
        if (MODE == 1) { s.state = D3D_NO_INTERSECTION; return; }
        --s.n_points;  // undo add, keep all old points (_gjk_jolt.py:248-252)
        simplex = (1 << s.n_points) - 1;
    }
    if (simplex == 0xf) {
        if (MODE == 0) s.v_len_sq = R(0.0);
        s.state = D3D_INTERSECTION;
        return;
    }
    if (MODE == 1) {
        if (s.v_len_sq <= prm.tolerance_sq) { s.state = D3D_INTERSECTION; return; }
        if (s.v_len_sq <= D3D_EPS * max_y_len_sq(y0, y1, y2, y3, (1 << s.n_points) - 1)) {
            s.state = D3D_INTERSECTION;
            return;
        }
    }
    if (MODE == 0) {
        // update_simplex_ypq (_gjk_jolt.py:654-664): the rows stay where they are, the map changes
        s.n_points = S.compact(s.n_points, simplex);
        if (s.v_len_sq <= prm.tolerance_sq) { s.v_len_sq = R(0.0); s.state = D3D_INTERSECTION; return; }
        if (s.v_len_sq <= D3D_EPS * max_y_len_sq(y0, y1, y2, y3, simplex)) {
            s.v_len_sq = R(0.0);
            s.state = D3D_INTERSECTION;
            return;
        }
    }
    s.sd = s.sd * -R(1.0);
    if (!(s.prev_v_len_sq >= s.v_len_sq)) { s.state = D3D_MONOTONICITY; return; }
    if (s.prev_v_len_sq - s.v_len_sq <= D3D_EPS * s.prev_v_len_sq) {
        s.state = D3D_NO_INTERSECTION;
        return;
    }
    s.prev_v_len_sq = s.v_len_sq;
    if (MODE == 1) {
        // update_simplex_y (_gjk_jolt.py:643-651)
        s.n_points = S.compact(s.n_points, simplex);
    }
}

// Whole iteration with the per-lane solver (warp-per-pair kernel: all lanes are in step).
template <int MODE, int G, int STRIDE>
D3D_DEV void gjk_step(PairState<STRIDE> &s, Simplex<STRIDE> &S, const GjkParams &prm, int lane) {
    if (!gjk_pre<MODE, G, STRIDE, D3D_ALL_TYPES_MASK>(s, S, prm, lane)) return;
    v3 v_new;
    real v_len_sq_new;
    int simplex;
    bool ok = closest_point_to_origin(S.get(GJK_OFF_Y, 0), S.get(GJK_OFF_Y, 1), S.get(GJK_OFF_Y, 2),
                                      S.get(GJK_OFF_Y, 3), s.n_points, s.prev_v_len_sq, v_new,
                                      v_len_sq_new, simplex);
    gjk_post<MODE, STRIDE>(s, S, prm, ok, v_new, v_len_sq_new, simplex);
}

// ---------------------------------------------------------------------------
// Warp-collective simplex solve for the thread-per-pair kernel.
//
// A lane with a 3-point simplex has one candidate triangle, a lane with a tetrahedron
// one to four (the faces the origin lies outside of), lanes with 1-2 points none; walking
// them lane by lane leaves most of the warp idle (5.6 active threads per instruction in
// closest_triangle, profiles/r01_ncu_k_gjk_thread_v3_outlined_div.txt).  Here the
// (owner lane, face) work items of the whole warp are listed in shared memory and the warp
// evaluates them 32 at a time: any lane can take any item because the simplex points of
// every pair live in shared memory.  Owners then fold the results of their own items in
// ascending face order with the reference's strict '<' rule, so the outcome is bit-identical
// to closest_point_tetrahedron (_gjk_jolt.py:573-631).
struct TriResult {
    real qx, qy, qz, dist_sq;
};
struct WarpScratch {
    TriResult *res;       // [32] results of the current round
    int *set;             // [32] feature sets of the current round (already in tetra numbering)
    unsigned short *desc;  // [128] item -> (owner's slot permutation << 8 | owner lane << 2 | face)
};
#define GJK_SCRATCH_BYTES (32 * sizeof(TriResult) + 32 * sizeof(int) + 256)

template <int STRIDE>
D3D_DEV bool closest_point_to_origin_warp(bool solve, const Simplex<STRIDE> &S, int n_points,
                                          real prev_v_len_sq, const WarpScratch &W, int lane,
                                          v3 &v_out, real &v_len_sq_out, int &set_out) {
    const unsigned FULL = 0xffffffffu;
    v3 v = V3(R(0.0), R(0.0), R(0.0));
    int set = 1;
    int faces = 0;  // candidate faces of this lane (bit f)
    if (solve) {
        if (n_points == 1) {
            v = S.get(GJK_OFF_Y, 0);
        } else if (n_points == 2) {
            v = closest_line(S.get(GJK_OFF_Y, 0), S.get(GJK_OFF_Y, 1), set);
        } else if (n_points == 3) {
            faces = 1;
        } else {
            faces = origin_outside_planes(S.get(GJK_OFF_Y, 0), S.get(GJK_OFF_Y, 1), S.get(GJK_OFF_Y, 2),
                                          S.get(GJK_OFF_Y, 3));
            set = 0xf;  // inside all planes unless a face says otherwise
        }
    }
    // exclusive scan of the face counts over the warp
    int cnt = __popc(faces);
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int y = __shfl_up_sync(FULL, incl, off);
        if (lane >= off) incl += y;
    }
    int first = incl - cnt;
    int total = __shfl_sync(FULL, incl, 31);
    if (total == 0) {
        // nothing to share: finish locally
    } else {
        int j = first;
        for (int todo = faces; todo; todo &= todo - 1)
            W.desc[j++] = (unsigned short)(((S.perm & 0xff) << 8) | (lane << 2) | (__ffs(todo) - 1));
        __syncwarp();
        real best_dist_sq = D3D_MAX_FLOAT;
#pragma unroll 1
        for (int base = 0; base < total; base += 32) {
            int item = base + lane;
            if (item < total) {
                int d = W.desc[item];
                int o = (d >> 2) & 31, f = d & 3;
                Simplex<STRIDE> So;
                So.base = S.base + (o - lane);
                So.pq = nullptr;
                So.perm = d >> 8;
                // faces: abc, acd, adb, bdc (slot indices packed two bits each)
                const int ia = (f == 3) ? 1 : 0;
                const int ib = (f == 0) ? 1 : ((f == 1) ? 2 : 3);
                const int ic = (f == 0) ? 2 : ((f == 1) ? 3 : ((f == 2) ? 1 : 2));
                int sset;
                v3 q = closest_triangle(So.get(GJK_OFF_Y, ia), So.get(GJK_OFF_Y, ib), So.get(GJK_OFF_Y, ic), sset);
                int mapped;
                if (f == 0) mapped = sset;
                else if (f == 1) mapped = (sset & 1) + ((sset & 6) << 1);
                else if (f == 2) mapped = (sset & 1) + ((sset & 2) << 2) + ((sset & 4) >> 1);
                else mapped = ((sset & 1) << 1) + ((sset & 2) << 2) + (sset & 4);
                TriResult r;
                r.qx = q.x; r.qy = q.y; r.qz = q.z; r.dist_sq = dot_blas(q, q);
                W.res[lane] = r;
                W.set[lane] = mapped;
            }
            __syncwarp();
            // owners fold their items of this round, ascending face order
            int lo = max(first, base), hi = min(first + cnt, base + 32);
#pragma unroll 1
            for (int it = lo; it < hi; ++it) {
                TriResult r = W.res[it - base];
                int f = W.desc[it] & 3;
                if (f == 0 || r.dist_sq < best_dist_sq) {
                    best_dist_sq = r.dist_sq;
                    v = V3(r.qx, r.qy, r.qz);
                    set = W.set[it - base];
                }
            }
            __syncwarp();
        }
    }
    if (!solve) return false;
    real v_len_sq = dot_blas(v, v);
    if (v_len_sq < prev_v_len_sq) {
        v_out = v; v_len_sq_out = v_len_sq; set_out = set;
        return true;
    }
    return false;
}

// MeshGraph vertex caching across calls (mesh.py:85): report the vertex each mesh ended on.
template <int STRIDE>
static __device__ __noinline__ void write_mesh_last(const PairState<STRIDE> &s, const int32_t *pairs,
                                                    int32_t *mesh_last) {
    int2 pr = __ldg(reinterpret_cast<const int2 *>(pairs) + s.k);
    if (s.A.type == D3D_MESH) mesh_last[pr.x] = s.A.mesh_cur();
    if (s.B.type == D3D_MESH) mesh_last[pr.y] = s.B.mesh_cur();
}

// Closest points, sanity check and output (_gjk_jolt.py:209-221, 667-687).  PS: PairState or
// FinState, SX: Simplex or SimplexFin.
template <int MODE, class PS, class SX>
D3D_DEV void gjk_finish(const PS &s, const SX &S, const GjkParams &prm, bool writer) {
    int64_t k = s.k;
    int state = s.state;
    if (MODE == 1) {
        if (writer) {
            prm.out_hit[k] = (state == D3D_INTERSECTION) ? 1 : 0;
            if (prm.out_iters) prm.out_iters[k] = s.iters;
            if (prm.out_status) prm.out_status[k] = state;
        }
        return;
    }
    v3 a = V3(R(0.0), R(0.0), R(0.0)), b = a;
    real dist = D3D_MAX_FLOAT;
    if (state == D3D_NO_INTERSECTION || state == D3D_INTERSECTION) {
        int n = s.n_points;
        // barycentric weights of the closest point (zero weight for unused slots)
        real u = R(1.0), v = R(0.0), w = R(0.0), x = R(0.0);
        v3 y0 = S.get(GJK_OFF_Y, 0), y1 = S.get(GJK_OFF_Y, 1), y2 = S.get(GJK_OFF_Y, 2),
           y3 = S.get(GJK_OFF_Y, 3);
        if (n == 2) bary_line(y0, y1, u, v);
        else if (n == 3) bary_plane(y0, y1, y2, u, v, w);
        else if (n == 4) bary_tetra(y0, y1, y2, y3, u, v, w, x);
        // u*P0 + v*P1 + w*P2 + x*P3, left to right, only over the valid slots
        a = S.get(GJK_OFF_P, 0);
        b = S.get(GJK_OFF_Q, 0);
        if (n >= 2) {
            a = a * u + S.get(GJK_OFF_P, 1) * v;
            b = b * u + S.get(GJK_OFF_Q, 1) * v;
        }
        if (n >= 3) { a = a + S.get(GJK_OFF_P, 2) * w; b = b + S.get(GJK_OFF_Q, 2) * w; }
        if (n >= 4) { a = a + S.get(GJK_OFF_P, 3) * x; b = b + S.get(GJK_OFF_Q, 3) * x; }
        real check_value = fabs(dot_blas(s.sd, s.sd) - s.v_len_sq);
        if (!(check_value < prm.sanity_check)) state = D3D_SANITY_FAILED;
        dist = sqrt(s.v_len_sq);
        if (dist < D3D_EPS) { a = (a + b) * R(0.5); b = a; }
    }
    if (!writer) return;
    prm.out_dist[k] = dist;
    if (prm.out_a) st3(prm.out_a + 3 * k, a);
    if (prm.out_b) st3(prm.out_b + 3 * k, b);
    if (prm.out_Y) {
#pragma unroll 1
        for (int i = 0; i < 4; ++i)
            st3(prm.out_Y + 12 * k + 3 * i, i < s.n_points ? S.get(GJK_OFF_Y, i) : V3(0.0, 0.0, 0.0));
    }
    if (prm.out_npoints) prm.out_npoints[k] = s.n_points;
    if (prm.out_iters) prm.out_iters[k] = s.iters;
    if (prm.out_status) prm.out_status[k] = state;
}

#ifndef GJK_THREADS
#define GJK_THREADS 128
#endif
#ifndef GJK_BLOCKS_PER_SM
#ifdef D3D_F32
#define GJK_BLOCKS_PER_SM 5  // fp32 state is half the size: 40 KB per CTA
#else
#define GJK_BLOCKS_PER_SM 3
#endif
#endif
// the primitive-only instance fits four CTAs per SM once P / Q are out of shared memory
// (44 values per thread = 45 KB + scratch per CTA, 120 registers)
#ifndef GJK_BLOCKS_PRIM
#ifdef D3D_F32
#define GJK_BLOCKS_PRIM GJK_BLOCKS_PER_SM
#else
#define GJK_BLOCKS_PRIM (GJK_PQ_GLOBAL ? 4 : 3)
#endif
#endif
#ifndef GJK_BLOCKS_CONVEX
#define GJK_BLOCKS_CONVEX GJK_BLOCKS_PRIM  // primitives + hulls instance: 126 registers, fits a 4th CTA as well
#endif
#define GJK_BLOCKS_FOR(TM) ((TM) == D3D_PRIMITIVE_MASK ? GJK_BLOCKS_PRIM : ((TM) == GJK_CONVEX_MASK ? GJK_BLOCKS_CONVEX : GJK_BLOCKS_PER_SM))
#ifndef GJK_REFILL_MIN
#define GJK_REFILL_MIN 8
#endif
#ifndef GJK_CHUNK
#define GJK_CHUNK 0  // pairs per warp-private chunk of the sorted order; 0 = one global cursor.
// Measured (C1 mix, Mpairs/s at 1 Mi / 4 Mi pairs): cursor 277 / 335, chunks of 128: 284 / 332,
// 256: 269 / 334, 512: 209 / 312 - chunks keep a warp inside one type bin but unbalance small
// batches; the global cursor is the default.
#endif
// refill when at least this many lanes of the warp are idle

// ---------------------------------------------------------------------------
// Deferred finish (thread kernel, distance mode).  gjk_finish is ~480 instructions, run by
// the few lanes of a warp that have just finished a pair; inside the persistent kernel it
// is 28 % of the loop's instruction footprint, and the kernel is instruction-fetch bound
// (B200, 1 Mi mixed pairs: 2.17e8 pairs/s with the finish inline, 2.53e8 with the finish
// stubbed out).  So the kernel only parks the final state and k_gjk_finish, one convergent
// thread per pair, produces the outputs.
struct FinState {
    v3 sd;
    real v_len_sq;
    int n_points, iters, k, state;
};
struct SimplexFin {  // Y, P, Q of one parked pair (record of GJK_FIN_FIELDS values)
    const real *base;
    int perm;  // Y is parked in simplex order; P / Q sit in their physical slots (GJK_PQ_GLOBAL)
    D3D_DEV v3 get(int off, int s) const {
        const int ps = (GJK_PQ_GLOBAL && off >= GJK_OFF_P) ? ((perm >> (2 * s)) & 3) : s;
        const real *p = base + (off - GJK_OFF_Y + 3 * ps);
        return V3(p[0], p[1], p[2]);
    }
};
#ifdef D3D_F32
D3D_DEV real fin_pack(int v) { return __int_as_float(v); }
D3D_DEV int fin_unpack(real v) { return __float_as_int(v); }
#else
D3D_DEV real fin_pack(int v) { return __longlong_as_double((long long)v); }
D3D_DEV int fin_unpack(real v) { return (int)__double_as_longlong(v); }
#endif

template <int MODE>
D3D_DEV void gjk_finish_or_park(const PairState<GJK_THREADS> &s, const Simplex<GJK_THREADS> &S,
                                const GjkParams &prm, const GjkWorkspace &w) {
    if (MODE == 1) {
        gjk_finish<1>(s, S, prm, true);
        return;
    }
    // one contiguous record per pair, indexed by the pair's number k: the stores of a lane
    // fill whole sectors and k_gjk_finish reads and writes in pair order
    real *o = w.fin + (int64_t)s.k * GJK_FIN_FIELDS;
#pragma unroll 1
    for (int i = 0; i < s.n_points; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            o[3 * i + j] = S.at(GJK_OFF_Y, i, j);
            if (!GJK_PQ_GLOBAL) {
                o[12 + 3 * i + j] = S.at(GJK_OFF_P, i, j);
                o[24 + 3 * i + j] = S.at(GJK_OFF_Q, i, j);
            }
        }
    }
    o[36] = s.sd.x;
    o[37] = s.sd.y;
    o[38] = s.sd.z;
    o[39] = s.v_len_sq;
    // n_points (3 bits) | state (4 bits) | slot permutation (8 bits) | iterations
    o[40] = fin_pack(s.n_points | (s.state << 3) | ((GJK_PQ_GLOBAL ? (S.perm & 0xff) : 0xe4) << 7) | (s.iters << 15));
}

// Pairs of the warp kernel (wide hulls) are finished by that kernel: their key is the wide bin.
__global__ void __launch_bounds__(256) k_gjk_finish(GjkWorkspace w, GjkParams prm) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < w.n_pairs;
         k += (int64_t)gridDim.x * blockDim.x) {
        if (w.keys[k] == D3D_WIDE_BIN) continue;
        SimplexFin S;
        S.base = w.fin + k * GJK_FIN_FIELDS;
        FinState s;
        s.sd = V3(S.base[36], S.base[37], S.base[38]);
        s.v_len_sq = S.base[39];
        int packed = fin_unpack(S.base[40]);
        s.n_points = packed & 7;
        s.state = (packed >> 3) & 15;
        S.perm = (packed >> 7) & 0xff;
        s.iters = packed >> 15;
        s.k = (int)k;
        gjk_finish<0>(s, S, prm, true);
    }
}

// One thread per pair, persistent, lanes refill from a warp-private chunk.  Three instances per
// mode: TM = D3D_PRIMITIVE_MASK walks the sorted range of primitive-only bins with a support
// switch compiled for those five types (the kernel is instruction-fetch bound: +5 % on the C1
// mix), TM = GJK_CONVEX_MASK the bins of primitives and hulls, TM = D3D_ALL_TYPES_MASK the
// remaining thread bins (k_bin_scan decides which ranges exist).  They are separate launches: one
// kernel that runs both loops back to back measured 10 % SLOWER than the generic kernel alone
// (2.40e8 vs 2.53e8 vs 2.66e8 pairs/s for the split), the larger kernel image costs more
// than the saved launch tail.
template <int MODE, int TM>
__global__ void __launch_bounds__(GJK_THREADS, GJK_BLOCKS_FOR(TM))
k_gjk_thread(d3d_colliders c, const int32_t *__restrict__ pairs, GjkWorkspace w, GjkParams prm) {
    extern __shared__ real smem[];
    real *base = smem + threadIdx.x;
    Simplex<GJK_THREADS> S;
    S.base = base;
    S.pq = nullptr;
    S.perm = 0;
    WarpScratch W;
    {
        char *scratch = reinterpret_cast<char *>(smem + GJK_FIELDS_THREAD * GJK_THREADS) +
                        (threadIdx.x >> 5) * GJK_SCRATCH_BYTES;
        W.res = reinterpret_cast<TriResult *>(scratch);
        W.set = reinterpret_cast<int *>(scratch + 32 * sizeof(TriResult));
        W.desc = reinterpret_cast<unsigned short *>(scratch + 32 * sizeof(TriResult) + 32 * sizeof(int));
    }
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1;
    const int inst = GJK_INSTANCE_OF(TM);
    const int first = inst == 0 ? 0 : w.counters[inst == 1 ? 4 : 6];
    const int total = w.counters[inst == 0 ? 4 : (inst == 1 ? 6 : 2)];
    int *next = &w.counters[inst == 0 ? 0 : (inst == 1 ? 7 : 5)];
    bool exhausted = false;
#if GJK_CHUNK > 0
    int chunk_pos = 0, chunk_end = 0;
#endif
    PairState<GJK_THREADS> s;
    s.state = D3D_UNKNOWN;
    bool running = false;   // lane owns a pair that still iterates
    bool finished = false;  // lane owns a pair whose result is not written yet

    for (;;) {
        unsigned run_mask = __ballot_sync(0xffffffffu, running);
        int idle = 32 - __popc(run_mask);
        if (idle >= GJK_REFILL_MIN || run_mask == 0) {
            if (finished) {
                if (((TM >> D3D_MESH) & 1) && prm.mesh_last) write_mesh_last(s, pairs, prm.mesh_last);
                gjk_finish_or_park<MODE>(s, S, prm, w);
                finished = false;
            }
            if (!exhausted) {
                unsigned need = ~run_mask;
                int want = __popc(need), rank = __popc(need & lt_mask);
                int mine = -1;
#if GJK_CHUNK > 0
                // warp-private chunk of the sorted order: the lanes of a warp hold neighbouring
                // pairs (one type bin) even when the bins are small next to the number of
                // pairs in flight on the whole device
                int avail = chunk_end - chunk_pos;
                if (rank < avail) mine = chunk_pos + rank;
                if (want > avail) {
                    int start = 0;
                    if (lane == 0) start = atomicAdd(next, GJK_CHUNK);
                    start = __shfl_sync(0xffffffffu, start, 0) + first;
                    if (start >= total) {
                        exhausted = true;
                        chunk_pos = chunk_end = 0;
                    } else {
                        int end = min(start + GJK_CHUNK, total);
                        if (rank >= avail && start + (rank - avail) < end) mine = start + (rank - avail);
                        chunk_pos = min(start + (want - avail), end);
                        chunk_end = end;
                    }
                } else {
                    chunk_pos += want;
                }
#else
                // the idle lanes take the next `want` pairs of the sorted order (one atomic per refill)
                int start = 0;
                if (lane == 0) start = atomicAdd(next, want);
                start = __shfl_sync(0xffffffffu, start, 0) + first;
                exhausted = start + want >= total;
                if (start + rank < total) mine = start + rank;
#endif
                if (!running && mine >= 0) {
                    init_pair<GJK_THREADS>(s, S, c, pairs, __ldg(w.perm + mine), base);
                    if (GJK_PQ_GLOBAL && MODE == 0) S.pq = w.fin + (int64_t)s.k * GJK_FIN_FIELDS + 12;
                    running = true;
                }
            }
            run_mask = __ballot_sync(0xffffffffu, running);
            if (run_mask == 0) break;
        }
        bool solve = running && gjk_pre<MODE, 1, GJK_THREADS, TM>(s, S, prm, 0);
        v3 v_new = V3(R(0.0), R(0.0), R(0.0));
        real v_len_sq_new = R(0.0);
        int simplex = 0;
        bool ok = closest_point_to_origin_warp<GJK_THREADS>(solve, S, s.n_points, s.prev_v_len_sq, W, lane,
                                                            v_new, v_len_sq_new, simplex);
        if (solve) gjk_post<MODE, GJK_THREADS>(s, S, prm, ok, v_new, v_len_sq_new, simplex);
        if (running && s.state != D3D_UNKNOWN) { running = false; finished = true; }
    }
}

// One warp per pair (wide hulls); every lane holds the same state, the shared
// record is written redundantly (same values) by all lanes.
template <int MODE>
__global__ void __launch_bounds__(GJK_THREADS, 4)
k_gjk_warp(d3d_colliders c, const int32_t *__restrict__ pairs, GjkWorkspace w, GjkParams prm) {
    __shared__ real smem[(GJK_THREADS / 32) * GJK_FIELDS];
    const int lane = threadIdx.x & 31;
    real *base = smem + (threadIdx.x >> 5) * GJK_FIELDS;
    Simplex<1> S;
    S.base = base;
    S.pq = nullptr;
    const int first = w.counters[2], total = w.counters[3];
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(&w.counters[1], 1);
        idx = __shfl_sync(0xffffffffu, idx, 0) + first;
        if (idx >= total) break;
        PairState<1> s;
        init_pair<1>(s, S, c, pairs, __ldg(w.perm + idx), base);
        __syncwarp();
        while (s.state == D3D_UNKNOWN) {
            gjk_step<MODE, 32, 1>(s, S, prm, lane);
            __syncwarp();
        }
        if (prm.mesh_last && lane == 0) write_mesh_last(s, pairs, prm.mesh_last);
        gjk_finish<MODE>(s, S, prm, lane == 0);
        __syncwarp();
    }
}

template <int MODE>
int launch_gjk(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs, const GjkParams &prm,
               void *workspace, size_t ws_bytes, cudaStream_t stream) {
    if (n_pairs == 0) return 0;
    if (n_pairs > 0x7fffffff) return d3d_set_error("d3d_gjk: more than 2^31-1 pairs in one call");
    if (ws_bytes < gjk_ws_bytes(n_pairs, MODE == 0)) return d3d_set_error("d3d_gjk: workspace too small");
    GjkWorkspace w = carve(workspace, n_pairs);
    D3D_CUDA_CHECK(cudaMemsetAsync(workspace, 0, 4096, stream));
    int sms = d3d_sm_count();
    int bin_blocks = (int)d3d_min64((n_pairs + 255) / 256, (int64_t)sms * 8);
    static int hull_max = -1;
    if (hull_max < 0) {
        const char *e = getenv("D3D_THREAD_HULL_MAX");
        hull_max = e ? atoi(e) : D3D_THREAD_HULL_MAX_DEFAULT;
    }
    k_pair_keys<<<bin_blocks, 256, 0, stream>>>(*c, pairs, n_pairs, w, hull_max);
    // smallest group of bins that gets a launch of its own (tests force 1 through the environment
    // so that small batches exercise all three thread instances)
    int split_min = sms * GJK_BLOCKS_PER_SM * GJK_THREADS * 8;
    if (const char *e = getenv("D3D_GJK_SPLIT_MIN")) split_min = atoi(e);
    k_bin_scan<<<1, 32, 0, stream>>>(w, n_pairs, split_min);
    k_bin_scatter<<<(int)d3d_min64((n_pairs + 256 * BIN_TILE_ITEMS - 1) / (256 * BIN_TILE_ITEMS), (int64_t)sms * 8), 256, 0, stream>>>(n_pairs, w);
    size_t smem = sizeof(real) * GJK_FIELDS_THREAD * GJK_THREADS + (GJK_THREADS / 32) * GJK_SCRATCH_BYTES;
    // function attributes are per device: remember which devices have been set up
    static bool attr_set[64] = {false};  // per MODE instance of this function
    int dev = 0;
    D3D_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        D3D_CUDA_CHECK(cudaFuncSetAttribute(k_gjk_thread<MODE, D3D_PRIMITIVE_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        D3D_CUDA_CHECK(cudaFuncSetAttribute(k_gjk_thread<MODE, GJK_CONVEX_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        D3D_CUDA_CHECK(cudaFuncSetAttribute(k_gjk_thread<MODE, D3D_ALL_TYPES_MASK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    int blocks = (int)d3d_min64((n_pairs + GJK_THREADS - 1) / GJK_THREADS, (int64_t)sms * GJK_BLOCKS_PER_SM);
    int blocks_prim = (int)d3d_min64((n_pairs + GJK_THREADS - 1) / GJK_THREADS, (int64_t)sms * GJK_BLOCKS_PRIM);
    // ranges are read on the device; an instance whose range is empty exits at once
    k_gjk_thread<MODE, D3D_PRIMITIVE_MASK><<<blocks_prim, GJK_THREADS, smem, stream>>>(*c, pairs, w, prm);
    int blocks_convex = (int)d3d_min64((n_pairs + GJK_THREADS - 1) / GJK_THREADS, (int64_t)sms * GJK_BLOCKS_CONVEX);
    k_gjk_thread<MODE, GJK_CONVEX_MASK><<<blocks_convex, GJK_THREADS, smem, stream>>>(*c, pairs, w, prm);
    k_gjk_thread<MODE, D3D_ALL_TYPES_MASK><<<blocks, GJK_THREADS, smem, stream>>>(*c, pairs, w, prm);
    if (MODE == 0)
        k_gjk_finish<<<(int)d3d_min64((n_pairs + 255) / 256, (int64_t)sms * 8), 256, 0, stream>>>(w, prm);
    int wblocks = (int)d3d_min64((n_pairs + 3) / 4, (int64_t)sms * 3);
    k_gjk_warp<MODE><<<wblocks, GJK_THREADS, 0, stream>>>(*c, pairs, w, prm);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace

// The fp32 instantiation of this file (gjk_f32.cu: -DD3D_F32) exports *_f32 entry points.
#ifdef D3D_F32
#define D3D_GJK_SYM(name) name##_f32
#else
#define D3D_GJK_SYM(name) name
#endif

extern "C" {

#ifndef D3D_F32
size_t d3d_gjk_workspace_bytes(int64_t n_pairs) { return gjk_ws_bytes(n_pairs, true); }
size_t d3d_gjk_intersection_workspace_bytes(int64_t n_pairs) { return gjk_ws_bytes(n_pairs, false); }
#endif

int D3D_GJK_SYM(d3d_gjk_distance)(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs,
                     double tolerance, double max_distance_squared, double sanity_check,
                     double *out_dist, double *out_a, double *out_b, double *out_Y,
                     int32_t *out_npoints, int32_t *out_iters, int32_t *out_status,
                     void *workspace, size_t ws_bytes, void *stream) {
    if (n_pairs == 0) return 0;
    if (!c || !pairs || !out_dist) return d3d_set_error("d3d_gjk_distance: null argument");
    GjkParams prm;
    prm.tolerance_sq = tolerance * tolerance;
    prm.max_distance_squared = max_distance_squared;
    prm.sanity_check = sanity_check;
    prm.out_dist = out_dist; prm.out_a = out_a; prm.out_b = out_b; prm.out_Y = out_Y;
    prm.out_npoints = out_npoints; prm.out_iters = out_iters; prm.out_status = out_status;
    prm.out_hit = nullptr;
    prm.graph = c->graph; prm.mesh_last = c->mesh_last;
    return launch_gjk<0>(c, pairs, n_pairs, prm, workspace, ws_bytes, (cudaStream_t)stream);
}

int D3D_GJK_SYM(d3d_gjk_intersection)(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs,
                         double tolerance, uint8_t *out_hit, int32_t *out_iters,
                         int32_t *out_status, void *workspace, size_t ws_bytes, void *stream) {
    if (n_pairs == 0) return 0;
    if (!c || !pairs || !out_hit) return d3d_set_error("d3d_gjk_intersection: null argument");
    GjkParams prm = {};
    prm.tolerance_sq = tolerance * tolerance;
    prm.out_hit = out_hit; prm.out_iters = out_iters; prm.out_status = out_status;
    prm.graph = c->graph; prm.mesh_last = c->mesh_last;
    return launch_gjk<1>(c, pairs, n_pairs, prm, workspace, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
