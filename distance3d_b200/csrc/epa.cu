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
    const int *perm;  // processing order: pairs grouped by (typeA, typeB)
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

__global__ void k_epa_scan(EpaOrder w) {
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < EPA_NBINS; ++i) { w.cursor[i] = acc; acc += w.hist[i]; }
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
inline size_t epa_ws_bytes(int64_t n) { return 2048 + epa_au((size_t)n) + epa_au((size_t)n * 4); }

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
#define EPA_BLOCKS_PER_SM 6  // 80 registers, 24 warps per SM; r02 sweep (scripts/r02_run2.sh) C3 / C5 ms at 4,5,6,8: 7.3 7.7 7.2 7.8 / 493 491 481 494
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

    for (;;) {
        int k = 0;
        if (lane == 0) {
            k = atomicAdd(prm.counter, 1);
            k = k < n_pairs ? __ldg(prm.perm + k) : -1;
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
            // ---- B: support point of A - B in the face normal (epa.py:62-65)
            v3 sd = W.fget(closest, 3);
            v3 new_point = support_call<32, 1, D3D_ALL_TYPES_MASK>(A.type, A.nv, A.V, recA, c.graph, sd.x, sd.y, sd.z, lane) -
                           support_call<32, 1, D3D_ALL_TYPES_MASK>(B.type, B.nv, B.V, recB, c.graph, -sd.x, -sd.y, -sd.z, lane);
            __syncwarp();  // MeshGraph: lane 0 stored the vertex the climb ended on
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
#pragma unroll 1
                    for (int j = 0; j < 3; ++j) {
                        v3 e0 = j == 0 ? fv0 : (j == 1 ? fv1 : fv2), e1 = j == 0 ? fv1 : (j == 1 ? fv2 : fv0);
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
                            if (n_loose >= ml) break;
                            if (lane == n_loose) { la = e0; lb = e1; }
                            ++n_loose;
                        }
                    }
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
        __syncwarp();
    }
}

// smallest double t with sqrt(t) >= x (sqrt is correctly rounded and monotone on host and device)
double sqrt_threshold(double x) {
    double t = x * x;
    while (t > 0.0 && sqrt(nextafter(t, 0.0)) >= x) t = nextafter(t, 0.0);
    while (sqrt(t) < x) t = nextafter(t, 1e308);
    return t;
}

}  // namespace

extern "C" {

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
    {
        int sms = d3d_sm_count();
        int kb = (int)d3d_min64((n_pairs + 255) / 256, (int64_t)sms * 8);
        k_epa_keys<<<kb, 256, 0, stream>>>(*c, pairs, npoints, n_pairs, ord);
        k_epa_scan<<<1, 32, 0, stream>>>(ord);
        k_epa_scatter<<<(int)d3d_min64((n_pairs + 2047) / 2048, (int64_t)sms * 8), 256, 0, stream>>>(n_pairs, ord);
    }
    size_t per_warp = (size_t)12 * max_faces + (max_faces + 1) / 2 + 2 + 2 * D3D_COLLIDER_FIELDS;
    size_t smem = per_warp * EPA_WARPS * sizeof(double);
    int blocks = (int)d3d_min64((n_pairs + EPA_WARPS - 1) / EPA_WARPS, (int64_t)d3d_sm_count() * (EPA_BLOCKS_PER_SM + 2));
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
