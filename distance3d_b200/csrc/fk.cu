// Batched forward kinematics of a URDF kinematic tree and the candidate-pair
// plumbing of robot self-collision (BASELINE config 4).
//
// The reference obtains collider poses one joint configuration at a time from
// pytransform3d (UrdfTransformManager.set_joint / get_transform; call sites
// distance3d/broad_phase.py:111,148) and then, per collider, asks the AABB tree for
// candidates, drops white-listed frames and runs gjk_intersection
// (distance3d/self_collision.py:5-64).  Here
//   k_fk           one thread per (configuration, frame): pose = prod_s fixed_s * joint_s(q)
//   k_filter_pairs tests a fixed list of candidate (frame, frame) pairs per
//                  configuration with the closed AABB predicate (aabb_tree.py:503-527)
//                  and compacts the overlapping ones (warp-aggregated append)
//   k_scatter_hits flags both colliders of every intersecting pair
// and the narrow phase in between is d3d_gjk_intersection.
#include "d3d_common.cuh"
#include "d3d_aabb.cuh"

namespace {

struct FkModel {
    int n_frames, n_joints;
    const double *joint_axis;    // [J,3] unit axes
    const double *joint_limits;  // [J,2]
    const int32_t *joint_type;   // [J] 0 revolute, 1 prismatic
    const int32_t *chain_off;    // [K+1]
    const double *chain_fixed;   // [S,4,4]
    const int32_t *chain_joint;  // [S] joint applied after the fixed transform, -1 none
};

__device__ __forceinline__ void mat_mul(const double *A, const double *B, double *C) {
    // rows 0..2 of two rigid transforms (row 3 = 0 0 0 1)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int col = 0; col < 4; ++col) {
            double s = A[4 * r] * B[col] + A[4 * r + 1] * B[4 + col] + A[4 * r + 2] * B[8 + col];
            if (col == 3) s += A[4 * r + 3];
            C[4 * r + col] = s;
        }
    }
}

__global__ void k_fk(FkModel m, const double *__restrict__ q, int64_t n_cfg, double *__restrict__ out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_cfg * m.n_frames) return;
    int64_t b = t / m.n_frames;
    int k = (int)(t % m.n_frames);
    double T[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    double U[12], J[12];
    for (int s = m.chain_off[k]; s < m.chain_off[k + 1]; ++s) {
        const double *F = m.chain_fixed + 16 * (int64_t)s;
        mat_mul(T, F, U);
        int j = m.chain_joint[s];
        if (j >= 0) {
            double v = q[b * m.n_joints + j];
            v = fmin(fmax(v, m.joint_limits[2 * j]), m.joint_limits[2 * j + 1]);
            double ux = m.joint_axis[3 * j], uy = m.joint_axis[3 * j + 1], uz = m.joint_axis[3 * j + 2];
            if (m.joint_type[j] == 0) {  // Rodrigues rotation about the unit axis
                double c = cos(v), sn = sin(v), ci = 1.0 - c;
                J[0] = ci * ux * ux + c;       J[1] = ci * ux * uy - uz * sn; J[2] = ci * ux * uz + uy * sn;  J[3] = 0.0;
                J[4] = ci * uy * ux + uz * sn; J[5] = ci * uy * uy + c;       J[6] = ci * uy * uz - ux * sn;  J[7] = 0.0;
                J[8] = ci * uz * ux - uy * sn; J[9] = ci * uz * uy + ux * sn; J[10] = ci * uz * uz + c;       J[11] = 0.0;
            } else {
                J[0] = 1; J[1] = 0; J[2] = 0; J[3] = v * ux;
                J[4] = 0; J[5] = 1; J[6] = 0; J[7] = v * uy;
                J[8] = 0; J[9] = 0; J[10] = 1; J[11] = v * uz;
            }
            mat_mul(U, J, T);
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) T[i] = U[i];
        }
    }
    double2 *o = reinterpret_cast<double2 *>(out + 16 * t);
#pragma unroll
    for (int i = 0; i < 6; ++i) o[i] = make_double2(T[2 * i], T[2 * i + 1]);
    o[6] = make_double2(0.0, 0.0);
    o[7] = make_double2(0.0, 1.0);
}

// The same product with the chains' common prefixes evaluated once: the steps of all frames form
// a tree (urdf.compile_kinematics: a node = (parent node, fixed transform, joint), two chains
// share a node when they agree in every step up to it), one thread walks the tree for one
// configuration.  Every frame's pose is produced by exactly the operations k_fk performs for it
// (same matrices, same order), so the results are bit-identical; a 6-joint arm with 8 collider
// frames needs 6 joint steps per configuration instead of 27 (sin / cos in fp64 dominate).
#define FK_MAX_KEEP 16  // nodes with children: their pose stays in the thread's local store
struct FkTree {
    int n_nodes, n_frames, n_joints;
    const double *joint_axis, *joint_limits;
    const int32_t *joint_type;
    const int32_t *node_parent;  // [N] -1 = base frame
    const double *node_fixed;    // [N,4,4]
    const int32_t *node_joint;   // [N] -1 none
    const int32_t *node_keep;    // [N] slot in the local store, -1 = no children
    const int32_t *node_out_off; // [N+1] frames that END at the node ...
    const int32_t *node_out;     // ... their indices
};

__global__ void __launch_bounds__(128) k_fk_tree(FkTree m, const double *__restrict__ q, int64_t n_cfg,
                                                 double *__restrict__ out) {
    int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (b >= n_cfg) return;
    double S[FK_MAX_KEEP][12];
    double T[12], U[12], J[12];
    for (int node = 0; node < m.n_nodes; ++node) {
        const int parent = m.node_parent[node];
        if (parent < 0) {
            const double I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
#pragma unroll
            for (int i = 0; i < 12; ++i) T[i] = I[i];
        } else {
            const double *P = S[m.node_keep[parent]];
#pragma unroll
            for (int i = 0; i < 12; ++i) T[i] = P[i];
        }
        mat_mul(T, m.node_fixed + 16 * (int64_t)node, U);
        const int j = m.node_joint[node];
        if (j >= 0) {
            double v = q[b * m.n_joints + j];
            v = fmin(fmax(v, m.joint_limits[2 * j]), m.joint_limits[2 * j + 1]);
            double ux = m.joint_axis[3 * j], uy = m.joint_axis[3 * j + 1], uz = m.joint_axis[3 * j + 2];
            if (m.joint_type[j] == 0) {  // Rodrigues rotation about the unit axis
                double c = cos(v), sn = sin(v), ci = 1.0 - c;
                J[0] = ci * ux * ux + c;       J[1] = ci * ux * uy - uz * sn; J[2] = ci * ux * uz + uy * sn;  J[3] = 0.0;
                J[4] = ci * uy * ux + uz * sn; J[5] = ci * uy * uy + c;       J[6] = ci * uy * uz - ux * sn;  J[7] = 0.0;
                J[8] = ci * uz * ux - uy * sn; J[9] = ci * uz * uy + ux * sn; J[10] = ci * uz * uz + c;       J[11] = 0.0;
            } else {
                J[0] = 1; J[1] = 0; J[2] = 0; J[3] = v * ux;
                J[4] = 0; J[5] = 1; J[6] = 0; J[7] = v * uy;
                J[8] = 0; J[9] = 0; J[10] = 1; J[11] = v * uz;
            }
            mat_mul(U, J, T);
        } else {
#pragma unroll
            for (int i = 0; i < 12; ++i) T[i] = U[i];
        }
        const int keep = m.node_keep[node];
        if (keep >= 0) {
#pragma unroll
            for (int i = 0; i < 12; ++i) S[keep][i] = T[i];
        }
        for (int e = m.node_out_off[node]; e < m.node_out_off[node + 1]; ++e) {
            double2 *o = reinterpret_cast<double2 *>(out + 16 * (b * m.n_frames + m.node_out[e]));
#pragma unroll
            for (int i = 0; i < 6; ++i) o[i] = make_double2(T[2 * i], T[2 * i + 1]);
            o[6] = make_double2(0.0, 0.0);
            o[7] = make_double2(0.0, 1.0);
        }
    }
}

__device__ __forceinline__ void append_pair(int a, int b, int32_t *out_pairs, int64_t cap,
                                            unsigned long long *count) {
    unsigned m = __activemask();
    int lane = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    unsigned long long pos = base + __popc(m & ((1u << lane) - 1));
    if ((int64_t)pos < cap) reinterpret_cast<int2 *>(out_pairs)[pos] = make_int2(a, b);
}

// candidate pairs pattern[n_pattern,2] (frame indices) replicated over n_groups groups of
// group_size consecutive boxes; emits global box indices of the overlapping pairs
__global__ void k_filter_pairs(const double *__restrict__ aabb, int64_t n_groups, int group_size,
                               const int32_t *__restrict__ pattern, int n_pattern,
                               int32_t *out_pairs, int64_t cap, unsigned long long *count) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_groups * n_pattern) return;
    int64_t g = t / n_pattern;
    int p = (int)(t % n_pattern);
    int64_t ia = g * group_size + pattern[2 * p], ib = g * group_size + pattern[2 * p + 1];
    const double2 *a = reinterpret_cast<const double2 *>(aabb + 6 * ia);
    const double2 *b = reinterpret_cast<const double2 *>(aabb + 6 * ib);
    bool ov = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double2 x = __ldg(a + k), y = __ldg(b + k);
        ov = ov && x.x <= y.y && x.y >= y.x;
    }
    if (ov) append_pair((int)ia, (int)ib, out_pairs, cap, count);
}

// k_aabb + k_filter_pairs in one pass for the self-collision step: a block takes G groups
// (configurations) of group_size colliders, one thread per collider computes its box (stored to
// `aabb` as d3d_aabb would, and to shared memory), then the block tests the G x n_pattern candidate
// pairs from shared memory.  The boxes are not re-read from global memory and one launch goes away.
#define FILTER_THREADS 128
__global__ void __launch_bounds__(FILTER_THREADS)
k_aabb_filter(d3d_colliders c, int64_t n_groups, int group_size, int groups_per_block,
              const int32_t *__restrict__ pattern, int n_pattern, double *__restrict__ aabb,
              int32_t *out_pairs, int64_t cap, unsigned long long *count) {
    __shared__ double box[FILTER_THREADS * 6];
    const int64_t g0 = blockIdx.x * (int64_t)groups_per_block;
    const int n_local = (int)d3d_min64(groups_per_block, n_groups - g0) * group_size;
    if ((int)threadIdx.x < n_local) {
        const int64_t i = g0 * group_size + threadIdx.x;
        double lo[3], hi[3];
        collider_aabb(c, i, lo, hi);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            box[6 * threadIdx.x + 2 * k] = lo[k];
            box[6 * threadIdx.x + 2 * k + 1] = hi[k];
        }
        if (aabb) {
            double2 *o = reinterpret_cast<double2 *>(aabb + 6 * i);
            o[0] = make_double2(lo[0], hi[0]);
            o[1] = make_double2(lo[1], hi[1]);
            o[2] = make_double2(lo[2], hi[2]);
        }
    }
    __syncthreads();
    const int n_tests = (n_local / group_size) * n_pattern;
    for (int t0 = 0; t0 < n_tests; t0 += FILTER_THREADS) {  // uniform trip count: append_pair uses __activemask
        const int t = t0 + threadIdx.x;
        if (t < n_tests) {
            const int g = t / n_pattern, p = t % n_pattern;
            const int a = g * group_size + __ldg(pattern + 2 * p), b = g * group_size + __ldg(pattern + 2 * p + 1);
            bool ov = true;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                ov = ov && box[6 * a + 2 * k] <= box[6 * b + 2 * k + 1] && box[6 * a + 2 * k + 1] >= box[6 * b + 2 * k];
            if (ov) append_pair((int)(g0 * group_size) + a, (int)(g0 * group_size) + b, out_pairs, cap, count);
        }
    }
}

__global__ void k_scatter_hits(const int32_t *__restrict__ pairs, const uint8_t *__restrict__ hit,
                               const unsigned long long *count, int64_t cap, uint8_t *mask) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t n = (int64_t)*count;
    if (n > cap) n = cap;
    if (t >= n || !hit[t]) return;
    int2 pr = reinterpret_cast<const int2 *>(pairs)[t];
    mask[pr.x] = 1;
    mask[pr.y] = 1;
}

// ---------------------------------------------------------------------------
// self_collision.detect with the reference's ORDER (self_collision.py:22-36).  With
// white-lists that are not symmetric (a link with several child links white-lists only one
// of them, urdf_utils.py:79-81) the reference's result depends on the order in which
// `aabb_overlapping_colliders` lists the candidates of a frame: the first intersecting one is
// flagged together with the frame, then the loop breaks.  That order is the depth-first
// order of the reference's incremental AABB tree (aabb_tree.py:194-341 insertion, :381-403
// query: children pushed left then right, popped right first), rebuilt for every pose update
// by inserting the colliders one at a time (broad_phase.py:144-151).  One thread per joint
// configuration rebuilds that tree over its K boxes and replays the loop on the precomputed
// intersection bits.
#define D3D_DETECT_MAX_K 64

__global__ void k_scatter_hit_bits(const int32_t *__restrict__ pairs, const uint8_t *__restrict__ hit,
                                   const unsigned long long *count, int64_t cap, int group_size,
                                   unsigned long long *bits) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t n = (int64_t)*count;
    if (n > cap) n = cap;
    if (t >= n || !hit[t]) return;
    int2 pr = reinterpret_cast<const int2 *>(pairs)[t];
    atomicOr(bits + pr.x, 1ull << (pr.y % group_size));
    atomicOr(bits + pr.y, 1ull << (pr.x % group_size));
}

struct Box6 { double v[6]; };  // lo_x hi_x lo_y hi_y lo_z hi_z (the (3,2) layout)

__device__ __forceinline__ Box6 merge_boxes(const Box6 &a, const Box6 &b) {  // aabb_tree.py:536-551
    Box6 o;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.v[2 * k] = a.v[2 * k] < b.v[2 * k] ? a.v[2 * k] : b.v[2 * k];
        o.v[2 * k + 1] = a.v[2 * k + 1] > b.v[2 * k + 1] ? a.v[2 * k + 1] : b.v[2 * k + 1];
    }
    return o;
}
__device__ __forceinline__ double box_volume(const Box6 &m) {
    return (m.v[1] - m.v[0]) * (m.v[3] - m.v[2]) * (m.v[5] - m.v[4]);
}
__device__ __forceinline__ bool boxes_overlap(const Box6 &a, const Box6 &b) {  // aabb_tree.py:503-527
    return a.v[0] <= b.v[1] && a.v[1] >= b.v[0] && a.v[2] <= b.v[3] && a.v[3] >= b.v[2] &&
           a.v[4] <= b.v[5] && a.v[5] >= b.v[4];
}

template <int MAXK>
__global__ void __launch_bounds__(64)
k_detect_ordered(const double *__restrict__ aabb, int64_t n_groups, int K,
                 const unsigned long long *__restrict__ bits,
                 const unsigned long long *__restrict__ whitelist, uint8_t *mask) {
    int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    // leaves 0..K-1, branches K..2K-2
    Box6 box[2 * MAXK - 1];
    signed char parent[2 * MAXK - 1], left[2 * MAXK - 1], right[2 * MAXK - 1];
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int c = 0; c < 6; ++c) box[i].v[c] = __ldg(aabb + 6 * (g * K + i) + c);
    int root = -1, filled = K;
    for (int leaf = 0; leaf < K; ++leaf) {
        left[leaf] = right[leaf] = -1;
        if (root < 0) { root = leaf; parent[leaf] = -1; continue; }
        int t = root;
        while (t >= K) {  // descend into the child whose merged box is smaller, ties go right
            double cl = box_volume(merge_boxes(box[leaf], box[left[t]]));
            double cr = box_volume(merge_boxes(box[leaf], box[right[t]]));
            t = (cl < cr) ? left[t] : right[t];
        }
        int sib = t, old = parent[sib], np_ = filled++;
        parent[np_] = (signed char)old; left[np_] = (signed char)sib; right[np_] = (signed char)leaf;
        box[np_] = merge_boxes(box[leaf], box[sib]);
        parent[leaf] = parent[sib] = (signed char)np_;
        if (old < 0) root = np_;
        else if (left[old] == sib) left[old] = (signed char)np_;
        else right[old] = (signed char)np_;
        for (int u = parent[np_]; u >= 0; u = parent[u]) box[u] = merge_boxes(box[left[u]], box[right[u]]);
    }
    unsigned long long marked = 0;
    signed char stack[2 * MAXK];
    for (int f = 0; f < K; ++f) {
        if ((marked >> f) & 1) continue;  // contact was detected before
        const unsigned long long allowed = ~__ldg(whitelist + f);
        const unsigned long long hits = __ldg(bits + g * K + f) | (1ull << f);  // a collider intersects itself
        int sp = 0;
        stack[sp++] = (signed char)root;
        while (sp) {
            int n = stack[--sp];
            if (!boxes_overlap(box[n], box[f])) continue;
            if (n < K) {
                if (((allowed >> n) & 1) && ((hits >> n) & 1)) {
                    marked |= (1ull << f) | (1ull << n);
                    break;
                }
            } else {
                stack[sp++] = left[n];
                stack[sp++] = right[n];
            }
        }
    }
    for (int f = 0; f < K; ++f) mask[g * K + f] = (uint8_t)((marked >> f) & 1);
}

}  // namespace

extern "C" {

int d3d_detect_ordered(const double *aabb, int64_t n_groups, int group_size, const int32_t *pairs,
                       const uint8_t *hit, const unsigned long long *count, int64_t cap,
                       const unsigned long long *whitelist, unsigned long long *hit_bits,
                       uint8_t *mask, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_groups == 0 || group_size == 0) return 0;
    if (group_size > D3D_DETECT_MAX_K) return d3d_set_error("d3d_detect_ordered: more than 64 colliders per group");
    if (!aabb || !whitelist || !hit_bits || !mask || (cap > 0 && (!pairs || !hit || !count)))
        return d3d_set_error("d3d_detect_ordered: null argument");
    D3D_CUDA_CHECK(cudaMemsetAsync(hit_bits, 0, sizeof(unsigned long long) * n_groups * group_size, stream));
    if (cap > 0)
        k_scatter_hit_bits<<<(unsigned)((cap + 255) / 256), 256, 0, stream>>>(pairs, hit, count, cap, group_size, hit_bits);
    unsigned blocks = (unsigned)((n_groups + 63) / 64);
    if (group_size <= 16)
        k_detect_ordered<16><<<blocks, 64, 0, stream>>>(aabb, n_groups, group_size, hit_bits, whitelist, mask);
    else
        k_detect_ordered<D3D_DETECT_MAX_K><<<blocks, 64, 0, stream>>>(aabb, n_groups, group_size, hit_bits, whitelist, mask);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_fk_urdf(int n_frames, int n_joints, const double *joint_axis, const double *joint_limits,
                const int32_t *joint_type, const int32_t *chain_off, const double *chain_fixed,
                const int32_t *chain_joint, const double *q, int64_t n_cfg, double *out_pose,
                void *stream) {
    if (n_cfg == 0 || n_frames == 0) return 0;
    if (!chain_off || !chain_fixed || !chain_joint || !q || !out_pose)
        return d3d_set_error("d3d_fk_urdf: null argument");
    FkModel m;
    m.n_frames = n_frames; m.n_joints = n_joints; m.joint_axis = joint_axis;
    m.joint_limits = joint_limits; m.joint_type = joint_type; m.chain_off = chain_off;
    m.chain_fixed = chain_fixed; m.chain_joint = chain_joint;
    int64_t threads = n_cfg * n_frames;
    k_fk<<<(unsigned)((threads + 127) / 128), 128, 0, (cudaStream_t)stream>>>(m, q, n_cfg, out_pose);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_fk_urdf_tree(int n_frames, int n_joints, const double *joint_axis, const double *joint_limits,
                     const int32_t *joint_type, int n_nodes, int n_keep, const int32_t *node_parent,
                     const double *node_fixed, const int32_t *node_joint, const int32_t *node_keep,
                     const int32_t *node_out_off, const int32_t *node_out, const double *q, int64_t n_cfg,
                     double *out_pose, void *stream) {
    if (n_cfg == 0 || n_frames == 0) return 0;
    if (!node_parent || !node_fixed || !node_joint || !node_keep || !node_out_off || !node_out || !q || !out_pose)
        return d3d_set_error("d3d_fk_urdf_tree: null argument");
    if (n_keep > FK_MAX_KEEP)
        return d3d_set_error("d3d_fk_urdf_tree: %d chain nodes with children, at most %d (use d3d_fk_urdf)", n_keep, FK_MAX_KEEP);
    FkTree m;
    m.n_nodes = n_nodes; m.n_frames = n_frames; m.n_joints = n_joints; m.joint_axis = joint_axis;
    m.joint_limits = joint_limits; m.joint_type = joint_type; m.node_parent = node_parent;
    m.node_fixed = node_fixed; m.node_joint = node_joint; m.node_keep = node_keep;
    m.node_out_off = node_out_off; m.node_out = node_out;
    k_fk_tree<<<(unsigned)((n_cfg + 127) / 128), 128, 0, (cudaStream_t)stream>>>(m, q, n_cfg, out_pose);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_filter_pairs(const double *aabb, int64_t n_groups, int group_size, const int32_t *pattern,
                     int n_pattern, int32_t *out_pairs, int64_t cap, unsigned long long *out_count,
                     void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!out_count) return d3d_set_error("d3d_filter_pairs: null argument");
    D3D_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(unsigned long long), stream));
    int64_t threads = n_groups * n_pattern;
    if (threads == 0) return 0;
    if (!aabb || !pattern || !out_pairs) return d3d_set_error("d3d_filter_pairs: null argument");
    k_filter_pairs<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(aabb, n_groups, group_size, pattern,
                                                                        n_pattern, out_pairs, cap, out_count);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_aabb_filter_pairs(const d3d_colliders *c, int64_t n_groups, int group_size, const int32_t *pattern,
                          int n_pattern, double *out_aabb, int32_t *out_pairs, int64_t cap,
                          unsigned long long *out_count, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!out_count) return d3d_set_error("d3d_aabb_filter_pairs: null argument");
    D3D_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(unsigned long long), stream));
    if (n_groups == 0) return 0;
    if (!c || !pattern || !out_pairs) return d3d_set_error("d3d_aabb_filter_pairs: null argument");
    if (group_size < 1 || group_size > FILTER_THREADS)
        return d3d_set_error("d3d_aabb_filter_pairs: group_size must be in [1, %d] (use d3d_aabb + d3d_filter_pairs)", FILTER_THREADS);
    if (c->n != n_groups * group_size) return d3d_set_error("d3d_aabb_filter_pairs: collider count != n_groups * group_size");
    const int gpb = FILTER_THREADS / group_size;
    k_aabb_filter<<<(unsigned)((n_groups + gpb - 1) / gpb), FILTER_THREADS, 0, stream>>>(
        *c, n_groups, group_size, gpb, pattern, n_pattern, out_aabb, out_pairs, cap, out_count);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_scatter_hits(const int32_t *pairs, const uint8_t *hit, const unsigned long long *count,
                     int64_t cap, uint8_t *mask, void *stream) {
    if (cap == 0) return 0;
    if (!pairs || !hit || !count || !mask) return d3d_set_error("d3d_scatter_hits: null argument");
    k_scatter_hits<<<(unsigned)((cap + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pairs, hit, count, cap, mask);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
