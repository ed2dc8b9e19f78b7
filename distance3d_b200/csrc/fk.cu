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

}  // namespace

extern "C" {

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

int d3d_scatter_hits(const int32_t *pairs, const uint8_t *hit, const unsigned long long *count,
                     int64_t cap, uint8_t *mask, void *stream) {
    if (cap == 0) return 0;
    if (!pairs || !hit || !count || !mask) return d3d_set_error("d3d_scatter_hits: null argument");
    k_scatter_hits<<<(unsigned)((cap + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pairs, hit, count, cap, mask);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
