// C-ABI plumbing (errors, device info) and the small per-collider kernels:
// box vertex generation, batched support map evaluation, bounding boxes.
#include <stdarg.h>
#include <string.h>

#include "d3d_common.cuh"
#include "d3d_support.cuh"
#include "d3d_aabb.cuh"

static thread_local char g_error[512] = "";

int d3d_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return -1;
}

int d3d_sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

namespace {

// geometry.py:138-157 convert_box_to_vertices for every box of the set
__global__ void k_prepare(d3d_colliders c, double *verts_out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t i = t >> 3;
    int v = (int)(t & 7);
    if (i >= c.n || c.type[i] != D3D_BOX) return;
    Collider col = load_collider(c, i);
    st3(verts_out + 3 * ((int64_t)c.vert_off[i] + v), box_vertex(col, v));
}

// Offsets of the wire records from their types alone (exclusive scan of the record sizes), so
// that the 4-byte offset per collider does not have to cross PCIe: tiles of 1024 colliders, tile
// sums -> one-block scan of the sums -> offsets inside every tile.
__device__ __forceinline__ int wire_doubles_dev(int t) {
    // sphere capsule box ellipsoid cylinder hull mesh disk ellipse cone (d3d_pack_wire_host)
    const unsigned lo = 0x2E7C1C4u;  // 5 bits each, types 0..5:  4 14 16 15 14 1
    const unsigned hi = 0x72CEDu;                 //              types 6..9: 13  7 11 14
    return t < 6 ? (int)((lo >> (5 * t)) & 31u) : (int)((hi >> (5 * (t - 6))) & 31u);
}
#define WIRE_TILE 1024
__global__ void __launch_bounds__(256) k_wire_tile_sums(const uint8_t *__restrict__ wtype, int64_t n, int32_t *tile_sum) {
    __shared__ int warp_sum[8];
    const int64_t base = blockIdx.x * (int64_t)WIRE_TILE + 4 * threadIdx.x;
    int s = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (base + u < n) s += wire_doubles_dev(wtype[base + u]);
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += warp_sum[w];
        tile_sum[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) k_wire_scan_sums(int32_t *tile_sum, int64_t n_tiles) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < n_tiles; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const int v = i < n_tiles ? tile_sum[i] : 0;
        int incl = v;
        for (int off = 1; off < 32; off <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, incl, off);
            if ((threadIdx.x & 31) >= off) incl += y;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_tot[threadIdx.x], wi = w;
            for (int off = 1; off < 32; off <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, wi, off);
                if ((int)threadIdx.x >= off) wi += y;
            }
            warp_tot[threadIdx.x] = wi - w;  // exclusive
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + warp_tot[threadIdx.x >> 5] + incl - v;
        if (i < n_tiles) tile_sum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_wire_offsets(const uint8_t *__restrict__ wtype, int64_t n,
                                                      const int32_t *__restrict__ tile_base, int32_t *woff) {
    __shared__ int warp_tot[8];
    const int64_t base = blockIdx.x * (int64_t)WIRE_TILE + 4 * threadIdx.x;
    int sz[4], s = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        sz[u] = base + u < n ? wire_doubles_dev(wtype[base + u]) : 0;
        s += sz[u];
    }
    int incl = s;
    for (int off = 1; off < 32; off <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, off);
        if ((threadIdx.x & 31) >= off) incl += y;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    int acc = tile_base[blockIdx.x] + incl - s;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) acc += warp_tot[w];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (base + u < n) woff[base + u] = acc;
        acc += sz[u];
    }
}

// Compact wire records -> the structure-of-arrays collider set (include/d3d_b200.h
// d3d_unpack_colliders).  One thread per collider; the record is at wire[wire_off[i]].
__global__ void k_unpack(const uint8_t *__restrict__ wtype, const int32_t *__restrict__ woff,
                         const double *__restrict__ wire, int64_t n, int32_t *type, double *pose,
                         double *param, int32_t *vert_off, int32_t *vert_len) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = wtype[i];
    const double *r = wire + woff[i];
    double T[12] = {1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 0.0};
    double p[3] = {0.0, 0.0, 0.0};
    int2 vr = make_int2(0, 0);
    const bool full_pose = t == D3D_CAPSULE || t == D3D_CYLINDER || t == D3D_ELLIPSOID || t == D3D_BOX ||
                           t == D3D_MESH || t == D3D_CONE;
    int k = 0;
    if (full_pose) {
#pragma unroll
        for (int j = 0; j < 12; ++j) T[j] = r[j];
        k = 12;
    }
    switch (t) {
    case D3D_SPHERE: T[3] = r[0]; T[7] = r[1]; T[11] = r[2]; p[0] = r[3]; break;
    case D3D_CAPSULE: case D3D_CYLINDER: case D3D_CONE: p[0] = r[k]; p[1] = r[k + 1]; break;
    case D3D_ELLIPSOID: p[0] = r[k]; p[1] = r[k + 1]; p[2] = r[k + 2]; break;
    case D3D_BOX:
        p[0] = r[k]; p[1] = r[k + 1]; p[2] = r[k + 2];
        vr = *reinterpret_cast<const int2 *>(r + k + 3);
        break;
    case D3D_HULL: vr = *reinterpret_cast<const int2 *>(r); break;
    case D3D_MESH: vr = *reinterpret_cast<const int2 *>(r + k); break;
    case D3D_DISK:  // pack.py: centre = pose[:3,3], normal = pose[:3,2] of an identity matrix
        T[3] = r[0]; T[7] = r[1]; T[11] = r[2]; T[2] = r[3]; T[6] = r[4]; T[10] = r[5]; p[0] = r[6];
        break;
    case D3D_ELLIPSE:  // centre, axes = pose[:3,0], pose[:3,1] of an identity matrix
        T[3] = r[0]; T[7] = r[1]; T[11] = r[2];
        T[0] = r[3]; T[4] = r[4]; T[8] = r[5]; T[1] = r[6]; T[5] = r[7]; T[9] = r[8];
        p[0] = r[9]; p[1] = r[10];
        break;
    }
    type[i] = t;
    double2 *o = reinterpret_cast<double2 *>(pose + 16 * i);
#pragma unroll
    for (int j = 0; j < 6; ++j) o[j] = make_double2(T[2 * j], T[2 * j + 1]);
    o[6] = make_double2(0.0, 0.0);
    o[7] = make_double2(0.0, 1.0);
    param[3 * i] = p[0]; param[3 * i + 1] = p[1]; param[3 * i + 2] = p[2];
    vert_off[i] = vr.x;
    vert_len[i] = vr.y;
}

__global__ void k_support(d3d_colliders c, const int32_t *idx, const double *dirs, int64_t n,
                          double *out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    Collider col = load_collider(c, idx[t]);
    st3(out + 3 * t, support<1>(col, ld3(dirs + 3 * t), 0));
    if (c.mesh_last && col.type == D3D_MESH) c.mesh_last[idx[t]] = col.cur;  // mesh.py:85
}

__global__ void k_center(d3d_colliders c, double *out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= c.n) return;
    Collider col = load_collider(c, t);
    st3(out + 3 * t, center_of(col));
}

// One thread per collider (d3d_aabb.cuh: collider_aabb).
__global__ void k_aabb(d3d_colliders c, double *out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= c.n) return;
    double lo[3], hi[3];
    collider_aabb(c, i, lo, hi);
    double2 *o = reinterpret_cast<double2 *>(out + 6 * i);
    o[0] = make_double2(lo[0], hi[0]);
    o[1] = make_double2(lo[1], hi[1]);
    o[2] = make_double2(lo[2], hi[2]);
}

// Dependent-free FP64 FMA loop: 8 independent accumulator chains per thread.
__global__ void k_fp64_peak(double *out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0,
           a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s;
}

__global__ void k_debug_vdiv(const double *v, const double *s, int64_t n, double *out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) st3(out + 3 * k, ld3(v + 3 * k) / s[k]);
}

__global__ void k_debug_norm(const double *v, int64_t n, double *out, int mode) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    double x = v[3 * t], y = v[3 * t + 1], z = v[3 * t + 2];
    if (mode == 0) { out[t] = norm_x87(x, y, z); return; }
    if (mode == 2) {  // the integer emulation seeded by the double-double estimate, for every vector
        double big = fmax(fabs(x), fmax(fabs(y), fabs(z)));
        out[t] = (big > 1e-140 && big < 1e140) ? norm_x87_core_t<true>(x, y, z) : norm_x87(x, y, z);
        return;
    }
    // exact emulation alone, seeded with a plain fp64 estimate
    double g = sqrt(x * x + y * y + z * z);
    out[t] = (g > 1e-140 && g < 1e140) ? norm_x87_exact(x, y, z, g, 0.0) : norm_x87(x, y, z);
}

}  // namespace

extern "C" {

const char *d3d_last_error_string(void) { return g_error; }

int d3d_prepare(const d3d_colliders *c, double *verts_out, void *stream) {
    if (!c || !verts_out) return d3d_set_error("d3d_prepare: null argument");
    if (c->n == 0) return 0;
    int64_t threads = c->n * 8;
    k_prepare<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*c, verts_out);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_unpack_colliders(const uint8_t *wire_type, const int32_t *wire_off, const double *wire,
                         int64_t n, int32_t *type, double *pose, double *param, int32_t *vert_off,
                         int32_t *vert_len, void *stream) {
    if (n == 0) return 0;
    if (!wire_type || !wire_off || !wire || !type || !pose || !param || !vert_off || !vert_len)
        return d3d_set_error("d3d_unpack_colliders: null argument");
    k_unpack<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(wire_type, wire_off, wire, n, type,
                                                                           pose, param, vert_off, vert_len);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_wire_offsets(const uint8_t *wire_type, int64_t n, int32_t *wire_off, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return 0;
    if (!wire_type || !wire_off) return d3d_set_error("d3d_wire_offsets: null argument");
    const int64_t n_tiles = (n + WIRE_TILE - 1) / WIRE_TILE;
    int32_t *tile_sum = nullptr;
    D3D_CUDA_CHECK(cudaMallocAsync(reinterpret_cast<void **>(&tile_sum), sizeof(int32_t) * (size_t)n_tiles, stream));
    k_wire_tile_sums<<<(unsigned)n_tiles, 256, 0, stream>>>(wire_type, n, tile_sum);
    k_wire_scan_sums<<<1, 1024, 0, stream>>>(tile_sum, n_tiles);
    k_wire_offsets<<<(unsigned)n_tiles, 256, 0, stream>>>(wire_type, n, tile_sum, wire_off);
    cudaError_t err = cudaGetLastError();
    cudaFreeAsync(tile_sum, stream);
    if (err != cudaSuccess) return d3d_set_error("d3d_wire_offsets: %s", cudaGetErrorString(err));
    return 0;
}

int d3d_support(const d3d_colliders *c, const int32_t *idx, const double *dirs, int64_t n,
                double *out, void *stream) {
    if (!c || !idx || !dirs || !out) return d3d_set_error("d3d_support: null argument");
    if (n == 0) return 0;
    k_support<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*c, idx, dirs, n, out);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_center(const d3d_colliders *c, double *out, void *stream) {
    if (!c || !out) return d3d_set_error("d3d_center: null argument");
    if (c->n == 0) return 0;
    k_center<<<(unsigned)((c->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*c, out);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_aabb(const d3d_colliders *c, double *out, void *stream) {
    if (!c || !out) return d3d_set_error("d3d_aabb: null argument");
    if (c->n == 0) return 0;
    k_aabb<<<(unsigned)((c->n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*c, out);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

/* test hook: out[k,:] = v[k,:] / s[k] through the vector division of d3d_math.cuh */
int d3d_debug_vdiv(const double *v, const double *s, int64_t n, double *out, void *stream) {
    if (n == 0) return 0;
    k_debug_vdiv<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(v, s, n, out);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

/* test hook: out[k] = norm of v[k,:] with the dnrm2 (x87) semantics of d3d_math.cuh;
 * mode 0 = production path, 1 = exact integer emulation only */
int d3d_debug_norm(const double *v, int64_t n, double *out, int mode, void *stream) {
    if (n == 0) return 0;
    k_debug_norm<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(v, n, out, mode);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

/* FP64 roofline probe: launches `blocks` x 256 threads doing 8*iters FMAs each
 * (2 flop per FMA); time it with CUDA events on `stream`. */
int d3d_fp64_peak_probe(double *scratch, int blocks, int iters, void *stream) {
    k_fp64_peak<<<blocks, 256, 0, (cudaStream_t)stream>>>(scratch, iters);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    D3D_CUDA_CHECK(cudaGetDevice(&dev));
    if (sm_count) *sm_count = d3d_sm_count();
    if (cc_major) D3D_CUDA_CHECK(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_minor) D3D_CUDA_CHECK(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
    return 0;
}

}  // extern "C"
