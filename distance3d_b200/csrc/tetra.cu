// Tetrahedron-pair intersection of the hydroelastic contact model: the other consumer of the
// AABB broad phase in the reference.
//
// Replaces, for MANY candidate pairs per call,
//   hydroelastic_contact/_mesh_processing.py:4-20     tetrahedral_mesh_aabbs
//   hydroelastic_contact/_barycentric_transform.py:4-9 barycentric_transforms (numpy pinv)
//   hydroelastic_contact/_tetrahedron_intersection.py:7-423 intersect_tetrahedron_pairs ->
//       contact_plane, check_tetrahedra_intersect_contact_plane, compute_contact_polygon
//       (make_halfplanes, order_points, filter_unique_points, project_polygon_to_3d)
//   hydroelastic_contact/_halfplanes.py:9-71           intersect_halfplanes
// which the reference runs one pair at a time from a Python loop (:64-80).  One thread per
// candidate pair; everything (two 4x4 barycentric transforms, eight half-planes, up to 24
// candidate vertices) lives in registers / local memory of the thread.  The candidate pairs
// come from d3d_bvh_overlap over the tetrahedron boxes (hydroelastic_contact.py).
//
// Parity with the reference is by tolerance (planes and polygon REGIONS within 1e-9): the
// reference inverts with LAPACK's pinv and sorts by libm's atan2, which are not reproduced bit
// for bit.  Compiled with -fmad=false like the rest of the library; same arithmetic as the C
// oracle (oracle/src/tetra.c).
#include <math.h>

#include "d3d_common.cuh"

namespace {

#define TETRA_EPS 2.220446049250313e-16

struct d3 {
    double x, y, z;
};
__device__ __forceinline__ d3 D3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return D3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return D3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator*(d3 a, double s) { return D3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dotp(d3 a, d3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ d3 crossp(d3 a, d3 b) {
    return D3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ d3 ldp(const double *p) { return D3(__ldg(p), __ldg(p + 1), __ldg(p + 2)); }
__device__ __forceinline__ double cross2d(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }

// _mesh_processing.py:4-20
__global__ void k_tetra_aabb(const double *__restrict__ points, int64_t n, double *out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const double *p = points + 12 * t;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double lo = __ldg(p + k), hi = lo;
#pragma unroll
        for (int v = 1; v < 4; ++v) {
            double x = __ldg(p + 3 * v + k);
            lo = fmin(lo, x);
            hi = fmax(hi, x);
        }
        out[6 * t + 2 * k] = lo;
        out[6 * t + 2 * k + 1] = hi;
    }
}

// _barycentric_transform.py:4-9: X = inverse of [[p0 p1 p2 p3], [1 1 1 1]].  Row i is the
// barycentric coordinate function of vertex i: the plane through the other three vertices,
// scaled to 1 at vertex i (closed form instead of the reference's pinv).
__device__ void barycentric(const double *p, double *X) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        d3 pi = ldp(p + 3 * i);
        d3 a = ldp(p + 3 * ((i + 1) & 3)), b = ldp(p + 3 * ((i + 2) & 3)), c = ldp(p + 3 * ((i + 3) & 3));
        d3 n = crossp(b - a, c - a);
        double w = dotp(n, pi - a);
        X[4 * i] = n.x / w;
        X[4 * i + 1] = n.y / w;
        X[4 * i + 2] = n.z / w;
        X[4 * i + 3] = -dotp(n, a) / w;
    }
}

__global__ void k_barycentric(const double *__restrict__ points, int64_t n, double *out) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    double X[16];
    barycentric(points + 12 * t, X);
#pragma unroll
    for (int k = 0; k < 16; ++k) out[16 * t + k] = X[k];
}

// utils.py:78-122 plane_basis_from_normal
__device__ void plane_basis(d3 n, d3 &x, d3 &y) {
    if (fabs(n.x) >= fabs(n.y)) {
        double len = sqrt(n.x * n.x + n.z * n.z);
        x = D3(-n.z / len, 0.0, n.x / len);
        y = D3(n.y * x.z, n.z * x.x - n.x * x.z, -n.y * x.x);
    } else {
        double len = sqrt(n.y * n.y + n.z * n.z);
        x = D3(0.0, n.z / len, -n.y / len);
        y = D3(n.y * x.z - n.z * x.y, -n.x * x.z, n.x * x.y);
    }
}

struct TetraParams {
    const double *points1, *eps1, *X1, *points2, *eps2, *X2;
    double ym1, ym2;
    int max_vertices;
    uint8_t *out_hit;
    double *out_plane;
    int32_t *out_nverts;
    double *out_poly;
    int32_t *out_status;
};

// _tetrahedron_intersection.py:87-140 intersect_tetrahedron_pair
__global__ void __launch_bounds__(128)
k_tetra_pairs(const int32_t *__restrict__ pairs, int64_t n_pairs, TetraParams prm) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    int2 pr = __ldg(reinterpret_cast<const int2 *>(pairs) + k);
    const double *t1 = prm.points1 + 12 * (int64_t)pr.x, *t2 = prm.points2 + 12 * (int64_t)pr.y;
    const double *e1 = prm.eps1 + 4 * (int64_t)pr.x, *e2 = prm.eps2 + 4 * (int64_t)pr.y;
    double X[32];  // rows 0-3: tetrahedron 1, rows 4-7: tetrahedron 2
    if (prm.X1) {
        for (int q = 0; q < 16; ++q) X[q] = __ldg(prm.X1 + 16 * (int64_t)pr.x + q);
    } else {
        barycentric(t1, X);
    }
    if (prm.X2) {
        for (int q = 0; q < 16; ++q) X[16 + q] = __ldg(prm.X2 + 16 * (int64_t)pr.y + q);
    } else {
        barycentric(t2, X + 16);
    }
    double *poly = prm.out_poly + (size_t)k * prm.max_vertices * 3;
    double plane[4];
    int n_vertices = 0, status = 0, hit = 0;
    // contact_plane :165-216
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            s1 += (__ldg(e1 + r) * prm.ym1) * X[4 * r + c];
            s2 += (__ldg(e2 + r) * prm.ym2) * X[16 + 4 * r + c];
        }
        plane[c] = s1 - s2;
    }
    double norm = sqrt(plane[0] * plane[0] + plane[1] * plane[1] + plane[2] * plane[2]);
    bool same = false;
    if (norm == 0.0) {
        same = true;
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) plane[c] /= norm;
        plane[3] *= -1.0;
        if (fabs(plane[3]) < 10.0 * TETRA_EPS) same = true;  // sic: a plane through the origin counts as "same"
    }
    if (same) {  // _handle_same_tetrahedron :143-162
        double sum = ((__ldg(e2) + __ldg(e2 + 1)) + __ldg(e2 + 2)) + __ldg(e2 + 3);
        d3 pp = D3(0.0, 0.0, 0.0);
        for (int r = 0; r < 4; ++r) pp = pp + ldp(t2 + 3 * r) * (__ldg(e2 + r) / sum);
        double d = sqrt(pp.x * pp.x + pp.y * pp.y + pp.z * pp.z);
        d3 n = d > 0.0 ? D3(pp.x / d, pp.y / d, pp.z / d) : D3(0.0, 0.0, 1.0);
        plane[0] = n.x; plane[1] = n.y; plane[2] = n.z; plane[3] = d;
        status = 1;
        hit = 1;
        n_vertices = min(3, prm.max_vertices);
        for (int q = 0; q < n_vertices; ++q) { poly[3 * q] = pp.x; poly[3 * q + 1] = pp.y; poly[3 * q + 2] = pp.z; }
    } else {
        d3 n = D3(plane[0], plane[1], plane[2]);
        double d = plane[3];
        // check_tetrahedra_intersect_contact_plane :219-252
        const double tol = 1e-6;
        double lo1 = 0, hi1 = 0, lo2 = 0, hi2 = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double a = dotp(ldp(t1 + 3 * r), n) - d, b = dotp(ldp(t2 + 3 * r), n) - d;
            if (r == 0 || a < lo1) lo1 = a;
            if (r == 0 || a > hi1) hi1 = a;
            if (r == 0 || b < lo2) lo2 = b;
            if (r == 0 || b > hi2) hi2 = b;
        }
        if (lo1 < -tol && hi1 > tol && lo2 < -tol && hi2 > tol) {
            // compute_contact_polygon :377-423
            d3 pp = n * d, bx, by;
            plane_basis(n, bx, by);
            double hp[8][4];  // make_halfplanes :255-292
            int nh = 0;
            for (int i = 0; i < 8; ++i) {
                d3 fn = D3(X[4 * i], X[4 * i + 1], X[4 * i + 2]);
                double nx = dotp(fn, bx), ny = dotp(fn, by);
                double ds = -X[4 * i + 3] - dotp(fn, pp);
                double nn = sqrt(nx * nx + ny * ny);
                if (nn > TETRA_EPS) {
                    hp[nh][0] = nx * ds / (nn * nn);
                    hp[nh][1] = ny * ds / (nn * nn);
                    hp[nh][2] = ny;
                    hp[nh][3] = -nx;
                    ++nh;
                }
            }
            double pts[24][2];  // intersect_halfplanes (_halfplanes.py:35-71)
            int np_ = 0;
            for (int i = 0; i < nh; ++i)
                for (int j = i + 1; j < nh; ++j) {
                    double denom = cross2d(hp[i][2], hp[i][3], hp[j][2], hp[j][3]);
                    if (fabs(denom) < TETRA_EPS) continue;
                    double t = cross2d(hp[j][0] - hp[i][0], hp[j][1] - hp[i][1], hp[j][2], hp[j][3]) / denom;
                    double px = hp[i][0] + hp[i][2] * t, py = hp[i][1] + hp[i][3] * t;
                    bool valid = true;
                    for (int q = 0; q < nh && valid; ++q)
                        if (q != i && q != j && cross2d(hp[q][2], hp[q][3], px - hp[q][0], py - hp[q][1]) < -TETRA_EPS)
                            valid = false;
                    if (valid && np_ < 24) { pts[np_][0] = px; pts[np_][1] = py; ++np_; }
                }
            if (np_ >= 3) {
                // order_points :295-313
                double cx = 0.0, cy = 0.0;
                for (int q = 0; q < np_; ++q) { cx += pts[q][0]; cy += pts[q][1]; }
                cx /= np_; cy /= np_;
                double ang[24];
                int ord[24];
                for (int q = 0; q < np_; ++q) { ang[q] = atan2(pts[q][1] - cy, pts[q][0] - cx); ord[q] = q; }
                for (int a = 1; a < np_; ++a) {
                    int o = ord[a], b = a - 1;
                    while (b >= 0 && ang[ord[b]] > ang[o]) { ord[b + 1] = ord[b]; --b; }
                    ord[b + 1] = o;
                }
                // filter_unique_points :316-343 + project_polygon_to_3d :346-374
                int nu = 0;
                for (int q = 0; q < np_; ++q) {
                    double x = pts[ord[q]][0], y = pts[ord[q]][1];
                    if (q > 0) {
                        double dx = x - pts[ord[q - 1]][0], dy = y - pts[ord[q - 1]][1];
                        if (!(sqrt(dx * dx + dy * dy) > 10.0 * TETRA_EPS)) continue;
                    }
                    if (nu < prm.max_vertices) {
                        d3 v = (bx * x + by * y) + pp;
                        poly[3 * nu] = v.x; poly[3 * nu + 1] = v.y; poly[3 * nu + 2] = v.z;
                    }
                    ++nu;
                }
                if (nu >= 3) {
                    hit = 1;
                    if (nu > prm.max_vertices) { status = 2; nu = prm.max_vertices; }
                    n_vertices = nu;
                }
            }
        }
    }
    prm.out_hit[k] = (uint8_t)hit;
#pragma unroll
    for (int c = 0; c < 4; ++c) prm.out_plane[4 * k + c] = plane[c];
    prm.out_nverts[k] = n_vertices;
    if (prm.out_status) prm.out_status[k] = status;
}

}  // namespace

extern "C" {

int d3d_tetra_aabb(const double *points, int64_t n, double *out_aabb, void *stream) {
    if (n == 0) return 0;
    if (!points || !out_aabb) return d3d_set_error("d3d_tetra_aabb: null argument");
    k_tetra_aabb<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(points, n, out_aabb);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_tetra_barycentric(const double *points, int64_t n, double *out_X, void *stream) {
    if (n == 0) return 0;
    if (!points || !out_X) return d3d_set_error("d3d_tetra_barycentric: null argument");
    k_barycentric<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(points, n, out_X);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int d3d_tetra_intersect_pairs(const int32_t *pairs, int64_t n_pairs, const double *points1,
                              const double *eps1, const double *X1, const double *points2,
                              const double *eps2, const double *X2, double youngs_modulus1,
                              double youngs_modulus2, int max_vertices, uint8_t *out_hit,
                              double *out_plane, int32_t *out_nverts, double *out_poly,
                              int32_t *out_status, void *stream) {
    if (n_pairs == 0) return 0;
    if (!pairs || !points1 || !eps1 || !points2 || !eps2 || !out_hit || !out_plane || !out_nverts || !out_poly)
        return d3d_set_error("d3d_tetra_intersect_pairs: null argument");
    if (max_vertices < 3 || max_vertices > 24)
        return d3d_set_error("d3d_tetra_intersect_pairs: max_vertices must be in [3, 24]");
    TetraParams prm;
    prm.points1 = points1; prm.eps1 = eps1; prm.X1 = X1; prm.points2 = points2; prm.eps2 = eps2; prm.X2 = X2;
    prm.ym1 = youngs_modulus1; prm.ym2 = youngs_modulus2; prm.max_vertices = max_vertices;
    prm.out_hit = out_hit; prm.out_plane = out_plane; prm.out_nverts = out_nverts; prm.out_poly = out_poly;
    prm.out_status = out_status;
    k_tetra_pairs<<<(unsigned)((n_pairs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(pairs, n_pairs, prm);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
