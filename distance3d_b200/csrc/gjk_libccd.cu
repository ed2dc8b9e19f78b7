// libccd-style boolean GJK, one thread per pair: the reference's second intersection test
// (SURVEY.md section 8f #4).  The reference uses it as an independent cross-check of the Jolt
// variant (distance3d/test/test_gjk.py:341-354); batched here it plays the same role for
// d3d_gjk_intersection on the device.
//
// Replaces distance3d/gjk/_gjk_libccd.py:14-266 (gjk_intersection_libccd, _gjk, _line_segment,
// _triangle, _triangle_ab, _tetrahedron, _rearrange_simplex_to_triangle, _triple_cross),
// distance3d/distance/_triangle.py:12-89 (point_to_triangle, distance only) and the
// first_vertex() methods of distance3d/colliders.py.  Only the Minkowski-difference points
// decide the boolean, so v1 / v2 of the reference's Simplex are not carried.
#include "d3d_common.cuh"
#include "d3d_support.cuh"

namespace {

#define CCD_EPS_SQRT 1.4901161193847656e-08  // math.sqrt(EPSILON), _gjk_libccd.py:11
enum { CCD_NO_CONTACT = -1, CCD_CONTINUE = 0, CCD_CONTACT = 1 };

// colliders.py:128,217,269,322,370,422,475,530,586,626
D3D_DEV v3 first_vertex(const Collider &c) {
    v3 t = V3(c.tx(), c.ty(), c.tz()), z = V3(c.r02(), c.r12(), c.r22());
    switch (c.type) {
    case D3D_SPHERE: return t + V3(0.0, 0.0, c.p0());
    case D3D_CAPSULE: return t - z * (c.p0() + 0.5 * c.p1());
    case D3D_ELLIPSOID: return t + z * c.p2();
    case D3D_CYLINDER: return t + z * (0.5 * c.p1());
    case D3D_CONE: return t + z * c.p1();
    case D3D_BOX:
    case D3D_HULL: return ld3(c.V);
    case D3D_MESH: return xform(c, ld3(c.V));
    case D3D_DISK: {
        v3 x, y;
        plane_basis(z, x, y);
        return t + x * c.p0();
    }
    case D3D_ELLIPSE: return t + V3(c.r00(), c.r10(), c.r20()) * c.p0();
    }
    return t;
}

// distance/_triangle.py:12-89
static __device__ __noinline__ double point_to_triangle(v3 P, v3 A, v3 B, v3 C) {
    v3 ab = B - A, ac = C - A;
    v3 ap = P - A;
    double d1 = dot_blas(ab, ap), d2 = dot_blas(ac, ap);
    v3 bp = P - B;
    double d3 = dot_blas(ab, bp), d4 = dot_blas(ac, bp);
    v3 cp = P - C;
    double d5 = dot_blas(ab, cp), d6 = dot_blas(ac, cp);
    v3 q;
    if (d1 <= 0.0 && d2 <= 0.0) q = A;
    else if (d3 >= 0.0 && d4 <= d3) q = B;
    else {
        double vc = d1 * d4 - d3 * d2;
        if (vc <= 0.0 && 0.0 <= d1 && d3 <= 0.0) q = A + ab * (d1 / (d1 - d3));
        else if (d6 >= 0.0 && d5 <= d6) q = C;
        else {
            double vb = d5 * d2 - d1 * d6;
            if (vb <= 0.0 && 0.0 <= d2 && d6 <= 0.0) q = A + ac * (d2 / (d2 - d6));
            else {
                double va = d3 * d6 - d5 * d4;
                if (va <= 0.0 && 0.0 <= d4 - d3 && d5 - d6 >= 0.0)
                    q = B + (C - B) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)));
                else {
                    double denom = 1.0 / (va + vb + vc);
                    q = (A + ab * (vb * denom)) + ac * (vc * denom);
                }
            }
        }
    }
    return norm3(P - q);
}

D3D_DEV v3 triple_cross(v3 a, v3 b, v3 c) { return cross(cross(a, b), c); }
D3D_DEV bool all_close(v3 a, v3 b) {
    return fabs(a.x - b.x) < D3D_EPS && fabs(a.y - b.y) < D3D_EPS && fabs(a.z - b.z) < D3D_EPS;
}
D3D_DEV int sgn(double x) { return (x > 0.0) - (x < 0.0); }

// _gjk_libccd.py:112-134
D3D_DEV int line_segment(v3 *v, v3 &sd, int &n) {
    v3 A = v[1], B = v[0];
    v3 AB = B - A, AO = -A;
    double on_ab = dot_blas(AB, AO);
    v3 tmp = cross(AB, AO);
    if (fabs(dot_blas(tmp, tmp)) < D3D_EPS && on_ab > 0.0) { n = 2; return CCD_CONTACT; }
    if (on_ab < D3D_EPS) { v[0] = A; n = 1; sd = AO; }
    else { sd = triple_cross(AB, AO, AB); n = 2; }
    return CCD_CONTINUE;
}

// :177-187
D3D_DEV void triangle_ab(v3 A, v3 B, v3 AB, v3 AO, v3 *v, v3 &sd, int &n) {
    if (dot_blas(AB, AO) > -D3D_EPS) { v[0] = B; v[1] = A; n = 2; sd = triple_cross(AB, AO, AB); }
    else { v[0] = A; n = 1; sd = AO; }
}

// :137-174
static __device__ __noinline__ int triangle(v3 *v, v3 &sd, int &n) {
    v3 A = v[2], B = v[1], C = v[0];
    if (fabs(point_to_triangle(V3(0.0, 0.0, 0.0), A, B, C)) < CCD_EPS_SQRT) { n = 1; return CCD_CONTACT; }
    if (all_close(A, B) || all_close(A, C)) { n = 0; return CCD_NO_CONTACT; }
    v3 AO = -A, AB = B - A, AC = C - A;
    v3 ABC = cross(AB, AC);
    if (dot_blas(cross(ABC, AC), AO) > -D3D_EPS) {
        if (dot_blas(AC, AO) > -D3D_EPS) { v[1] = A; n = 2; sd = triple_cross(AC, AO, AC); }
        else triangle_ab(A, B, AB, AO, v, sd, n);
    } else {
        if (dot_blas(cross(AB, ABC), AO) > -D3D_EPS) triangle_ab(A, B, AB, AO, v, sd, n);
        else if (dot_blas(ABC, AO) > -D3D_EPS) { n = 3; sd = ABC; }
        else { v[0] = B; v[1] = C; n = 3; sd = -ABC; }
    }
    return CCD_CONTINUE;
}

// :190-260
D3D_DEV int tetrahedron(v3 *v, v3 &sd, int &n) {
    v3 A = v[3], B = v[2], C = v[1], D = v[0];
    if (fabs(point_to_triangle(A, B, C, D)) < CCD_EPS_SQRT) { n = 0; return CCD_NO_CONTACT; }
    v3 O = V3(0.0, 0.0, 0.0);
    if (point_to_triangle(O, A, B, C) < CCD_EPS_SQRT || point_to_triangle(O, A, C, D) < CCD_EPS_SQRT ||
        point_to_triangle(O, A, B, D) < CCD_EPS_SQRT || point_to_triangle(O, B, C, D) < CCD_EPS_SQRT) {
        n = 3;
        return CCD_CONTACT;
    }
    v3 AO = -A, AB = B - A, AC = C - A, AD = D - A;
    v3 ABC = cross(AB, AC), ACD = cross(AC, AD), ADB = cross(AD, AB);
    int b_on_acd = sgn(dot_blas(ACD, AB)), c_on_adb = sgn(dot_blas(ADB, AC)), d_on_abc = sgn(dot_blas(ABC, AD));
    bool ab_o = sgn(dot_blas(ACD, AO)) == b_on_acd, ac_o = sgn(dot_blas(ADB, AO)) == c_on_adb,
         ad_o = sgn(dot_blas(ABC, AO)) == d_on_abc;
    if (ab_o && ac_o && ad_o) { n = 4; return CCD_CONTACT; }
    if (!ab_o) { v[2] = A; }                                 // :249-260
    else if (!ac_o) { v[1] = D; v[0] = B; v[2] = A; }
    else { v[0] = C; v[1] = B; v[2] = A; }
    return triangle(v, sd, n);
}

__global__ void __launch_bounds__(128)
k_gjk_libccd(d3d_colliders c, const int32_t *__restrict__ pairs, const int32_t *__restrict__ perm,
             int64_t n_pairs, int max_iterations, uint8_t *out_hit, int32_t *out_iters) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    int64_t k = perm ? perm[t] : t;
    int2 pr = __ldg(reinterpret_cast<const int2 *>(pairs) + k);
    Collider A = load_collider(c, pr.x), B = load_collider(c, pr.y);
    v3 v[4];
    v[0] = first_vertex(A) - first_vertex(B);
    v[1] = v[2] = v[3] = V3(0.0, 0.0, 0.0);
    int n = 1, hit = 0, it = 0;
    v3 sd = -v[0];
    for (it = 0; it < max_iterations; ++it) {  // _gjk :56-91
        v3 sp = support_ni<1>(A, sd.x, sd.y, sd.z, 0) - support_ni<1>(B, -sd.x, -sd.y, -sd.z, 0);
        if (dot_blas(sp, sp) < D3D_EPS) { hit = 1; ++it; break; }
        if (dot_blas(sp, sd) < -CCD_EPS_SQRT) { ++it; break; }
        v[n++] = sp;
        int state = n == 2 ? line_segment(v, sd, n) : (n == 3 ? triangle(v, sd, n) : tetrahedron(v, sd, n));
        if (state == CCD_CONTACT) { hit = 1; ++it; break; }
        if (state == CCD_NO_CONTACT) { ++it; break; }
        if (fabs(dot_blas(sd, sd)) < D3D_EPS) { ++it; break; }
    }
    out_hit[k] = (uint8_t)hit;
    if (out_iters) out_iters[k] = it;
    if (c.mesh_last) {  // mesh.py:85
        if (A.type == D3D_MESH) c.mesh_last[pr.x] = A.cur;
        if (B.type == D3D_MESH) c.mesh_last[pr.y] = B.cur;
    }
}

}  // namespace

extern "C" {

int d3d_gjk_intersection_libccd(const d3d_colliders *c, const int32_t *pairs, const int32_t *perm,
                                int64_t n_pairs, int max_iterations, uint8_t *out_hit,
                                int32_t *out_iters, void *stream) {
    if (n_pairs == 0) return 0;
    if (!c || !pairs || !out_hit) return d3d_set_error("d3d_gjk_intersection_libccd: null argument");
    k_gjk_libccd<<<(unsigned)((n_pairs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        *c, pairs, perm, n_pairs, max_iterations, out_hit, out_iters);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
