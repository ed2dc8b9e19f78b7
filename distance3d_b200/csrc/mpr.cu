// Minkowski portal refinement (XenoCollide), one thread per pair.
//
// Replaces distance3d/mpr.py:21-393 (mpr_intersection, mpr_penetration and their
// helpers), distance3d/minkowski.py:23-55 and distance3d/distance/_triangle.py:12-89
// (point_to_triangle with the point at the origin).  The reference's numba array
// aliasing in _swap_vertices (mpr.py:233-243: both rows end up equal to the old row
// idx2) is reproduced.  Pairs are processed in (typeA, typeB) order supplied by the
// caller (optional permutation), so warps evaluate one kind of support map.
#include "d3d_common.cuh"
#include "d3d_support.cuh"

namespace {

#define MPR_REFINE_CAP 4096
enum { ORIGIN_OUTSIDE = -1, PORTAL_BUILT = 0, ORIGIN_ON_V1 = 1, ORIGIN_ON_SEGMENT = 2 };

struct Portal {
    v3 v[4], v1[4], v2[4];
};

struct MprParams {
    double tol;
    int max_iterations;
    int want_pen;
    uint8_t *out_hit;
    double *out_depth;
    double *out_dir;
    double *out_pos;
    int32_t *out_status;
};

// minkowski.py:23-55
D3D_DEV void mink_support(const Collider &A, const Collider &B, v3 d, v3 &v, v3 &v1, v3 &v2) {
    v1 = support_ni<1>(A, d.x, d.y, d.z, 0);
    v2 = support_ni<1>(B, -d.x, -d.y, -d.z, 0);
    v = v1 - v2;
}

// mpr.py:273-279
D3D_DEV v3 portal_direction(const Portal &p) {
    return normalized(cross(p.v[2] - p.v[1], p.v[3] - p.v[1]));
}
// mpr.py:282-285
D3D_DEV bool encapsulates_origin(v3 v, v3 sd) { return dot_blas(v, sd) > -10.0 * D3D_EPS; }
// mpr.py:288-296
D3D_DEV bool reach_tolerance(const Portal &p, v3 v4, v3 sd, double tol) {
    double dv4 = dot_blas(v4, sd);
    double m = dv4 - gemv_row(p.v[1].x, p.v[1].y, p.v[1].z, sd);
    double m2 = dv4 - gemv_row(p.v[2].x, p.v[2].y, p.v[2].z, sd);
    double m3 = dv4 - gemv_row(p.v[3].x, p.v[3].y, p.v[3].z, sd);
    if (m2 < m) m = m2;
    if (m3 < m) m = m3;
    return m < tol + D3D_EPS;
}
// mpr.py:299-315
D3D_DEV void expand_portal(Portal &p, v3 v4, v3 v14, v3 v24) {
    v3 v4v0 = cross(v4, p.v[0]);
    int k;
    if (dot_blas(p.v[1], v4v0) > 0.0) k = (dot_blas(p.v[2], v4v0) > 0.0) ? 1 : 3;
    else k = (dot_blas(p.v[3], v4v0) > 0.0) ? 2 : 1;
    p.v[k] = v4; p.v1[k] = v14; p.v2[k] = v24;
}

// mpr.py:120-243
D3D_DEV int discover_portal(const Collider &A, const Collider &B, int max_iterations, Portal &p) {
    p.v1[0] = center_of(A);
    p.v2[0] = center_of(B);
    p.v[0] = p.v1[0] - p.v2[0];
    if (all_zero(p.v[0])) p.v[0].x += D3D_EPS * 10.0;
    v3 sd = normalized(-p.v[0]);
    mink_support(A, B, sd, p.v[1], p.v1[1], p.v2[1]);
    if (!all_zero(p.v[1]) && dot_blas(p.v[1], sd) < D3D_EPS) return ORIGIN_OUTSIDE;
    sd = cross(p.v[0], p.v[1]);
    if (dot_blas(sd, sd) < D3D_EPS) return all_zero(p.v[1]) ? ORIGIN_ON_V1 : ORIGIN_ON_SEGMENT;
    sd = normalized(sd);
    mink_support(A, B, sd, p.v[2], p.v1[2], p.v2[2]);
    if (dot_blas(p.v[2], sd) < D3D_EPS) return ORIGIN_OUTSIDE;
    sd = normalized(cross(p.v[1] - p.v[0], p.v[2] - p.v[0]));
    if (dot_blas(sd, p.v[0]) > 0.0) {
        p.v[1] = p.v[2]; p.v1[1] = p.v1[2]; p.v2[1] = p.v2[2];  // aliased "swap"
        sd = sd * -1.0;
    }
    int n_points = 3, it = 0;
    while (n_points < 4) {
        mink_support(A, B, sd, p.v[3], p.v1[3], p.v2[3]);
        if (dot_blas(p.v[3], sd) < D3D_EPS) return ORIGIN_OUTSIDE;
        bool cont = false;
        if (dot_blas(cross(p.v[1], p.v[3]), p.v[0]) < D3D_EPS) {
            p.v[2] = p.v[3]; p.v1[2] = p.v1[3]; p.v2[2] = p.v2[3];
            cont = true;
        }
        if (!cont && dot_blas(cross(p.v[3], p.v[2]), p.v[0]) < D3D_EPS) {
            p.v[1] = p.v[3]; p.v1[1] = p.v1[3]; p.v2[1] = p.v2[3];
            cont = true;
        }
        if (cont) sd = normalized(cross(p.v[1] - p.v[0], p.v[2] - p.v[0]));
        else n_points = 4;
        if (++it >= max_iterations) break;
    }
    return PORTAL_BUILT;
}

// mpr.py:246-270; -1 when the (reference-unbounded) loop hits the cap
D3D_DEV int refine_portal(const Collider &A, const Collider &B, Portal &p, double tol) {
    for (int it = 0; it < MPR_REFINE_CAP; ++it) {
        v3 sd = portal_direction(p);
        if (encapsulates_origin(p.v[1], sd)) return 1;
        v3 n, n1, n2;
        mink_support(A, B, sd, n, n1, n2);
        if (!encapsulates_origin(n, sd) || reach_tolerance(p, n, sd, tol)) return 0;
        expand_portal(p, n, n1, n2);
    }
    return -1;
}

// distance/_triangle.py:12-89 with point = 0
D3D_DEV double point_to_triangle_origin(v3 A, v3 B, v3 C, v3 &closest) {
    v3 zero = V3(0.0, 0.0, 0.0);
    v3 ab = B - A, ac = C - A;
    v3 ap = zero - A;
    double d1 = dot_blas(ab, ap), d2 = dot_blas(ac, ap);
    v3 bp = zero - B;
    double d3 = dot_blas(ab, bp), d4 = dot_blas(ac, bp);
    v3 cp = zero - C;
    double d5 = dot_blas(ab, cp), d6 = dot_blas(ac, cp);
    double vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
    v3 r;
    if (d1 <= 0.0 && d2 <= 0.0) r = A;
    else if (d3 >= 0.0 && d4 <= d3) r = B;
    else if (vc <= 0.0 && 0.0 <= d1 && d3 <= 0.0) r = A + ab * (d1 / (d1 - d3));
    else if (d6 >= 0.0 && d5 <= d6) r = C;
    else if (vb <= 0.0 && 0.0 <= d2 && d6 <= 0.0) r = A + ac * (d2 / (d2 - d6));
    else if (va <= 0.0 && 0.0 <= d4 - d3 && d5 - d6 >= 0.0)
        r = B + (C - B) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)));
    else {
        double denom = 1.0 / (va + vb + vc);
        double v = vb * denom, w = vc * denom;
        r = (A + ab * v) + ac * w;
    }
    closest = r;
    return norm3(zero - r);
}

// b.dot(M), b[4], M[4,3] inside numba (dgemv): fma(b0,M0,b1*M1) + fma(b2,M2,b3*M3)
D3D_DEV v3 vec4_mat(const double *b, const v3 *M) {
    return V3(fma(b[0], M[0].x, b[1] * M[1].x) + fma(b[2], M[2].x, b[3] * M[3].x),
              fma(b[0], M[0].y, b[1] * M[1].y) + fma(b[2], M[2].y, b[3] * M[3].y),
              fma(b[0], M[0].z, b[1] * M[1].z) + fma(b[2], M[2].z, b[3] * M[3].z));
}

// mpr.py:368-393
D3D_DEV v3 contact_position(const Portal &p, v3 sd) {
    double b[4];
    b[0] = dot_blas(cross(p.v[1], p.v[2]), p.v[3]);
    b[1] = dot_blas(cross(p.v[3], p.v[2]), p.v[0]);
    b[2] = dot_blas(cross(p.v[0], p.v[1]), p.v[3]);
    b[3] = dot_blas(cross(p.v[2], p.v[1]), p.v[0]);
    double sum = ((b[0] + b[1]) + b[2]) + b[3];
    if (sum < D3D_EPS) {
        b[0] = 0.0;
        b[1] = dot_blas(cross(p.v[2], p.v[3]), sd);
        b[2] = dot_blas(cross(p.v[3], p.v[1]), sd);
        b[3] = dot_blas(cross(p.v[1], p.v[2]), sd);
        sum = ((b[0] + b[1]) + b[2]) + b[3];
    }
    for (int i = 0; i < 4; ++i) b[i] /= sum;
    return (vec4_mat(b, p.v1) + vec4_mat(b, p.v2)) * 0.5;
}

__global__ void __launch_bounds__(128)
k_mpr(d3d_colliders c, const int32_t *__restrict__ pairs, const int32_t *__restrict__ perm,
      int64_t n_pairs, MprParams prm) {
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    int64_t k = perm ? perm[t] : t;
    int2 pr = __ldg(reinterpret_cast<const int2 *>(pairs) + k);
    Collider A = load_collider(c, pr.x), B = load_collider(c, pr.y);
    Portal p;
    for (int i = 0; i < 4; ++i) p.v[i] = p.v1[i] = p.v2[i] = V3(0.0, 0.0, 0.0);
    int res = discover_portal(A, B, prm.max_iterations, p);
    double depth = 0.0;
    v3 dir = V3(0.0, 0.0, 0.0), pos = dir;
    int hit, status = D3D_UNKNOWN;
    if (res == ORIGIN_OUTSIDE) {
        hit = 0;
    } else if (res == ORIGIN_ON_V1) {  // mpr.py:347-353
        hit = 1;
        pos = (p.v1[1] + p.v2[1]) * 0.5;
    } else if (res == ORIGIN_ON_SEGMENT) {  // mpr.py:356-365
        hit = 1;
        pos = (p.v1[1] + p.v2[1]) * 0.5;
        depth = norm3(p.v[1]);
        dir = normalized(p.v[1]);
    } else {
        hit = refine_portal(A, B, p, prm.tol);
        if (hit < 0) { hit = 0; status = D3D_ITER_CAP; }
        if (hit && prm.want_pen) {  // mpr.py:318-344
            int iterations = 0;
            for (;;) {
                v3 sd = portal_direction(p);
                v3 n, n1, n2;
                mink_support(A, B, sd, n, n1, n2);
                if (reach_tolerance(p, n, sd, prm.tol) || iterations > prm.max_iterations) {
                    v3 cp;
                    depth = point_to_triangle_origin(p.v[1], p.v[2], p.v[3], cp);
                    if (fabs(depth) < D3D_EPS) cp = V3(0.0, 0.0, 0.0);
                    pos = contact_position(p, portal_direction(p));
                    dir = normalized(cp);
                    break;
                }
                expand_portal(p, n, n1, n2);
                ++iterations;
            }
        }
    }
    if (status == D3D_UNKNOWN) status = hit ? D3D_INTERSECTION : D3D_NO_INTERSECTION;
    prm.out_hit[k] = (uint8_t)hit;
    if (c.mesh_last) {  // mesh.py:85
        if (A.type == D3D_MESH) c.mesh_last[pr.x] = A.cur;
        if (B.type == D3D_MESH) c.mesh_last[pr.y] = B.cur;
    }
    if (prm.out_status) prm.out_status[k] = status;
    if (prm.want_pen) {
        prm.out_depth[k] = depth;
        st3(prm.out_dir + 3 * k, dir);
        st3(prm.out_pos + 3 * k, pos);
    }
}

}  // namespace

extern "C" {

int d3d_mpr(const d3d_colliders *c, const int32_t *pairs, const int32_t *perm, int64_t n_pairs,
            double tolerance, int max_iterations, int want_penetration, uint8_t *out_hit,
            double *out_depth, double *out_dir, double *out_pos, int32_t *out_status,
            void *stream) {
    if (n_pairs == 0) return 0;
    if (!c || !pairs || !out_hit) return d3d_set_error("d3d_mpr: null argument");
    if (want_penetration && (!out_depth || !out_dir || !out_pos))
        return d3d_set_error("d3d_mpr: penetration outputs missing");
    MprParams prm;
    prm.tol = tolerance; prm.max_iterations = max_iterations; prm.want_pen = want_penetration;
    prm.out_hit = out_hit; prm.out_depth = out_depth; prm.out_dir = out_dir; prm.out_pos = out_pos;
    prm.out_status = out_status;
    k_mpr<<<(unsigned)((n_pairs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*c, pairs, perm, n_pairs, prm);
    D3D_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // extern "C"
