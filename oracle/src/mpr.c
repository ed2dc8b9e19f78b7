/* TEST INFRASTRUCTURE (CPU oracle): Minkowski portal refinement.
 * Follows distance3d/mpr.py:21-393, distance3d/minkowski.py:23-55 and
 * distance3d/distance/_triangle.py:12-89, including the aliasing behaviour of
 * _swap_vertices (mpr.py:233-243: both rows end up equal to the old row idx2). */
#include "d3d_oracle.h"
#include "vec.h"

v3 d3do_support_s(const d3d_colliders *c, int64_t i, v3 d, int32_t *cur);
void d3do_pair_begin(const d3d_colliders *c, int64_t ia, int64_t ib, int32_t *cur);
void d3do_pair_end(const d3d_colliders *c, int64_t ia, int64_t ib, const int32_t *cur);
/* MeshGraph vertex of collider A / B of the pair this thread is working on */
static _Thread_local int32_t mesh_cur[2];
v3 d3do_center_v(const d3d_colliders *c, int64_t i);

#define EPS D3D_EPS
#define MPR_REFINE_CAP 4096

typedef struct { v3 v[4], v1[4], v2[4]; } portal_t;
enum { ORIGIN_OUTSIDE = -1, PORTAL_BUILT = 0, ORIGIN_ON_V1 = 1, ORIGIN_ON_SEGMENT = 2 };

/* minkowski.py:23-55 */
static void mink_support(const d3d_colliders *c, int64_t ia, int64_t ib, v3 d, v3 *v, v3 *v1,
                         v3 *v2) {
    *v1 = d3do_support_s(c, ia, d, &mesh_cur[0]);
    *v2 = d3do_support_s(c, ib, vneg(d), &mesh_cur[1]);
    *v = vsub(*v1, *v2);
}

static int all_zero(v3 a) { return a.x == 0.0 && a.y == 0.0 && a.z == 0.0; }

/* mpr.py:120-158 */
static int discover_portal(const d3d_colliders *c, int64_t ia, int64_t ib, int max_iterations,
                           portal_t *p) {
    /* _find_origin_ray :161-175 */
    p->v1[0] = d3do_center_v(c, ia);
    p->v2[0] = d3do_center_v(c, ib);
    p->v[0] = vsub(p->v1[0], p->v2[0]);
    if (all_zero(p->v[0])) p->v[0].x += EPS * 10.0;
    /* _find_support_in_direction_of_origin_ray :178-186 */
    v3 sd = vnormalized(vneg(p->v[0]));
    mink_support(c, ia, ib, sd, &p->v[1], &p->v1[1], &p->v2[1]);
    if (!all_zero(p->v[1]) && vdot(p->v[1], sd) < EPS) return ORIGIN_OUTSIDE;
    /* _find_support_perpendicular_to_plane_containing_origin_v01 :189-202 */
    sd = vcross(p->v[0], p->v[1]);
    if (vdot(sd, sd) < EPS) return all_zero(p->v[1]) ? ORIGIN_ON_V1 : ORIGIN_ON_SEGMENT;
    sd = vnormalized(sd);
    mink_support(c, ia, ib, sd, &p->v[2], &p->v1[2], &p->v2[2]);
    if (vdot(p->v[2], sd) < EPS) return ORIGIN_OUTSIDE;
    /* _search_direction_perpendicular_to_plane_containing_v012 :205-211 */
    sd = vnormalized(vcross(vsub(p->v[1], p->v[0]), vsub(p->v[2], p->v[0])));
    if (vdot(sd, p->v[0]) > 0.0) {
        p->v[1] = p->v[2]; p->v1[1] = p->v1[2]; p->v2[1] = p->v2[2]; /* aliased "swap" */
        sd = vscale(sd, -1.0);
    }
    int n_points = 3, it = 0;
    while (n_points < 4) {
        mink_support(c, ia, ib, sd, &p->v[3], &p->v1[3], &p->v2[3]);
        if (vdot(p->v[3], sd) < EPS) return ORIGIN_OUTSIDE;
        /* _iterate_discover_portal :214-230 */
        int cont = 0;
        if (vdot(vcross(p->v[1], p->v[3]), p->v[0]) < EPS) {
            p->v[2] = p->v[3]; p->v1[2] = p->v1[3]; p->v2[2] = p->v2[3];
            cont = 1;
        }
        if (!cont && vdot(vcross(p->v[3], p->v[2]), p->v[0]) < EPS) {
            p->v[1] = p->v[3]; p->v1[1] = p->v1[3]; p->v2[1] = p->v2[3];
            cont = 1;
        }
        if (cont) sd = vnormalized(vcross(vsub(p->v[1], p->v[0]), vsub(p->v[2], p->v[0])));
        else n_points = 4;
        if (++it >= max_iterations) break;
    }
    return PORTAL_BUILT;
}

/* mpr.py:273-279 */
static v3 portal_direction(const portal_t *p) {
    return vnormalized(vcross(vsub(p->v[2], p->v[1]), vsub(p->v[3], p->v[1])));
}
/* mpr.py:282-285 */
static int encapsulates_origin(v3 v, v3 sd) { return vdot(v, sd) > -10.0 * EPS; }
/* mpr.py:288-296: min(v4.dot(sd) - v[1:].dot(sd)) < tol + EPS (ddot minus dgemv rows) */
static int reach_tolerance(const portal_t *p, v3 v4, v3 sd, double tol) {
    double dv4 = vdot(v4, sd);
    double m = dv4 - gemv_row(p->v[1].x, p->v[1].y, p->v[1].z, sd);
    double m2 = dv4 - gemv_row(p->v[2].x, p->v[2].y, p->v[2].z, sd);
    double m3 = dv4 - gemv_row(p->v[3].x, p->v[3].y, p->v[3].z, sd);
    if (m2 < m) m = m2;
    if (m3 < m) m = m3;
    return m < tol + EPS;
}
/* mpr.py:299-315 */
static void expand_portal(portal_t *p, v3 v4, v3 v14, v3 v24) {
    v3 v4v0 = vcross(v4, p->v[0]);
    int k;
    if (vdot(p->v[1], v4v0) > 0.0) k = (vdot(p->v[2], v4v0) > 0.0) ? 1 : 3;
    else k = (vdot(p->v[3], v4v0) > 0.0) ? 2 : 1;
    p->v[k] = v4; p->v1[k] = v14; p->v2[k] = v24;
}

/* mpr.py:246-270; returns 1/0, or -1 when the (reference-unbounded) loop hits our cap */
static int refine_portal(const d3d_colliders *c, int64_t ia, int64_t ib, portal_t *p, double tol) {
    for (int it = 0; it < MPR_REFINE_CAP; ++it) {
        v3 sd = portal_direction(p);
        if (encapsulates_origin(p->v[1], sd)) return 1;
        v3 n, n1, n2;
        mink_support(c, ia, ib, sd, &n, &n1, &n2);
        if (!encapsulates_origin(n, sd) || reach_tolerance(p, n, sd, tol)) return 0;
        expand_portal(p, n, n1, n2);
    }
    return -1;
}

/* distance/_triangle.py:12-89 with point = 0 */
static double point_to_triangle_origin(v3 A, v3 B, v3 C, v3 *closest) {
    v3 zero = V3(0.0, 0.0, 0.0);
    v3 ab = vsub(B, A), ac = vsub(C, A);
    v3 ap = vsub(zero, A);
    double d1 = vdot(ab, ap), d2 = vdot(ac, ap);
    v3 cp_;
    if (d1 <= 0.0 && d2 <= 0.0) { cp_ = A; goto done; }
    v3 bp = vsub(zero, B);
    double d3 = vdot(ab, bp), d4 = vdot(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) { cp_ = B; goto done; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && 0.0 <= d1 && d3 <= 0.0) {
        double v = d1 / (d1 - d3);
        cp_ = vadd(A, vscale(ab, v));
        goto done;
    }
    v3 cp = vsub(zero, C);
    double d5 = vdot(ab, cp), d6 = vdot(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) { cp_ = C; goto done; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && 0.0 <= d2 && d6 <= 0.0) {
        double w = d2 / (d2 - d6);
        cp_ = vadd(A, vscale(ac, w));
        goto done;
    }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && 0.0 <= d4 - d3 && d5 - d6 >= 0.0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        cp_ = vadd(B, vscale(vsub(C, B), w));
        goto done;
    }
    {
        double denom = 1.0 / (va + vb + vc);
        double v = vb * denom, w = vc * denom;
        cp_ = vadd(vadd(A, vscale(ab, v)), vscale(ac, w));
    }
done:
    *closest = cp_;
    return vnorm_blas(vsub(zero, cp_));
}

/* b.dot(M) for b[4], M[4,3] inside njit (dgemv): fma(b0,M0,b1*M1) + fma(b2,M2,b3*M3) */
static v3 vec4_mat(const double *b, const v3 *M) {
    v3 r;
    r.x = __builtin_fma(b[0], M[0].x, b[1] * M[1].x) + __builtin_fma(b[2], M[2].x, b[3] * M[3].x);
    r.y = __builtin_fma(b[0], M[0].y, b[1] * M[1].y) + __builtin_fma(b[2], M[2].y, b[3] * M[3].y);
    r.z = __builtin_fma(b[0], M[0].z, b[1] * M[1].z) + __builtin_fma(b[2], M[2].z, b[3] * M[3].z);
    return r;
}

/* mpr.py:368-393 */
static v3 contact_position(const portal_t *p, v3 sd) {
    double b[4];
    b[0] = vdot(vcross(p->v[1], p->v[2]), p->v[3]);
    b[1] = vdot(vcross(p->v[3], p->v[2]), p->v[0]);
    b[2] = vdot(vcross(p->v[0], p->v[1]), p->v[3]);
    b[3] = vdot(vcross(p->v[2], p->v[1]), p->v[0]);
    double sum = ((b[0] + b[1]) + b[2]) + b[3];
    if (sum < EPS) {
        b[0] = 0.0;
        b[1] = vdot(vcross(p->v[2], p->v[3]), sd);
        b[2] = vdot(vcross(p->v[3], p->v[1]), sd);
        b[3] = vdot(vcross(p->v[1], p->v[2]), sd);
        sum = ((b[0] + b[1]) + b[2]) + b[3];
    }
    for (int i = 0; i < 4; ++i) b[i] /= sum;
    v3 p1 = vec4_mat(b, p->v1), p2 = vec4_mat(b, p->v2);
    return vscale(vadd(p1, p2), 0.5);
}

static void mpr_one(const d3d_colliders *c, int64_t ia, int64_t ib, double tol,
                    int max_iterations, int want_pen, uint8_t *out_hit, double *out_depth,
                    double *out_dir, double *out_pos, int32_t *out_status) {
    portal_t p;
    for (int i = 0; i < 4; ++i) p.v[i] = p.v1[i] = p.v2[i] = V3(0, 0, 0);
    int res = discover_portal(c, ia, ib, max_iterations, &p);
    double depth = 0.0;
    v3 dir = V3(0, 0, 0), pos = V3(0, 0, 0);
    int hit;
    *out_status = D3D_UNKNOWN;
    if (res == ORIGIN_OUTSIDE) {
        hit = 0;
    } else if (res == ORIGIN_ON_V1) {
        hit = 1; /* mpr.py:347-353 */
        pos = vscale(vadd(p.v1[1], p.v2[1]), 0.5);
    } else if (res == ORIGIN_ON_SEGMENT) {
        hit = 1; /* mpr.py:356-365 */
        pos = vscale(vadd(p.v1[1], p.v2[1]), 0.5);
        depth = vnorm_blas(p.v[1]);
        dir = vnormalized(p.v[1]);
    } else {
        hit = refine_portal(c, ia, ib, &p, tol);
        if (hit < 0) { hit = 0; *out_status = D3D_ITER_CAP; }
        if (hit && want_pen) { /* mpr.py:318-344 */
            int iterations = 0;
            for (;;) {
                v3 sd = portal_direction(&p);
                v3 n, n1, n2;
                mink_support(c, ia, ib, sd, &n, &n1, &n2);
                if (reach_tolerance(&p, n, sd, tol) || iterations > max_iterations) {
                    v3 cp;
                    depth = point_to_triangle_origin(p.v[1], p.v[2], p.v[3], &cp);
                    if (fabs(depth) < EPS) cp = V3(0, 0, 0);
                    pos = contact_position(&p, portal_direction(&p));
                    dir = vnormalized(cp);
                    break;
                }
                expand_portal(&p, n, n1, n2);
                ++iterations;
            }
        }
    }
    if (*out_status == D3D_UNKNOWN) *out_status = hit ? D3D_INTERSECTION : D3D_NO_INTERSECTION;
    *out_hit = (uint8_t)hit;
    if (want_pen) {
        *out_depth = depth;
        vstore(out_dir, dir);
        vstore(out_pos, pos);
    }
}

void d3do_mpr(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs, double tol,
              int max_iterations, int want_penetration, uint8_t *out_hit, double *out_depth,
              double *out_dir, double *out_pos, int32_t *out_status, int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
    for (int64_t k = 0; k < n_pairs; ++k) {
        d3do_pair_begin(c, pairs[2 * k], pairs[2 * k + 1], mesh_cur);
        mpr_one(c, pairs[2 * k], pairs[2 * k + 1], tol, max_iterations, want_penetration,
                out_hit + k, out_depth + k, out_dir + 3 * k, out_pos + 3 * k, out_status + k);
        d3do_pair_end(c, pairs[2 * k], pairs[2 * k + 1], mesh_cur);
    }
}
