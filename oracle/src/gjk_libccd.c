/* TEST INFRASTRUCTURE (CPU oracle): libccd-style boolean GJK, the reference's second
 * intersection test, used as an independent cross-check of the Jolt
 * variant (distance3d/test/test_gjk.py:341-354).
 * Follows distance3d/gjk/_gjk_libccd.py:14-266, distance3d/distance/_triangle.py:12-89,
 * distance3d/colliders.py first_vertex() of every collider type. */
#include <omp.h>
#include "d3d_oracle.h"
#include "vec.h"

v3 d3do_support_s(const d3d_colliders *c, int64_t i, v3 d, int32_t *cur);
void d3do_pair_begin(const d3d_colliders *c, int64_t ia, int64_t ib, int32_t *cur);
void d3do_pair_end(const d3d_colliders *c, int64_t ia, int64_t ib, const int32_t *cur);
static _Thread_local int32_t mesh_cur[2];

#define EPS D3D_EPS
#define EPS_SQRT 1.4901161193847656e-08 /* math.sqrt(EPSILON), _gjk_libccd.py:11 */
enum { NO_CONTACT = -1, CONTINUE = 0, CONTACT = 1 };

/* colliders.py:128,217,269,322,370,422,475,530,586,626 first_vertex() */
v3 d3do_first_vertex(const d3d_colliders *c, int64_t i) {
    const double *T = c->pose + 16 * i, *p = c->param + 3 * i;
    v3 t = V3(T[3], T[7], T[11]), z = V3(T[2], T[6], T[10]);
    switch (c->type[i]) {
    case D3D_SPHERE: return vadd(t, V3(0.0, 0.0, p[0]));
    case D3D_CAPSULE: return vsub(t, vscale(z, p[0] + 0.5 * p[1]));
    case D3D_ELLIPSOID: return vadd(t, vscale(z, p[2]));
    case D3D_CYLINDER: return vadd(t, vscale(z, 0.5 * p[1]));
    case D3D_CONE: return vadd(t, vscale(z, p[1]));
    case D3D_BOX:
    case D3D_HULL: return vload(c->verts + 3 * (int64_t)c->vert_off[i]);
    case D3D_MESH: return transform_point(T, vload(c->verts + 3 * (int64_t)c->vert_off[i]));
    case D3D_DISK: { /* c + r * x, x from plane_basis_from_normal(normal) (utils.py:78-122) */
        v3 x;
        if (fabs(z.x) >= fabs(z.y)) {
            double len = sqrt(z.x * z.x + z.z * z.z);
            x = V3(-z.z / len, 0.0, z.x / len);
        } else {
            double len = sqrt(z.y * z.y + z.z * z.z);
            x = V3(0.0, z.z / len, -z.y / len);
        }
        return vadd(t, vscale(x, p[0]));
    }
    case D3D_ELLIPSE: return vadd(t, vscale(V3(T[0], T[4], T[8]), p[0]));
    }
    return t;
}

/* distance/_triangle.py:12-89: distance only */
static double point_to_triangle(v3 P, v3 A, v3 B, v3 C) {
    v3 ab = vsub(B, A), ac = vsub(C, A);
    v3 ap = vsub(P, A);
    double d1 = vdot(ab, ap), d2 = vdot(ac, ap);
    v3 cp_;
    if (d1 <= 0.0 && d2 <= 0.0) { cp_ = A; goto done; }
    v3 bp = vsub(P, B);
    double d3 = vdot(ab, bp), d4 = vdot(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) { cp_ = B; goto done; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && 0.0 <= d1 && d3 <= 0.0) { cp_ = vadd(A, vscale(ab, d1 / (d1 - d3))); goto done; }
    v3 cp = vsub(P, C);
    double d5 = vdot(ab, cp), d6 = vdot(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) { cp_ = C; goto done; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && 0.0 <= d2 && d6 <= 0.0) { cp_ = vadd(A, vscale(ac, d2 / (d2 - d6))); goto done; }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && 0.0 <= d4 - d3 && d5 - d6 >= 0.0) {
        cp_ = vadd(B, vscale(vsub(C, B), (d4 - d3) / ((d4 - d3) + (d5 - d6))));
        goto done;
    }
    {
        double denom = 1.0 / (va + vb + vc);
        cp_ = vadd(vadd(A, vscale(ab, vb * denom)), vscale(ac, vc * denom));
    }
done:
    return vnorm_blas(vsub(P, cp_));
}

static v3 triple_cross(v3 a, v3 b, v3 c) { return vcross(vcross(a, b), c); } /* :263-266 */
static int all_close(v3 a, v3 b) {
    return fabs(a.x - b.x) < EPS && fabs(a.y - b.y) < EPS && fabs(a.z - b.z) < EPS;
}
static int sgn(double x) { return (x > 0.0) - (x < 0.0); } /* np.sign */

/* :112-134 */
static int line_segment(v3 *v, v3 *sd, int *n) {
    v3 A = v[1], B = v[0];
    v3 AB = vsub(B, A), AO = vneg(A);
    double on_ab = vdot(AB, AO);
    v3 tmp = vcross(AB, AO);
    if (fabs(vdot(tmp, tmp)) < EPS && on_ab > 0.0) { *n = 2; return CONTACT; }
    if (on_ab < EPS) { v[0] = A; *n = 1; *sd = AO; }
    else { *sd = triple_cross(AB, AO, AB); *n = 2; }
    return CONTINUE;
}

/* :177-187 */
static void triangle_ab(v3 A, v3 B, v3 AB, v3 AO, v3 *v, v3 *sd, int *n) {
    if (vdot(AB, AO) > -EPS) { v[0] = B; v[1] = A; *n = 2; *sd = triple_cross(AB, AO, AB); }
    else { v[0] = A; *n = 1; *sd = AO; }
}

/* :137-174 */
static int triangle(v3 *v, v3 *sd, int *n) {
    v3 A = v[2], B = v[1], C = v[0];
    if (fabs(point_to_triangle(V3(0, 0, 0), A, B, C)) < EPS_SQRT) { *n = 1; return CONTACT; }
    if (all_close(A, B) || all_close(A, C)) { *n = 0; return NO_CONTACT; }
    v3 AO = vneg(A), AB = vsub(B, A), AC = vsub(C, A);
    v3 ABC = vcross(AB, AC);
    if (vdot(vcross(ABC, AC), AO) > -EPS) {
        if (vdot(AC, AO) > -EPS) { v[1] = A; *n = 2; *sd = triple_cross(AC, AO, AC); }
        else triangle_ab(A, B, AB, AO, v, sd, n);
    } else {
        if (vdot(vcross(AB, ABC), AO) > -EPS) triangle_ab(A, B, AB, AO, v, sd, n);
        else if (vdot(ABC, AO) > -EPS) { *n = 3; *sd = ABC; }
        else { v[0] = B; v[1] = C; *n = 3; *sd = vneg(ABC); }
    }
    return CONTINUE;
}

/* :190-260 */
static int tetrahedron(v3 *v, v3 *sd, int *n) {
    v3 A = v[3], B = v[2], C = v[1], D = v[0];
    if (fabs(point_to_triangle(A, B, C, D)) < EPS_SQRT) { *n = 0; return NO_CONTACT; }
    v3 O = V3(0, 0, 0);
    if (point_to_triangle(O, A, B, C) < EPS_SQRT || point_to_triangle(O, A, C, D) < EPS_SQRT ||
        point_to_triangle(O, A, B, D) < EPS_SQRT || point_to_triangle(O, B, C, D) < EPS_SQRT) {
        *n = 3;
        return CONTACT;
    }
    v3 AO = vneg(A), AB = vsub(B, A), AC = vsub(C, A), AD = vsub(D, A);
    v3 ABC = vcross(AB, AC), ACD = vcross(AC, AD), ADB = vcross(AD, AB);
    int b_on_acd = sgn(vdot(ACD, AB)), c_on_adb = sgn(vdot(ADB, AC)), d_on_abc = sgn(vdot(ABC, AD));
    int ab_o = sgn(vdot(ACD, AO)) == b_on_acd, ac_o = sgn(vdot(ADB, AO)) == c_on_adb,
        ad_o = sgn(vdot(ABC, AO)) == d_on_abc;
    if (ab_o && ac_o && ad_o) { *n = 4; return CONTACT; }
    if (!ab_o) { v[2] = A; }                                /* :249-260 */
    else if (!ac_o) { v[1] = D; v[0] = B; v[2] = A; }
    else { v[0] = C; v[1] = B; v[2] = A; }
    return triangle(v, sd, n);
}

/* :56-91 */
static int gjk_libccd_one(const d3d_colliders *c, int64_t ia, int64_t ib, int max_iterations,
                          int32_t *out_iters) {
    v3 v[4];
    v[0] = vsub(d3do_first_vertex(c, ia), d3do_first_vertex(c, ib));
    int n = 1;
    v3 sd = vneg(v[0]);
    int it;
    for (it = 0; it < max_iterations; ++it) {
        v3 sp = vsub(d3do_support_s(c, ia, sd, &mesh_cur[0]), d3do_support_s(c, ib, vneg(sd), &mesh_cur[1]));
        if (out_iters) *out_iters = it + 1;
        if (vdot(sp, sp) < EPS) return 1;
        if (vdot(sp, sd) < -EPS_SQRT) return 0;
        v[n++] = sp;
        int state = n == 2 ? line_segment(v, &sd, &n) : (n == 3 ? triangle(v, &sd, &n) : tetrahedron(v, &sd, &n));
        if (state == CONTACT) return 1;
        if (state == NO_CONTACT) return 0;
        if (fabs(vdot(sd, sd)) < EPS) return 0;
    }
    return 0;
}

void d3do_gjk_intersection_libccd(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs,
                                  int max_iterations, uint8_t *out_hit, int32_t *out_iters,
                                  int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
    for (int64_t k = 0; k < n_pairs; ++k) {
        int32_t it = 0;
        d3do_pair_begin(c, pairs[2 * k], pairs[2 * k + 1], mesh_cur);
        out_hit[k] = (uint8_t)gjk_libccd_one(c, pairs[2 * k], pairs[2 * k + 1], max_iterations, &it);
        d3do_pair_end(c, pairs[2 * k], pairs[2 * k + 1], mesh_cur);
        if (out_iters) out_iters[k] = it;
    }
}
