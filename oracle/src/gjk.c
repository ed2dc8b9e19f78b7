/* TEST INFRASTRUCTURE (CPU oracle): Jolt-variant GJK.
 * Follows distance3d/gjk/_gjk_jolt.py line by line (cited per function). */
#include <omp.h>
#include "d3d_oracle.h"
#include "vec.h"

v3 d3do_support_s(const d3d_colliders *c, int64_t i, v3 d, int32_t *cur);
void d3do_pair_begin(const d3d_colliders *c, int64_t ia, int64_t ib, int32_t *cur);
void d3do_pair_end(const d3d_colliders *c, int64_t ia, int64_t ib, const int32_t *cur);
/* MeshGraph vertex of collider A / B of the pair this thread is working on */
static _Thread_local int32_t mesh_cur[2];

#define EPS D3D_EPS
#define EPS_SQR (D3D_EPS * D3D_EPS)

/* _gjk_jolt.py:291-312 */
static void bary_line(v3 a, v3 b, double *u, double *v) {
    v3 ab = vsub(b, a);
    double denominator = vdot(ab, ab);
    if (denominator < EPS_SQR) {
        if (vdot(a, a) < vdot(b, b)) { *u = 1.0; *v = 0.0; }
        else { *u = 0.0; *v = 1.0; }
    } else {
        *v = -vdot(a, ab) / denominator;
        *u = 1.0 - *v;
    }
}

/* _gjk_jolt.py:315-372 */
static void bary_plane(v3 a, v3 b, v3 c, double *u, double *v, double *w) {
    v3 v0 = vsub(b, a), v1 = vsub(c, a), v2 = vsub(c, b);
    double d00 = vdot(v0, v0), d11 = vdot(v1, v1), d22 = vdot(v2, v2);
    if (d00 <= d22) {
        double d01 = vdot(v0, v1);
        double denominator = d00 * d11 - d01 * d01;
        if (fabs(denominator) < EPS) {
            if (d00 > d11) { bary_line(a, b, u, v); *w = 0.0; }
            else { bary_line(a, c, u, w); *v = 0.0; }
        } else {
            double a0 = vdot(a, v0), a1 = vdot(a, v1);
            *v = (d01 * a1 - d11 * a0) / denominator;
            *w = (d01 * a0 - d00 * a1) / denominator;
            *u = 1.0 - *v - *w;
        }
    } else {
        double d12 = vdot(v1, v2);
        double denominator = d11 * d22 - d12 * d12;
        if (fabs(denominator) < EPS) {
            if (d11 > d22) { bary_line(a, c, u, w); *v = 0.0; }
            else { bary_line(b, c, v, w); *u = 0.0; }
        } else {
            double c1 = vdot(c, v1), c2 = vdot(c, v2);
            *u = (d22 * c1 - d12 * c2) / denominator;
            *v = (d11 * c2 - d12 * c1) / denominator;
            *w = 1.0 - *u - *v;
        }
    }
}

/* utils.py:73 */
static double stp(v3 a, v3 b, v3 c) { return vdot(a, vcross(b, c)); }

/* _gjk_jolt.py:375-390 */
static void bary_tetra(v3 a, v3 b, v3 c, v3 d, double *u, double *v, double *w, double *x) {
    v3 vab = vsub(b, a), vac = vsub(c, a), vad = vsub(d, a);
    double va6 = -stp(b, vsub(d, b), vsub(c, b));
    double vb6 = -stp(a, vac, vad);
    double vc6 = -stp(a, vad, vab);
    double vd6 = -stp(a, vab, vac);
    double v6 = 1.0 / stp(vab, vac, vad);
    *u = va6 * v6; *v = vb6 * v6; *w = vc6 * v6; *x = vd6 * v6;
}

/* _gjk_jolt.py:393-412 */
static v3 closest_line(v3 a, v3 b, int *set) {
    double u, v;
    bary_line(a, b, &u, &v);
    if (v <= 0.0) { *set = 1; return a; }
    if (u <= 0.0) { *set = 2; return b; }
    *set = 3;
    return vadd(vscale(a, u), vscale(b, v));
}

/* _gjk_jolt.py:415-523 */
static v3 closest_triangle(v3 a, v3 b, v3 c, int *set) {
    v3 ab = vsub(b, a), ac = vsub(c, a), bc = vsub(c, b);
    int bc_shorter_than_ac = vdot(bc, bc) < vdot(ac, ac);
    v3 n = bc_shorter_than_ac ? vcross(ab, bc) : vcross(ab, ac);
    double n_len_sq = vdot(n, n);

    if (n_len_sq < EPS_SQR) {
        int closest_set, new_set;
        v3 closest_point = closest_line(a, b, &closest_set);
        double best_dist_sq = vdot(closest_point, closest_point);
        v3 q = closest_line(a, c, &new_set);
        double dist_sq = vdot(q, q);
        if (dist_sq < best_dist_sq) {
            closest_point = q;
            best_dist_sq = dist_sq;
            closest_set = (new_set & 1) + ((new_set & 2) << 1);
        }
        q = closest_line(b, c, &new_set);
        dist_sq = vdot(q, q);
        if (dist_sq < best_dist_sq) {
            closest_point = q;
            closest_set = new_set << 1;
        }
        *set = closest_set;
        return closest_point;
    }

    v3 ap = vneg(a);
    double d1 = vdot(ab, ap), d2 = vdot(ac, ap);
    if (d1 <= 0.0 && d2 <= 0.0) { *set = 1; return a; }

    v3 bp = vneg(b);
    double d3 = vdot(ab, bp), d4 = vdot(ac, bp);
    if (d3 >= 0.0 && d4 <= d3) { *set = 2; return b; }

    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && 0.0 <= d1 && d3 <= 0.0) {
        double v = d1 / (d1 - d3);
        *set = 3;
        return vadd(a, vscale(ab, v));
    }

    v3 cp = vneg(c);
    double d5 = vdot(ab, cp), d6 = vdot(ac, cp);
    if (d6 >= 0.0 && d5 <= d6) { *set = 4; return c; }

    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && 0.0 <= d2 && d6 <= 0.0) {
        double w = d2 / (d2 - d6);
        *set = 5;
        return vadd(a, vscale(ac, w));
    }

    double va = d3 * d6 - d5 * d4;
    double d4_d3 = d4 - d3, d5_d6 = d5 - d6;
    if (va <= 0.0 && 0.0 <= d4_d3 && d5_d6 >= 0.0) {
        double w = d4_d3 / (d4_d3 + d5_d6);
        *set = 6;
        return vadd(b, vscale(bc, w));
    }

    *set = 7;
    /* n * (a + b + c).dot(n) / (3.0 * n_len_sq) */
    double s = vdot(vadd(vadd(a, b), c), n);
    return vdiv(vscale(n, s), 3.0 * n_len_sq);
}

/* _gjk_jolt.py:526-570; returns 4-bit mask, bit i = origin outside plane i */
static int origin_outside_planes(v3 a, v3 b, v3 c, v3 d) {
    v3 ab = vsub(b, a), ac = vsub(c, a), ad = vsub(d, a), bd = vsub(d, b), bc = vsub(c, b);
    v3 ab_x_ac = vcross(ab, ac), ac_x_ad = vcross(ac, ad), ad_x_ab = vcross(ad, ab),
       bd_x_bc = vcross(bd, bc);
    double signp[4] = {vdot(a, ab_x_ac), vdot(a, ac_x_ad), vdot(a, ad_x_ab), vdot(b, bd_x_bc)};
    double signd[4] = {vdot(ad, ab_x_ac), vdot(ab, ac_x_ad), vdot(ac, ad_x_ab),
                       -vdot(ab, bd_x_bc)};
    int all_pos = 1, all_neg = 1, mask = 0;
    for (int i = 0; i < 4; ++i) {
        if (!(signd[i] > 0.0)) all_pos = 0;
        if (!(signd[i] < 0.0)) all_neg = 0;
    }
    if (all_pos) {
        for (int i = 0; i < 4; ++i) if (signp[i] >= -EPS) mask |= 1 << i;
    } else if (all_neg) {
        for (int i = 0; i < 4; ++i) if (signp[i] <= EPS) mask |= 1 << i;
    } else {
        mask = 0xf;
    }
    return mask;
}

/* _gjk_jolt.py:573-631 */
static v3 closest_tetrahedron(v3 a, v3 b, v3 c, v3 d, int *set) {
    int closest_set = 0xf, new_set;
    v3 closest_point = V3(0.0, 0.0, 0.0);
    double best_dist_sq = D3D_MAX_FLOAT;
    int out = origin_outside_planes(a, b, c, d);
    if (out & 1) {
        closest_point = closest_triangle(a, b, c, &closest_set);
        best_dist_sq = vdot(closest_point, closest_point);
    }
    if (out & 2) {
        v3 q = closest_triangle(a, c, d, &new_set);
        double dist_sq = vdot(q, q);
        if (dist_sq < best_dist_sq) {
            best_dist_sq = dist_sq;
            closest_point = q;
            closest_set = (new_set & 1) + ((new_set & 6) << 1);
        }
    }
    if (out & 4) {
        v3 q = closest_triangle(a, d, b, &new_set);
        double dist_sq = vdot(q, q);
        if (dist_sq < best_dist_sq) {
            best_dist_sq = dist_sq;
            closest_point = q;
            closest_set = (new_set & 1) + ((new_set & 2) << 2) + ((new_set & 4) >> 1);
        }
    }
    if (out & 8) {
        v3 q = closest_triangle(b, d, c, &new_set);
        double dist_sq = vdot(q, q);
        if (dist_sq < best_dist_sq) {
            closest_point = q;
            closest_set = ((new_set & 1) << 1) + ((new_set & 2) << 2) + (new_set & 4);
        }
    }
    *set = closest_set;
    return closest_point;
}

/* _gjk_jolt.py:690-711 */
static int closest_point_to_origin(const v3 *Y, int n_points, double prev_v_len_sq, v3 *v_out,
                                   double *v_len_sq_out, int *set_out) {
    v3 v;
    int set;
    switch (n_points) {
    case 1: set = 1; v = Y[0]; break;
    case 2: v = closest_line(Y[0], Y[1], &set); break;
    case 3: v = closest_triangle(Y[0], Y[1], Y[2], &set); break;
    default: v = closest_tetrahedron(Y[0], Y[1], Y[2], Y[3], &set); break;
    }
    double v_len_sq = vdot(v, v);
    if (v_len_sq < prev_v_len_sq) {
        *v_out = v; *v_len_sq_out = v_len_sq; *set_out = set;
        return 1;
    }
    return 0;
}

/* _gjk_jolt.py:634-640 */
static double max_y_len_sq(const v3 *Y, int n) {
    double m = vdot(Y[0], Y[0]);
    for (int i = 1; i < n; ++i) { double l = vdot(Y[i], Y[i]); m = m > l ? m : l; /* max(m, l) */ }
    return m;
}

/* one pair of _gjk_jolt.py:138-221 (+ :224-288) */
static void gjk_distance_one(const d3d_colliders *c, int64_t ia, int64_t ib, double tolerance,
                             double max_distance_squared, double sanity_check, double *out_dist,
                             double *out_a, double *out_b, double *out_Y, int32_t *out_npoints,
                             int32_t *out_iters, int32_t *out_status) {
    v3 Y[4], P[4], Q[4];
    for (int i = 0; i < 4; ++i) Y[i] = P[i] = Q[i] = V3(0, 0, 0);
    int n_points = 0;
    double tolerance_sq = tolerance * tolerance;
    v3 sd = V3(1.0, 0.0, 0.0);
    double v_len_sq = vdot(sd, sd);
    double prev_v_len_sq = D3D_MAX_FLOAT;
    int state = D3D_UNKNOWN, iters = 0;

    while (state == D3D_UNKNOWN) {
        if (iters >= D3D_GJK_ITER_CAP) { state = D3D_ITER_CAP; break; }
        ++iters;
        v3 p = d3do_support_s(c, ia, sd, &mesh_cur[0]);
        v3 q = d3do_support_s(c, ib, vneg(sd), &mesh_cur[1]);
        /* _distance_loop */
        v3 w = vsub(p, q);
        double dot = vdot(sd, w);
        if (dot < 0.0 && dot * dot > v_len_sq * max_distance_squared) { state = D3D_CLIPPED; break; }
        Y[n_points] = w; P[n_points] = p; Q[n_points] = q;
        ++n_points;
        v3 v_new; double v_len_sq_new; int simplex;
        if (closest_point_to_origin(Y, n_points, prev_v_len_sq, &v_new, &v_len_sq_new, &simplex)) {
            sd = v_new; v_len_sq = v_len_sq_new;
        } else {
            --n_points;
            simplex = 0;
            for (int i = 0; i < n_points; ++i) simplex |= 1 << i;
        }
        if (simplex == 0xf) { v_len_sq = 0.0; state = D3D_INTERSECTION; break; }
        /* update_simplex_ypq :654-664 */
        int nn = 0;
        for (int i = 0; i < n_points; ++i)
            if (simplex & (1 << i)) { Y[nn] = Y[i]; P[nn] = P[i]; Q[nn] = Q[i]; ++nn; }
        n_points = nn;
        if (v_len_sq <= tolerance_sq) { v_len_sq = 0.0; state = D3D_INTERSECTION; break; }
        if (v_len_sq <= EPS * max_y_len_sq(Y, n_points)) { v_len_sq = 0.0; state = D3D_INTERSECTION; break; }
        sd = vscale(sd, -1.0);
        if (!(prev_v_len_sq >= v_len_sq)) { state = D3D_MONOTONICITY; break; }
        if (prev_v_len_sq - v_len_sq <= EPS * prev_v_len_sq) { state = D3D_NO_INTERSECTION; break; }
        prev_v_len_sq = v_len_sq;
    }

    *out_iters = iters;
    *out_npoints = n_points;
    for (int i = 0; i < 4; ++i) vstore(out_Y + 3 * i, Y[i]);
    if (state == D3D_CLIPPED || state == D3D_ITER_CAP || state == D3D_MONOTONICITY) {
        *out_dist = D3D_MAX_FLOAT;
        vstore(out_a, V3(0, 0, 0)); vstore(out_b, V3(0, 0, 0));
        *out_status = state;
        return;
    }
    /* calculate_closest_points :667-687 */
    v3 a = V3(0, 0, 0), b = V3(0, 0, 0);
    if (n_points == 1) { a = P[0]; b = Q[0]; }
    else if (n_points == 2) {
        double u, v; bary_line(Y[0], Y[1], &u, &v);
        a = vadd(vscale(P[0], u), vscale(P[1], v));
        b = vadd(vscale(Q[0], u), vscale(Q[1], v));
    } else if (n_points == 3) {
        double u, v, w; bary_plane(Y[0], Y[1], Y[2], &u, &v, &w);
        a = vadd(vadd(vscale(P[0], u), vscale(P[1], v)), vscale(P[2], w));
        b = vadd(vadd(vscale(Q[0], u), vscale(Q[1], v)), vscale(Q[2], w));
    } else if (n_points == 4) {
        double u, v, w, x; bary_tetra(Y[0], Y[1], Y[2], Y[3], &u, &v, &w, &x);
        a = vadd(vadd(vadd(vscale(P[0], u), vscale(P[1], v)), vscale(P[2], w)), vscale(P[3], x));
        b = vadd(vadd(vadd(vscale(Q[0], u), vscale(Q[1], v)), vscale(Q[2], w)), vscale(Q[3], x));
    }
    double check_value = fabs(vdot(sd, sd) - v_len_sq);
    if (!(check_value < sanity_check)) state = D3D_SANITY_FAILED;
    double dist = sqrt(v_len_sq);
    if (dist < EPS) { a = b = vscale(vadd(a, b), 0.5); }
    *out_dist = dist;
    vstore(out_a, a); vstore(out_b, b);
    *out_status = state;
}

void d3do_gjk_distance(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs,
                       double tolerance, double max_distance_squared, double sanity_check,
                       double *out_dist, double *out_a, double *out_b, double *out_Y,
                       int32_t *out_npoints, int32_t *out_iters, int32_t *out_status,
                       int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
    for (int64_t k = 0; k < n_pairs; ++k) {
        d3do_pair_begin(c, pairs[2 * k], pairs[2 * k + 1], mesh_cur);
        gjk_distance_one(c, pairs[2 * k], pairs[2 * k + 1], tolerance, max_distance_squared,
                         sanity_check, out_dist + k, out_a + 3 * k, out_b + 3 * k, out_Y + 12 * k,
                         out_npoints + k, out_iters + k, out_status + k);
        d3do_pair_end(c, pairs[2 * k], pairs[2 * k + 1], mesh_cur);
    }
}

/* one pair of _gjk_jolt.py:29-135 */
static void gjk_intersection_one(const d3d_colliders *c, int64_t ia, int64_t ib, double tolerance,
                                 uint8_t *out_hit, int32_t *out_iters, int32_t *out_status) {
    v3 Y[4];
    for (int i = 0; i < 4; ++i) Y[i] = V3(0, 0, 0);
    int n_points = 0;
    double tolerance_sq = tolerance * tolerance;
    double prev_v_len_sq = D3D_MAX_FLOAT;
    v3 sd = V3(1.0, 0.0, 0.0);
    int state = D3D_UNKNOWN, iters = 0;
    while (state == D3D_UNKNOWN) {
        if (iters >= D3D_GJK_ITER_CAP) { state = D3D_ITER_CAP; break; }
        ++iters;
        v3 p = d3do_support_s(c, ia, sd, &mesh_cur[0]);
        v3 q = d3do_support_s(c, ib, vneg(sd), &mesh_cur[1]);
        v3 w = vsub(p, q);
        if (vdot(sd, w) < -EPS) { state = D3D_NO_INTERSECTION; break; }
        Y[n_points++] = w;
        v3 v_new; double v_len_sq; int simplex;
        if (!closest_point_to_origin(Y, n_points, prev_v_len_sq, &v_new, &v_len_sq, &simplex)) {
            state = D3D_NO_INTERSECTION; break;
        }
        sd = v_new;
        if (simplex == 0xf) { state = D3D_INTERSECTION; break; }
        if (v_len_sq <= tolerance_sq) { state = D3D_INTERSECTION; break; }
        if (v_len_sq <= EPS * max_y_len_sq(Y, n_points)) { state = D3D_INTERSECTION; break; }
        sd = vscale(sd, -1.0);
        if (!(prev_v_len_sq >= v_len_sq)) { state = D3D_MONOTONICITY; break; }
        if (prev_v_len_sq - v_len_sq <= EPS * prev_v_len_sq) { state = D3D_NO_INTERSECTION; break; }
        prev_v_len_sq = v_len_sq;
        int nn = 0; /* update_simplex_y :643-651 */
        for (int i = 0; i < n_points; ++i) if (simplex & (1 << i)) Y[nn++] = Y[i];
        n_points = nn;
    }
    *out_hit = (state == D3D_INTERSECTION);
    *out_iters = iters;
    *out_status = state;
}

void d3do_gjk_intersection(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs,
                           double tolerance, uint8_t *out_hit, int32_t *out_iters,
                           int32_t *out_status, int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
    for (int64_t k = 0; k < n_pairs; ++k) {
        d3do_pair_begin(c, pairs[2 * k], pairs[2 * k + 1], mesh_cur);
        gjk_intersection_one(c, pairs[2 * k], pairs[2 * k + 1], tolerance, out_hit + k,
                             out_iters + k, out_status + k);
        d3do_pair_end(c, pairs[2 * k], pairs[2 * k + 1], mesh_cur);
    }
}

int d3do_max_threads(void) { return omp_get_max_threads(); }
