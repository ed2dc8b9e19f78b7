// Jolt closest-point-to-origin simplex solver, all state in registers.
// Restates distance3d/gjk/_gjk_jolt.py:291-711 (cited per function) with the
// reference's evaluation order so that results are bit-identical.
#pragma once
#include "d3d_math.cuh"

// _gjk_jolt.py:291-312 (out of line: nine call sites)
static __device__ __noinline__ void bary_line(v3 a, v3 b, real &u, real &v) {
    v3 ab = b - a;
    real denominator = dot_blas(ab, ab);
    if (denominator < D3D_EPS_SQR) {
        if (dot_blas(a, a) < dot_blas(b, b)) { u = R(1.0); v = R(0.0); }
        else { u = R(0.0); v = R(1.0); }
    } else {
        v = ddiv(-dot_blas(a, ab), denominator);
        u = R(1.0) - v;
    }
}

// _gjk_jolt.py:315-372
D3D_DEV void bary_plane(v3 a, v3 b, v3 c, real &u, real &v, real &w) {
    v3 v0 = b - a, v1 = c - a, v2 = c - b;
    real d00 = dot_blas(v0, v0), d11 = dot_blas(v1, v1), d22 = dot_blas(v2, v2);
    if (d00 <= d22) {
        real d01 = dot_blas(v0, v1);
        real denominator = d00 * d11 - d01 * d01;
        if (fabs(denominator) < D3D_PLANE_DENOM_EPS) {
            if (d00 > d11) { bary_line(a, b, u, v); w = R(0.0); }
            else { bary_line(a, c, u, w); v = R(0.0); }
        } else {
            real a0 = dot_blas(a, v0), a1 = dot_blas(a, v1);
            v = ddiv(d01 * a1 - d11 * a0, denominator);
            w = ddiv(d01 * a0 - d00 * a1, denominator);
            u = R(1.0) - v - w;
        }
    } else {
        real d12 = dot_blas(v1, v2);
        real denominator = d11 * d22 - d12 * d12;
        if (fabs(denominator) < D3D_PLANE_DENOM_EPS) {
            if (d11 > d22) { bary_line(a, c, u, w); v = R(0.0); }
            else { bary_line(b, c, v, w); u = R(0.0); }
        } else {
            real c1 = dot_blas(c, v1), c2 = dot_blas(c, v2);
            u = ddiv(d22 * c1 - d12 * c2, denominator);
            v = ddiv(d11 * c2 - d12 * c1, denominator);
            w = R(1.0) - u - v;
        }
    }
}

// utils.py:73
D3D_DEV real triple(v3 a, v3 b, v3 c) { return dot_blas(a, cross(b, c)); }

// _gjk_jolt.py:375-390
D3D_DEV void bary_tetra(v3 a, v3 b, v3 c, v3 d, real &u, real &v, real &w, real &x) {
    v3 vab = b - a, vac = c - a, vad = d - a;
    real va6 = -triple(b, d - b, c - b);
    real vb6 = -triple(a, vac, vad);
    real vc6 = -triple(a, vad, vab);
    real vd6 = -triple(a, vab, vac);
    real v6 = ddiv(R(1.0), triple(vab, vac, vad));
    u = va6 * v6; v = vb6 * v6; w = vc6 * v6; x = vd6 * v6;
}

// _gjk_jolt.py:393-412
D3D_DEV v3 closest_line(v3 a, v3 b, int &set) {
    real u, v;
    bary_line(a, b, u, v);
    if (v <= R(0.0)) { set = 1; return a; }
    if (u <= R(0.0)) { set = 2; return b; }
    set = 3;
    return a * u + b * v;
}

// Degenerate triangle (_gjk_jolt.py:450-472): best of the three edges.
static __device__ __noinline__ v3 closest_triangle_degenerate(v3 a, v3 b, v3 c, int &set) {
    int closest_set, new_set;
    v3 closest_point = closest_line(a, b, closest_set);
    real best_dist_sq = dot_blas(closest_point, closest_point);
    v3 q = closest_line(a, c, new_set);
    real dist_sq = dot_blas(q, q);
    if (dist_sq < best_dist_sq) {
        closest_point = q;
        best_dist_sq = dist_sq;
        closest_set = (new_set & 1) + ((new_set & 2) << 1);
    }
    q = closest_line(b, c, new_set);
    dist_sq = dot_blas(q, q);
    if (dist_sq < best_dist_sq) {
        closest_point = q;
        closest_set = new_set << 1;
    }
    set = closest_set;
    return closest_point;
}

// _gjk_jolt.py:415-523.  The reference walks the Voronoi regions with early
// returns; here every lane computes all region predicates (same expressions,
// same order of evaluation of the tests) and only the short tails that build the
// result diverge, so lanes of a warp that sit in different regions stay together.
D3D_DEV v3 closest_triangle(v3 a, v3 b, v3 c, int &set) {
    v3 ab = b - a, ac = c - a, bc = c - b;
    bool bc_shorter_than_ac = dot_blas(bc, bc) < dot_blas(ac, ac);
    v3 second = bc_shorter_than_ac ? bc : ac;
    v3 n = cross(ab, second);
    real n_len_sq = dot_blas(n, n);
    if (n_len_sq < D3D_TRI_DEGENERATE_SQR) return closest_triangle_degenerate(a, b, c, set);

    real d1 = dot_blas(ab, -a), d2 = dot_blas(ac, -a);
    real d3 = dot_blas(ab, -b), d4 = dot_blas(ac, -b);
    real d5 = dot_blas(ab, -c), d6 = dot_blas(ac, -c);
    real vc = d1 * d4 - d3 * d2;
    real vb = d5 * d2 - d1 * d6;
    real va = d3 * d6 - d5 * d4;
    real d4_d3 = d4 - d3, d5_d6 = d5 - d6;

    int region;
    if (d1 <= R(0.0) && d2 <= R(0.0)) region = 1;
    else if (d3 >= R(0.0) && d4 <= d3) region = 2;
    else if (vc <= R(0.0) && R(0.0) <= d1 && d3 <= R(0.0)) region = 3;
    else if (d6 >= R(0.0) && d5 <= d6) region = 4;
    else if (vb <= R(0.0) && R(0.0) <= d2 && d6 <= R(0.0)) region = 5;
    else if (va <= R(0.0) && R(0.0) <= d4_d3 && d5_d6 >= R(0.0)) region = 6;
    else region = 7;
    set = region;

    if (region == 7) {
        real s = dot_blas((a + b) + c, n);
        return (n * s) / (R(3.0) * n_len_sq);
    }
    if (region == 3 || region == 5 || region == 6) {
        // a + v*ab | a + w*ac | b + w*bc
        v3 base = (region == 6) ? b : a;
        v3 dir = (region == 3) ? ab : ((region == 5) ? ac : bc);
        real num = (region == 3) ? d1 : ((region == 5) ? d2 : d4_d3);
        real den = (region == 3) ? (d1 - d3) : ((region == 5) ? (d2 - d6) : (d4_d3 + d5_d6));
        real t = ddiv(num, den);
        return base + dir * t;
    }
    return (region == 1) ? a : ((region == 2) ? b : c);
}

// _gjk_jolt.py:526-570: bit i set = origin outside plane i
D3D_DEV int origin_outside_planes(v3 a, v3 b, v3 c, v3 d) {
    v3 ab = b - a, ac = c - a, ad = d - a, bd = d - b, bc = c - b;
    v3 ab_x_ac = cross(ab, ac), ac_x_ad = cross(ac, ad), ad_x_ab = cross(ad, ab),
       bd_x_bc = cross(bd, bc);
    real p0 = dot_blas(a, ab_x_ac), p1 = dot_blas(a, ac_x_ad), p2 = dot_blas(a, ad_x_ab),
           p3 = dot_blas(b, bd_x_bc);
    real s0 = dot_blas(ad, ab_x_ac), s1 = dot_blas(ab, ac_x_ad), s2 = dot_blas(ac, ad_x_ab),
           s3 = -dot_blas(ab, bd_x_bc);
    if (s0 > R(0.0) && s1 > R(0.0) && s2 > R(0.0) && s3 > R(0.0))
        return (p0 >= -D3D_EPS ? 1 : 0) | (p1 >= -D3D_EPS ? 2 : 0) | (p2 >= -D3D_EPS ? 4 : 0) |
               (p3 >= -D3D_EPS ? 8 : 0);
    if (s0 < R(0.0) && s1 < R(0.0) && s2 < R(0.0) && s3 < R(0.0))
        return (p0 <= D3D_EPS ? 1 : 0) | (p1 <= D3D_EPS ? 2 : 0) | (p2 <= D3D_EPS ? 4 : 0) |
               (p3 <= D3D_EPS ? 8 : 0);
    return 0xf;
}

// _gjk_jolt.py:690-711 with closest_point_tetrahedron (:573-631) folded in.
//
// Triangle (3 points) and tetrahedron (4 points) lanes share ONE instance of
// closest_triangle: a 3-point simplex is treated as a tetrahedron whose only
// candidate face is (y0, y1, y2).  Each lane walks its own candidate faces in
// ascending order (the order matters for the strict '<' tie-break), the loop
// trip count is the number of candidate faces, not 4.
D3D_DEV bool closest_point_to_origin(v3 y0, v3 y1, v3 y2, v3 y3, int n_points,
                                     real prev_v_len_sq, v3 &v_out, real &v_len_sq_out,
                                     int &set_out) {
    v3 v = y0;
    int set = 1;
    if (n_points == 2) {
        v = closest_line(y0, y1, set);
    } else if (n_points >= 3) {
        int out = 1;
        if (n_points == 4) out = origin_outside_planes(y0, y1, y2, y3);
        set = 0xf;
        v = V3(R(0.0), R(0.0), R(0.0));
        real best_dist_sq = D3D_MAX_FLOAT;
        int todo = out;
#pragma unroll 1
        while (todo) {
            int f = __ffs(todo) - 1;
            todo &= todo - 1;
            // faces: abc, acd, adb, bdc
            v3 fa = (f == 3) ? y1 : y0;
            v3 fb = (f == 0) ? y1 : ((f == 1) ? y2 : y3);
            v3 fc = (f == 0) ? y2 : ((f == 1) ? y3 : ((f == 2) ? y1 : y2));
            int s;
            v3 q = closest_triangle(fa, fb, fc, s);
            real dist_sq = dot_blas(q, q);
            if (f == 0 || dist_sq < best_dist_sq) {
                best_dist_sq = dist_sq;
                v = q;
                if (f == 0) set = s;
                else if (f == 1) set = (s & 1) + ((s & 6) << 1);
                else if (f == 2) set = (s & 1) + ((s & 2) << 2) + ((s & 4) >> 1);
                else set = ((s & 1) << 1) + ((s & 2) << 2) + (s & 4);
            }
        }
    }
    real v_len_sq = dot_blas(v, v);
    if (v_len_sq < prev_v_len_sq) {
        v_out = v; v_len_sq_out = v_len_sq; set_out = set;
        return true;
    }
    return false;
}
