/* TEST INFRASTRUCTURE (CPU oracle): support maps, centres and AABBs.
 * Follows distance3d/geometry.py:138-454, distance3d/colliders.py:109-646,
 * distance3d/containment.py:6-229, distance3d/utils.py:12-143. */
#include "d3d_oracle.h"
#include "vec.h"

static const double BOX_COORDS[8][3] = {
    {-0.5, -0.5, -0.5}, {-0.5, -0.5, 0.5}, {-0.5, 0.5, -0.5}, {-0.5, 0.5, 0.5},
    {0.5, -0.5, -0.5},  {0.5, -0.5, 0.5},  {0.5, 0.5, -0.5},  {0.5, 0.5, 0.5}};

/* geometry.py:157: t + (BOX_COORDS * size).dot(R.T)  (dgemm = ddot convention) */
void d3do_box_vertices(const double *T, const double *size, double *out) {
    for (int i = 0; i < 8; ++i) {
        v3 l = V3(BOX_COORDS[i][0] * size[0], BOX_COORDS[i][1] * size[1],
                  BOX_COORDS[i][2] * size[2]);
        for (int k = 0; k < 3; ++k) {
            v3 r = V3(T[4 * k], T[4 * k + 1], T[4 * k + 2]);
            out[3 * i + k] = T[4 * k + 3] + vdot(l, r);
        }
    }
}

void d3do_prepare(const d3d_colliders *c, double *verts_out) {
    for (int64_t i = 0; i < c->n; ++i)
        if (c->type[i] == D3D_BOX)
            d3do_box_vertices(c->pose + 16 * i, c->param + 3 * i,
                              verts_out + 3 * (int64_t)c->vert_off[i]);
}

/* utils.py:78-122 plane_basis_from_normal */
static void plane_basis(v3 n, v3 *x, v3 *y) {
    if (fabs(n.x) >= fabs(n.y)) {
        double len = sqrt(n.x * n.x + n.z * n.z);
        *x = V3(-n.z / len, 0.0, n.x / len);
        *y = V3(n.y * x->z, n.z * x->x - n.x * x->z, -n.y * x->x);
    } else {
        double len = sqrt(n.y * n.y + n.z * n.z);
        *x = V3(0.0, n.z / len, -n.y / len);
        *y = V3(n.y * x->z - n.z * x->y, -n.x * x->z, n.x * x->y);
    }
}

/* colliders.py:132 / mesh.py:186: first argmax of V.dot(d); a single row goes
 * through ddot, two or more rows through dgemv. */
static int64_t argmax_dot(const double *V, int64_t n, v3 d) {
    if (n == 1) return 0;
    int64_t best = 0;
    double best_val = gemv_row(V[0], V[1], V[2], d);
    for (int64_t i = 1; i < n; ++i) {
        double val = gemv_row(V[3 * i], V[3 * i + 1], V[3 * i + 2], d);
        if (val > best_val) { best_val = val; best = i; }
    }
    return best;
}

/* mesh.py:90-139 hill_climb_mesh_extreme; g = adjacency record (include/d3d_types.h).
 * `search_direction.dot(vertex_diff)` is a BLAS ddot. */
static int64_t hill_climb(const double *V, const int32_t *g, v3 l, int64_t start) {
    const double eps = 10.0 * D3D_EPS; /* PROJECTION_LENGTH_EPSILON, mesh.py:9 */
    int64_t best = start;
    for (int k = 1; k <= 6; ++k) { /* shortcut_connections, mesh.py:44-47, 121-127 */
        int64_t ci = g[k];
        if (vdot(l, vsub(vload(V + 3 * ci), vload(V + 3 * best))) > eps) best = ci;
    }
    int converged = 0;
    while (!converged) { /* mesh.py:129-137 */
        converged = 1;
        int32_t lo = g[7 + best], hi = g[8 + best]; /* the list of the vertex the round starts on */
        for (int32_t j = lo; j < hi; ++j) {
            int64_t ci = g[j];
            if (vdot(l, vsub(vload(V + 3 * ci), vload(V + 3 * best))) > eps) {
                best = ci;
                converged = 0;
            }
        }
    }
    return best;
}

/* first_idx of a fresh reference object (mesh.py:29) unless the caller supplies one */
int32_t d3do_mesh_start(const d3d_colliders *c, int64_t i) {
    if (c->type[i] != D3D_MESH) return 0;
    if (c->mesh_start != 0 && c->mesh_start[i] >= 0) return c->mesh_start[i];
    if (c->graph_off != 0 && c->graph_off[i] >= 0) return c->graph[c->graph_off[i]];
    return 0;
}

/* per-pair MeshGraph state: cur[0] for collider ia, cur[1] for ib (see d3d_types.h) */
void d3do_pair_begin(const d3d_colliders *c, int64_t ia, int64_t ib, int32_t *cur) {
    cur[0] = d3do_mesh_start(c, ia);
    cur[1] = d3do_mesh_start(c, ib);
}
void d3do_pair_end(const d3d_colliders *c, int64_t ia, int64_t ib, const int32_t *cur) {
    if (c->mesh_last == 0) return;
    if (c->type[ia] == D3D_MESH) c->mesh_last[ia] = cur[0];
    if (c->type[ib] == D3D_MESH) c->mesh_last[ib] = cur[1];
}

static v3 support_unmargined(const d3d_colliders *c, int64_t i, v3 d, int32_t *cur) {
    const double *T = c->pose + 16 * i;
    const double *p = c->param + 3 * i;
    switch (c->type[i]) {
    case D3D_SPHERE: { /* geometry.py:341-346 */
        v3 center = V3(T[3], T[7], T[11]);
        double s = vnorm_blas(d);
        if (s == 0.0) return vadd(center, V3(0.0, 0.0, p[0]));
        return vadd(center, vscale(vdiv(d, s), p[0]));
    }
    case D3D_CAPSULE: { /* geometry.py:243-256 */
        v3 l = rot_t_apply(T, d);
        double s = sqrt(l.x * l.x + l.y * l.y + l.z * l.z);
        v3 v;
        if (s == 0.0) v = V3(p[0], 0.0, 0.0);
        else v = vscale(l, p[0] / s);
        if (l.z > 0.0) v.z += 0.5 * p[1];
        else v.z -= 0.5 * p[1];
        return transform_point(T, v);
    }
    case D3D_CYLINDER: { /* geometry.py:194-206 */
        v3 l = rot_t_apply(T, d);
        double s = sqrt(l.x * l.x + l.y * l.y);
        double z = (l.z < 0.0) ? -0.5 * p[1] : 0.5 * p[1];
        v3 v;
        if (s == 0.0) v = V3(p[0], 0.0, z);
        else { double k = p[0] / s; v = V3(l.x * k, l.y * k, z); }
        return transform_point(T, v);
    }
    case D3D_ELLIPSOID: { /* geometry.py:282-284 */
        v3 r = V3(p[0], p[1], p[2]);
        v3 l = rot_t_apply(T, d);
        v3 v = vmul(vnormalized(vmul(l, r)), r);
        return transform_point(T, v);
    }
    case D3D_BOX:
    case D3D_HULL: { /* colliders.py:131-132 */
        const double *V = c->verts + 3 * (int64_t)c->vert_off[i];
        return vload(V + 3 * argmax_dot(V, c->vert_len[i], d));
    }
    case D3D_MESH: { /* mesh.py:79-87 (hill climbing); without a graph mesh.py:182-189 */
        const double *V = c->verts + 3 * (int64_t)c->vert_off[i];
        v3 l = rot_t_apply(T, d);
        int64_t idx;
        if (c->graph_off != 0 && c->graph_off[i] >= 0) {
            idx = hill_climb(V, c->graph + c->graph_off[i], l, *cur);
            *cur = (int32_t)idx; /* vertex caching, mesh.py:85 */
        } else {
            idx = argmax_dot(V, c->vert_len[i], l);
        }
        return transform_point(T, vload(V + 3 * idx));
    }
    case D3D_DISK: { /* geometry.py:375-383 */
        v3 center = V3(T[3], T[7], T[11]);
        v3 n = V3(T[2], T[6], T[10]);
        v3 x, y;
        plane_basis(n, &x, &y);
        /* np.dot(R.T, d) with R F-viewed contiguous: dgemv 'n' kernel = ddot convention */
        v3 pt = V3(vdot(x, d), vdot(y, d), 0.0);
        double norm = vnorm_blas(pt);
        if (norm == 0.0) return center;
        double k = p[0] / norm;
        pt = vscale(pt, k);
        return V3(center.x + gemv_row(x.x, y.x, n.x, pt), center.y + gemv_row(x.y, y.y, n.y, pt),
                  center.z + gemv_row(x.z, y.z, n.z, pt));
    }
    case D3D_ELLIPSE: { /* geometry.py:412-414 */
        v3 center = V3(T[3], T[7], T[11]);
        v3 a0 = V3(T[0], T[4], T[8]), a1 = V3(T[1], T[5], T[9]);
        double l0 = gemv_row(a0.x, a0.y, a0.z, d), l1 = gemv_row(a1.x, a1.y, a1.z, d);
        double w0 = p[0] * l0, w1 = p[1] * l1;
        long double e0 = w0, e1 = w1;
        double norm = (double)sqrtl(e0 * e0 + e1 * e1);
        if (norm != 0.0) { w0 = w0 / norm; w1 = w1 / norm; }
        w0 *= p[0]; w1 *= p[1];
        /* np.dot(local_vertex, axes): dgemv over 2 terms */
        return V3(center.x + __builtin_fma(w1, a1.x, w0 * a0.x),
                  center.y + __builtin_fma(w1, a1.y, w0 * a0.y),
                  center.z + __builtin_fma(w1, a1.z, w0 * a0.z));
    }
    case D3D_CONE: { /* geometry.py:443-454 */
        v3 l = rot_t_apply(T, d);
        v3 dp = V3(l.x, l.y, 0.0);
        double norm = vnorm_blas(dp);
        if (norm == 0.0) dp = V3(0.0, 0.0, 0.0);
        else dp = vscale(dp, p[0] / norm);
        v3 pt;
        if (vdot(l, dp) >= l.z * p[1]) pt = dp;
        else pt = V3(0.0, 0.0, p[1]);
        return transform_point(T, pt);
    }
    }
    return V3(0, 0, 0);
}

v3 d3do_support_s(const d3d_colliders *c, int64_t i, v3 d, int32_t *cur) {
    v3 s = support_unmargined(c, i, d, cur);
    if (c->margin != 0 && c->margin[i] != 0.0) /* colliders.py:629-631 */
        s = vadd(s, vscale(vnormalized(d), c->margin[i]));
    return s;
}

void d3do_support(const d3d_colliders *c, int64_t idx, const double *d, double *out) {
    int32_t cur = d3do_mesh_start(c, idx);
    vstore(out, d3do_support_s(c, idx, vload(d), &cur));
    if (c->mesh_last != 0 && c->type[idx] == D3D_MESH) c->mesh_last[idx] = cur;
}

/* colliders.py center() */
v3 d3do_center_v(const d3d_colliders *c, int64_t i) {
    const double *T = c->pose + 16 * i;
    const double *p = c->param + 3 * i;
    v3 t = V3(T[3], T[7], T[11]);
    switch (c->type[i]) {
    case D3D_HULL: { /* colliders.py:135 np.mean(V, axis=0): sequential column sums / n */
        const double *V = c->verts + 3 * (int64_t)c->vert_off[i];
        int64_t n = c->vert_len[i];
        v3 s = V3(0, 0, 0);
        for (int64_t k = 0; k < n; ++k) s = vadd(s, vload(V + 3 * k));
        return vdiv(s, (double)n);
    }
    case D3D_MESH: { /* colliders.py:225-226 */
        const double *V = c->verts + 3 * (int64_t)c->vert_off[i];
        int64_t n = c->vert_len[i];
        v3 s = V3(0, 0, 0);
        for (int64_t k = 0; k < n; ++k) s = vadd(s, vload(V + 3 * k));
        return transform_point(T, vdiv(s, (double)n));
    }
    case D3D_CONE: /* colliders.py:583-584 */
        return V3(t.x + 0.5 * p[1] * T[2], t.y + 0.5 * p[1] * T[6], t.z + 0.5 * p[1] * T[10]);
    default:
        return t;
    }
}

void d3do_center(const d3d_colliders *c, int64_t idx, double *out) {
    vstore(out, d3do_center_v(c, idx));
}

static void aabb_of_points(const double *V, int64_t n, double *out) {
    for (int k = 0; k < 3; ++k) {
        double lo = V[k], hi = V[k];
        for (int64_t i = 1; i < n; ++i) {
            double x = V[3 * i + k];
            if (x < lo) lo = x;
            if (x > hi) hi = x;
        }
        out[2 * k] = lo;
        out[2 * k + 1] = hi;
    }
}

/* colliders.py aabb() -> containment.py */
void d3do_aabb(const d3d_colliders *c, double *out) {
    for (int64_t i = 0; i < c->n; ++i) {
        const double *T = c->pose + 16 * i;
        const double *p = c->param + 3 * i;
        double *o = out + 6 * i;
        double t[3] = {T[3], T[7], T[11]};
        double e[3];
        int have_extent = 1;
        switch (c->type[i]) {
        case D3D_SPHERE: /* containment.py:44 */
            e[0] = e[1] = e[2] = p[0];
            break;
        case D3D_CAPSULE: /* containment.py:121 */
            for (int k = 0; k < 3; ++k) e[k] = 0.5 * p[1] * fabs(T[4 * k + 2]) + p[0];
            break;
        case D3D_CYLINDER: /* containment.py:94-95 */
            for (int k = 0; k < 3; ++k) {
                double a = T[4 * k + 2];
                e[k] = 0.5 * p[1] * fabs(a) + p[0] * sqrt(1.0 - a * a);
            }
            break;
        case D3D_ELLIPSOID: { /* containment.py:144-147 */
            double E[3][3];
            for (int k = 0; k < 3; ++k) { /* column k */
                double col[3];
                for (int j = 0; j < 3; ++j) col[j] = T[4 * j + k] * p[k];
                /* np.linalg.norm(axis=0): sqrt of sequential sum of squares */
                double nrm = sqrt((col[0] * col[0] + col[1] * col[1]) + col[2] * col[2]);
                for (int j = 0; j < 3; ++j) E[j][k] = col[j] / nrm * p[k];
            }
            /* np.dot(R, E.T) (dgemm, ddot convention), max over axis 0 */
            for (int j = 0; j < 3; ++j) {
                double best = 0.0;
                for (int r = 0; r < 3; ++r) {
                    double val = vdot(V3(T[4 * r], T[4 * r + 1], T[4 * r + 2]),
                                      V3(E[j][0], E[j][1], E[j][2]));
                    if (r == 0 || val > best) best = val;
                }
                e[j] = best;
            }
            break;
        }
        case D3D_BOX: /* containment.py:66-67 (vertices regenerated) */
        case D3D_HULL: /* containment.py:22 */
            aabb_of_points(c->verts + 3 * (int64_t)c->vert_off[i], c->vert_len[i], o);
            have_extent = 0;
            break;
        case D3D_MESH: { /* colliders.py:234-237: t + V.dot(R.T) */
            const double *V = c->verts + 3 * (int64_t)c->vert_off[i];
            int64_t n = c->vert_len[i];
            for (int k = 0; k < 3; ++k) {
                double lo = 0, hi = 0;
                for (int64_t q = 0; q < n; ++q) {
                    double x = t[k] + vdot(vload(V + 3 * q), V3(T[4 * k], T[4 * k + 1], T[4 * k + 2]));
                    if (q == 0 || x < lo) lo = x;
                    if (q == 0 || x > hi) hi = x;
                }
                o[2 * k] = lo;
                o[2 * k + 1] = hi;
            }
            have_extent = 0;
            break;
        }
        case D3D_DISK: /* containment.py:173 */
            for (int k = 0; k < 3; ++k) {
                double nk = T[4 * k + 2];
                e[k] = p[0] * sqrt(1.0 - nk * nk);
            }
            break;
        case D3D_ELLIPSE: /* containment.py:228 */
            for (int k = 0; k < 3; ++k) {
                double u = p[0] * T[4 * k], w = p[1] * T[4 * k + 1];
                e[k] = sqrt(u * u + w * w);
            }
            break;
        case D3D_CONE: /* containment.py:199-203 */
            for (int k = 0; k < 3; ++k) {
                double pa = t[k];
                double pb = t[k] + p[1] * T[4 * k + 2];
                double a = pb - pa;
                double ee = sqrt(1.0 - a * a / (p[1] * p[1]));
                double lo = pa - ee * p[0], hi = pa + ee * p[0];
                o[2 * k] = lo < pb ? lo : pb;
                o[2 * k + 1] = hi > pb ? hi : pb;
            }
            have_extent = 0;
            break;
        default:
            e[0] = e[1] = e[2] = 0.0;
        }
        if (have_extent)
            for (int k = 0; k < 3; ++k) {
                o[2 * k] = t[k] - e[k];
                o[2 * k + 1] = t[k] + e[k];
            }
        if (c->margin != 0 && c->margin[i] != 0.0) /* colliders.py:639-643 */
            for (int k = 0; k < 3; ++k) {
                o[2 * k] -= c->margin[i];
                o[2 * k + 1] += c->margin[i];
            }
    }
}

/* dnrm2 semantics (x87) for n 3-vectors: test hook for the device emulation */
void d3do_norm(const double *v, int64_t n, double *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = vnorm_blas(vload(v + 3 * i));
}
