/* TEST INFRASTRUCTURE (CPU oracle): tetrahedron-pair intersection of the hydroelastic contact
 * model, the consumer of the AABB broad phase.
 * Follows distance3d/hydroelastic_contact/_tetrahedron_intersection.py:7-423,
 * _halfplanes.py:9-71, _barycentric_transform.py:4-9, _mesh_processing.py:4-20,
 * distance3d/utils.py:78-122.  Plain double arithmetic, no FMA contraction; parity with the
 * reference is by tolerance here (1e-9 on the plane and the polygon vertices): the reference
 * goes through LAPACK pinv and libm atan2, which are not reproduced bit for bit. */
#include <math.h>
#include <string.h>
#include "d3d_oracle.h"
#include "vec.h"

#define EPS D3D_EPS

/* _mesh_processing.py:4-20 tetrahedral_mesh_aabbs: out[n,3,2] */
void d3do_tetra_aabbs(const double *points, int64_t n, double *out) {
    for (int64_t t = 0; t < n; ++t)
        for (int k = 0; k < 3; ++k) {
            double lo = points[12 * t + k], hi = lo;
            for (int v = 1; v < 4; ++v) {
                double x = points[12 * t + 3 * v + k];
                if (x < lo) lo = x;
                if (x > hi) hi = x;
            }
            out[6 * t + 2 * k] = lo;
            out[6 * t + 2 * k + 1] = hi;
        }
}

/* _barycentric_transform.py:4-9: X = inverse of [[p0 p1 p2 p3], [1 1 1 1]]; row i of X is the
 * barycentric coordinate function of vertex i: lambda_i(r) = X[i,:3] . r + X[i,3].  Closed
 * form: the plane through the other three vertices, scaled to 1 at vertex i. */
void d3do_barycentric_transform(const double *p, double *X) {
    for (int i = 0; i < 4; ++i) {
        v3 pi = vload(p + 3 * i);
        v3 a = vload(p + 3 * ((i + 1) & 3)), b = vload(p + 3 * ((i + 2) & 3)), c = vload(p + 3 * ((i + 3) & 3));
        v3 n = vcross(vsub(b, a), vsub(c, a));
        double w = vdot_plain(n, vsub(pi, a));
        X[4 * i] = n.x / w;
        X[4 * i + 1] = n.y / w;
        X[4 * i + 2] = n.z / w;
        X[4 * i + 3] = -vdot_plain(n, a) / w;
    }
}

/* utils.py:78-122 */
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

static double cross2d(double ax, double ay, double bx, double by) { return ax * by - ay * bx; }

/* One pair (_tetrahedron_intersection.py:87-140).  Returns 1 when the pair intersects.
 * plane[4]: contact plane in Hesse normal form; poly[max_vertices,3], *n_vertices: contact
 * polygon (counter-clockwise); status: 0 ok, 1 = "same tetrahedron" branch, 2 = more polygon
 * vertices than max_vertices (the excess is dropped). */
static int tetra_pair(const double *t1, const double *e1, const double *X1, const double *t2,
                      const double *e2, const double *X2, double ym1, double ym2, double *plane,
                      double *poly, int max_vertices, int *n_vertices, int *status) {
    *n_vertices = 0;
    *status = 0;
    /* contact_plane :165-216 */
    for (int c = 0; c < 4; ++c) {
        double s1 = 0.0, s2 = 0.0;
        for (int r = 0; r < 4; ++r) {
            s1 += (e1[r] * ym1) * X1[4 * r + c];
            s2 += (e2[r] * ym2) * X2[4 * r + c];
        }
        plane[c] = s1 - s2;
    }
    double norm = sqrt(plane[0] * plane[0] + plane[1] * plane[1] + plane[2] * plane[2]);
    int same = 0;
    if (norm == 0.0) {
        same = 1;
    } else {
        for (int c = 0; c < 4; ++c) plane[c] /= norm;
        plane[3] *= -1.0;
        if (fabs(plane[3]) < 10.0 * EPS) same = 1;
    }
    if (same) { /* _handle_same_tetrahedron :143-162 */
        double sum = ((e2[0] + e2[1]) + e2[2]) + e2[3];
        v3 pp = V3(0, 0, 0);
        for (int r = 0; r < 4; ++r) pp = vadd(pp, vscale(vload(t2 + 3 * r), e2[r] / sum));
        double d = sqrt(pp.x * pp.x + pp.y * pp.y + pp.z * pp.z);
        v3 n = d > 0.0 ? vdiv(pp, d) : V3(0.0, 0.0, 1.0);
        plane[0] = n.x; plane[1] = n.y; plane[2] = n.z; plane[3] = d;
        *status = 1;
        for (int k = 0; k < 3 && k < max_vertices; ++k) vstore(poly + 3 * k, pp);
        *n_vertices = 3 < max_vertices ? 3 : max_vertices;
        return 1;
    }
    v3 n = V3(plane[0], plane[1], plane[2]);
    double d = plane[3];
    /* check_tetrahedra_intersect_contact_plane :219-252, tolerance 1e-6 */
    const double tol = 1e-6;
    double lo1 = 0, hi1 = 0, lo2 = 0, hi2 = 0;
    for (int r = 0; r < 4; ++r) {
        double a = vdot_plain(vload(t1 + 3 * r), n) - d, b = vdot_plain(vload(t2 + 3 * r), n) - d;
        if (r == 0 || a < lo1) lo1 = a;
        if (r == 0 || a > hi1) hi1 = a;
        if (r == 0 || b < lo2) lo2 = b;
        if (r == 0 || b > hi2) hi2 = b;
    }
    if (!(lo1 < -tol && hi1 > tol && lo2 < -tol && hi2 > tol)) return 0;
    /* compute_contact_polygon :377-423 */
    v3 pp = vscale(n, d), bx, by;
    plane_basis(n, &bx, &by);
    /* make_halfplanes :255-292: (point p, direction) per face whose projected normal is not ~0 */
    double hp[8][4];
    int nh = 0;
    for (int i = 0; i < 8; ++i) {
        const double *row = i < 4 ? X1 + 4 * i : X2 + 4 * (i - 4);
        v3 fn = V3(row[0], row[1], row[2]);
        double nx = vdot_plain(fn, bx), ny = vdot_plain(fn, by);
        double ds = -row[3] - vdot_plain(fn, pp);
        double nn = sqrt(nx * nx + ny * ny);
        if (nn > EPS) {
            hp[nh][0] = nx * ds / (nn * nn);
            hp[nh][1] = ny * ds / (nn * nn);
            hp[nh][2] = ny;
            hp[nh][3] = -nx;
            ++nh;
        }
    }
    /* intersect_halfplanes (_halfplanes.py:35-71) */
    double pts[24][2];
    int np_ = 0;
    for (int i = 0; i < nh; ++i)
        for (int j = i + 1; j < nh; ++j) {
            double denom = cross2d(hp[i][2], hp[i][3], hp[j][2], hp[j][3]);
            if (fabs(denom) < EPS) continue;
            double t = cross2d(hp[j][0] - hp[i][0], hp[j][1] - hp[i][1], hp[j][2], hp[j][3]) / denom;
            double px = hp[i][0] + hp[i][2] * t, py = hp[i][1] + hp[i][3] * t;
            int valid = 1;
            for (int k = 0; k < nh && valid; ++k)
                if (k != i && k != j && cross2d(hp[k][2], hp[k][3], px - hp[k][0], py - hp[k][1]) < -EPS)
                    valid = 0;
            if (valid && np_ < 24) { pts[np_][0] = px; pts[np_][1] = py; ++np_; }
        }
    if (np_ < 3) return 0;
    /* order_points :295-313: by angle around the mean */
    double cx = 0, cy = 0;
    for (int k = 0; k < np_; ++k) { cx += pts[k][0]; cy += pts[k][1]; }
    cx /= np_; cy /= np_;
    double ang[24];
    int ord[24];
    for (int k = 0; k < np_; ++k) { ang[k] = atan2(pts[k][1] - cy, pts[k][0] - cx); ord[k] = k; }
    for (int a = 1; a < np_; ++a) { /* insertion sort, stable */
        int o = ord[a], b = a - 1;
        while (b >= 0 && ang[ord[b]] > ang[o]) { ord[b + 1] = ord[b]; --b; }
        ord[b + 1] = o;
    }
    /* filter_unique_points :316-343: drop points within 10 EPS of their predecessor */
    double uq[24][2];
    int nu = 0;
    for (int k = 0; k < np_; ++k) {
        double x = pts[ord[k]][0], y = pts[ord[k]][1];
        if (k > 0) {
            double dx = x - pts[ord[k - 1]][0], dy = y - pts[ord[k - 1]][1];
            if (!(sqrt(dx * dx + dy * dy) > 10.0 * EPS)) continue;
        }
        uq[nu][0] = x; uq[nu][1] = y; ++nu;
    }
    if (nu < 3) return 0;
    /* project_polygon_to_3d :346-374 */
    if (nu > max_vertices) { *status = 2; nu = max_vertices; }
    for (int k = 0; k < nu; ++k)
        vstore(poly + 3 * k, vadd(vadd(vscale(bx, uq[k][0]), vscale(by, uq[k][1])), pp));
    *n_vertices = nu;
    return 1;
}

/* _tetrahedron_intersection.py:7-84 intersect_tetrahedron_pairs for pairs[n,2] (i of mesh 1, j of
 * mesh 2).  X1 / X2 (optional, [n_tetra,4,4]): barycentric transforms; NULL = computed here. */
void d3do_tetra_pairs(const int32_t *pairs, int64_t n_pairs, const double *points1, const double *eps1,
                      const double *X1, const double *points2, const double *eps2, const double *X2,
                      double ym1, double ym2, int max_vertices, uint8_t *out_hit, double *out_plane,
                      int32_t *out_nverts, double *out_poly, int32_t *out_status, int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads)
    for (int64_t k = 0; k < n_pairs; ++k) {
        int64_t i = pairs[2 * k], j = pairs[2 * k + 1];
        double Xa[16], Xb[16];
        const double *xa = X1 ? X1 + 16 * i : Xa, *xb = X2 ? X2 + 16 * j : Xb;
        if (!X1) d3do_barycentric_transform(points1 + 12 * i, Xa);
        if (!X2) d3do_barycentric_transform(points2 + 12 * j, Xb);
        int nv = 0, st = 0;
        out_hit[k] = (uint8_t)tetra_pair(points1 + 12 * i, eps1 + 4 * i, xa, points2 + 12 * j, eps2 + 4 * j,
                                         xb, ym1, ym2, out_plane + 4 * k,
                                         out_poly + (size_t)k * max_vertices * 3, max_vertices, &nv, &st);
        out_nverts[k] = nv;
        out_status[k] = st;
    }
}
