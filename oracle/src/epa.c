/* TEST INFRASTRUCTURE (CPU oracle): expanding polytope algorithm.
 * Follows distance3d/epa.py:9-202 including its array-aliasing behaviour
 * (fix_ccw_normal_direction "swap", epa.py:139-146; the stale closest-face view
 * used after the last iteration, epa.py:60,77). */
#include <stdlib.h>
#include <string.h>
#include "d3d_oracle.h"
#include "vec.h"

v3 d3do_support_s(const d3d_colliders *c, int64_t i, v3 d, int32_t *cur);
void d3do_pair_begin(const d3d_colliders *c, int64_t ia, int64_t ib, int32_t *cur);
void d3do_pair_end(const d3d_colliders *c, int64_t ia, int64_t ib, const int32_t *cur);
/* MeshGraph vertex of collider A / B of the pair this thread is working on */
static _Thread_local int32_t mesh_cur[2];

typedef struct { v3 v[3]; v3 n; } face_t;
typedef struct { v3 a, b; } edge_t;

/* numpy-level np.linalg.norm of a 1-D array: sqrt(x.dot(x)) */
static double norm_numpy(v3 a) { return sqrt(vdot(a, a)); }

/* epa.py:99-102 */
static void compute_normal(face_t *f) {
    f->n = vnormalized(vcross(vsub(f->v[1], f->v[0]), vsub(f->v[2], f->v[0])));
}

static void epa_one(const d3d_colliders *c, int64_t ia, int64_t ib, const double *Yp,
                    int max_iter, int max_loose_edges, int max_faces, double epsilon,
                    face_t *faces, edge_t *loose, double *out_mtv, uint8_t *out_success,
                    int32_t *out_nfaces, int32_t *out_iters, int32_t *out_status) {
    v3 Y[4];
    for (int i = 0; i < 4; ++i) Y[i] = vload(Yp + 3 * i);
    memset(faces, 0, sizeof(face_t) * (size_t)max_faces);
    /* epa.py:89-97 */
    faces[0].v[0] = Y[0]; faces[0].v[1] = Y[1]; faces[0].v[2] = Y[2];
    faces[1].v[0] = Y[0]; faces[1].v[1] = Y[2]; faces[1].v[2] = Y[3];
    faces[2].v[0] = Y[0]; faces[2].v[1] = Y[3]; faces[2].v[2] = Y[1];
    faces[3].v[0] = Y[1]; faces[3].v[1] = Y[3]; faces[3].v[2] = Y[2];
    int n_faces = 4;
    for (int i = 0; i < 4; ++i) compute_normal(&faces[i]);

    int closest = 0, it;
    *out_status = D3D_INTERSECTION;
    for (it = 0; it < max_iter; ++it) {
        /* epa.py:104-109: first argmin of sum(v0 * n) */
        double min_dist = 0.0;
        closest = 0;
        for (int i = 0; i < n_faces; ++i) {
            double d = vdot_plain(faces[i].v[0], faces[i].n);
            if (i == 0 || d < min_dist) { min_dist = d; closest = i; }
        }
        v3 sd = faces[closest].n;
        v3 new_point = vsub(d3do_support_s(c, ia, sd, &mesh_cur[0]), d3do_support_s(c, ib, vneg(sd), &mesh_cur[1]));
        double proj = vdot(new_point, sd);
        if (proj - min_dist < epsilon) { /* epa.py:67-70 */
            vstore(out_mtv, vscale(faces[closest].n, proj));
            *out_success = 1; *out_nfaces = n_faces; *out_iters = it + 1;
            return;
        }
        /* epa.py:157-165 */
        int n_loose = 0;
        for (int i = 0; i < n_faces; ++i) {
            if (vdot(faces[i].n, vsub(new_point, faces[i].v[0])) > epsilon) {
                /* epa.py:167-187 */
                for (int j = 0; j < 3; ++j) {
                    v3 e0 = faces[i].v[j], e1 = faces[i].v[(j + 1) % 3];
                    int found = 0;
                    for (int k = 0; k < n_loose; ++k) {
                        if (norm_numpy(vsub(loose[k].b, e0)) < epsilon &&
                            norm_numpy(vsub(loose[k].a, e1)) < epsilon) {
                            loose[k] = loose[n_loose - 1];
                            --n_loose;
                            found = 1;
                            break;
                        }
                    }
                    if (!found) {
                        if (n_loose >= max_loose_edges) break;
                        loose[n_loose].a = e0; loose[n_loose].b = e1;
                        ++n_loose;
                    }
                }
                faces[i] = faces[n_faces - 1]; /* epa.py:118-120 */
                --n_faces;
                --i;
            }
        }
        /* epa.py:126-137 */
        for (int i = 0; i < n_loose; ++i) {
            if (!(n_faces < max_faces)) {
                *out_status = D3D_EPA_MAX_FACES;
                vstore(out_mtv, V3(0, 0, 0));
                *out_success = 0; *out_nfaces = n_faces; *out_iters = it + 1;
                return;
            }
            face_t *f = &faces[n_faces];
            f->v[0] = loose[i].a; f->v[1] = loose[i].b; f->v[2] = new_point;
            compute_normal(f);
            if (norm_numpy(f->n) < 0.5) continue;
            if (vdot(f->v[0], f->n) + 1e-6 < 0.0) { /* epa.py:139-146 (aliasing: v0 <- v1 only) */
                f->v[0] = f->v[1];
                f->n = vneg(f->n);
            }
            ++n_faces;
        }
    }
    /* epa.py:76-78: `closest_face` is a view of slot `closest` as it is NOW */
    vstore(out_mtv, vscale(faces[closest].n, vdot(faces[closest].v[0], faces[closest].n)));
    *out_success = 0; *out_nfaces = n_faces; *out_iters = it;
}

void d3do_epa(const d3d_colliders *c, const int32_t *pairs, int64_t n_pairs, const double *Y,
              int max_iter, int max_loose_edges, int max_faces, double epsilon,
              double *out_mtv, uint8_t *out_success, int32_t *out_nfaces, int32_t *out_iters,
              int32_t *out_status, double *out_faces, int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
    {
        face_t *faces = malloc(sizeof(face_t) * (size_t)(max_faces + 1));
        edge_t *loose = malloc(sizeof(edge_t) * (size_t)(max_loose_edges + 1));
#pragma omp for schedule(dynamic, 64)
        for (int64_t k = 0; k < n_pairs; ++k) {
            d3do_pair_begin(c, pairs[2 * k], pairs[2 * k + 1], mesh_cur);
            epa_one(c, pairs[2 * k], pairs[2 * k + 1], Y + 12 * k, max_iter, max_loose_edges,
                    max_faces, epsilon, faces, loose, out_mtv + 3 * k, out_success + k,
                    out_nfaces + k, out_iters + k, out_status + k);
            d3do_pair_end(c, pairs[2 * k], pairs[2 * k + 1], mesh_cur);
            if (out_faces) {
                double *o = out_faces + (size_t)k * max_faces * 12;
                for (int i = 0; i < max_faces; ++i) {
                    vstore(o + 12 * i, faces[i].v[0]); vstore(o + 12 * i + 3, faces[i].v[1]);
                    vstore(o + 12 * i + 6, faces[i].v[2]); vstore(o + 12 * i + 9, faces[i].n);
                }
            }
        }
        free(faces);
        free(loose);
    }
}
