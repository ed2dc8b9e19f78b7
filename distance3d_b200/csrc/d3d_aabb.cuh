// Axis-aligned bounding box of one packed collider (shared by d3d_aabb and the fused
// self-collision filter).
#pragma once
#include "d3d_support.cuh"

// colliders.py aabb() -> containment.py:6-229 for collider i: lo[3], hi[3].  Hull and mesh vertex
// scans are serial (the LBVH path computes boxes once per pose update, the narrow phase dominates).
D3D_DEV void collider_aabb(const d3d_colliders &c, int64_t i, double *lo, double *hi) {
    Collider col = load_collider(c, i);
    double t[3] = {col.tx(), col.ty(), col.tz()};
    double e[3] = {0.0, 0.0, 0.0};
    bool have_extent = true;
    const double R[3][3] = {{col.r00(), col.r01(), col.r02()}, {col.r10(), col.r11(), col.r12()},
                            {col.r20(), col.r21(), col.r22()}};
    const double p[3] = {col.p0(), col.p1(), col.p2()};
    switch (col.type) {
    case D3D_SPHERE:  // containment.py:44
        e[0] = e[1] = e[2] = col.p0();
        break;
    case D3D_CAPSULE:  // containment.py:121
        for (int k = 0; k < 3; ++k) e[k] = 0.5 * col.p1() * fabs(R[k][2]) + col.p0();
        break;
    case D3D_CYLINDER:  // containment.py:94-95
        for (int k = 0; k < 3; ++k) {
            double a = R[k][2];
            e[k] = 0.5 * col.p1() * fabs(a) + col.p0() * sqrt(1.0 - a * a);
        }
        break;
    case D3D_ELLIPSOID: {  // containment.py:144-147 (reproduced as is)
        double E[3][3];
        for (int k = 0; k < 3; ++k) {
            double col_[3];
            for (int j = 0; j < 3; ++j) col_[j] = R[j][k] * p[k];
            double nrm = sqrt((col_[0] * col_[0] + col_[1] * col_[1]) + col_[2] * col_[2]);
            for (int j = 0; j < 3; ++j) E[j][k] = col_[j] / nrm * p[k];
        }
        for (int j = 0; j < 3; ++j) {
            double best = 0.0;
            for (int r = 0; r < 3; ++r) {
                double val = dot_blas(V3(R[r][0], R[r][1], R[r][2]), V3(E[j][0], E[j][1], E[j][2]));
                if (r == 0 || val > best) best = val;
            }
            e[j] = best;
        }
        break;
    }
    case D3D_BOX: {  // containment.py:66-67
        for (int v = 0; v < 8; ++v) {
            v3 x = box_vertex(col, v);
            double xs[3] = {x.x, x.y, x.z};
            for (int k = 0; k < 3; ++k) {
                if (v == 0 || xs[k] < lo[k]) lo[k] = xs[k];
                if (v == 0 || xs[k] > hi[k]) hi[k] = xs[k];
            }
        }
        have_extent = false;
        break;
    }
    case D3D_HULL:  // containment.py:22
    case D3D_MESH: {  // colliders.py:234-237
        for (int v = 0; v < col.nv; ++v) {
            v3 x = ld3(col.V + 3 * v);
            if (col.type == D3D_MESH)
                x = V3(col.tx() + dot_blas(x, V3(col.r00(), col.r01(), col.r02())),
                       col.ty() + dot_blas(x, V3(col.r10(), col.r11(), col.r12())),
                       col.tz() + dot_blas(x, V3(col.r20(), col.r21(), col.r22())));
            double xs[3] = {x.x, x.y, x.z};
            for (int k = 0; k < 3; ++k) {
                if (v == 0 || xs[k] < lo[k]) lo[k] = xs[k];
                if (v == 0 || xs[k] > hi[k]) hi[k] = xs[k];
            }
        }
        have_extent = false;
        break;
    }
    case D3D_DISK:  // containment.py:173
        for (int k = 0; k < 3; ++k) e[k] = col.p0() * sqrt(1.0 - R[k][2] * R[k][2]);
        break;
    case D3D_ELLIPSE:  // containment.py:228
        for (int k = 0; k < 3; ++k) {
            double u = col.p0() * R[k][0], w = col.p1() * R[k][1];
            e[k] = sqrt(u * u + w * w);
        }
        break;
    case D3D_CONE:  // containment.py:199-203
        for (int k = 0; k < 3; ++k) {
            double pa = t[k];
            double pb = t[k] + col.p1() * R[k][2];
            double a = pb - pa;
            double ee = sqrt(1.0 - a * a / (col.p1() * col.p1()));
            double l = pa - ee * col.p0(), h = pa + ee * col.p0();
            lo[k] = l < pb ? l : pb;
            hi[k] = h > pb ? h : pb;
        }
        have_extent = false;
        break;
    }
    if (have_extent)
        for (int k = 0; k < 3; ++k) { lo[k] = t[k] - e[k]; hi[k] = t[k] + e[k]; }
    if (col.margin() != 0.0)  // colliders.py:639-643
        for (int k = 0; k < 3; ++k) { lo[k] -= col.margin(); hi[k] += col.margin(); }
}
