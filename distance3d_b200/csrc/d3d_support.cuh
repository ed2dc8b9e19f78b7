// Support maps, centres and bounding boxes of the packed colliders.
//
// Replaces the per-object Python methods of the reference
// (distance3d/colliders.py:109-646 -> distance3d/geometry.py:138-454,
// distance3d/containment.py:6-229, distance3d/mesh.py:142-191).  One `Collider`
// record is held in registers by the thread (or by every lane of the cooperating
// group) that works on a pair.
#pragma once
#include "../../include/d3d_types.h"
#include "d3d_math.cuh"

struct Collider {
    int type;
    int nv;
    const double *V;  // vertex range in the pool (box / hull: world frame, mesh: local)
    double margin;
    double r00, r01, r02, tx;
    double r10, r11, r12, ty;
    double r20, r21, r22, tz;
    double p0, p1, p2;
};

// 128-bit vectorised loads of the 4x4 pose (rows 0..2) and the parameters.
D3D_DEV Collider load_collider(const d3d_colliders &c, int64_t i) {
    Collider o;
    o.type = __ldg(c.type + i);
    o.nv = __ldg(c.vert_len + i);
    o.V = c.verts + 3 * (int64_t)__ldg(c.vert_off + i);
    o.margin = c.margin ? __ldg(c.margin + i) : 0.0;
    const double2 *T = reinterpret_cast<const double2 *>(c.pose + 16 * i);
    double2 a = __ldg(T + 0), b = __ldg(T + 1), d = __ldg(T + 2), e = __ldg(T + 3),
            f = __ldg(T + 4), g = __ldg(T + 5);
    o.r00 = a.x; o.r01 = a.y; o.r02 = b.x; o.tx = b.y;
    o.r10 = d.x; o.r11 = d.y; o.r12 = e.x; o.ty = e.y;
    o.r20 = f.x; o.r21 = f.y; o.r22 = g.x; o.tz = g.y;
    const double *p = c.param + 3 * i;
    o.p0 = __ldg(p); o.p1 = __ldg(p + 1); o.p2 = __ldg(p + 2);
    return o;
}

// np.dot(pose[:3,:3].T, d)  (dgemv convention)
D3D_DEV v3 rot_t(const Collider &c, v3 d) {
    return V3(gemv_row(c.r00, c.r10, c.r20, d), gemv_row(c.r01, c.r11, c.r21, d),
              gemv_row(c.r02, c.r12, c.r22, d));
}
// utils.py:143 transform_point
D3D_DEV v3 xform(const Collider &c, v3 v) {
    return V3(c.tx + gemv_row(c.r00, c.r01, c.r02, v), c.ty + gemv_row(c.r10, c.r11, c.r12, v),
              c.tz + gemv_row(c.r20, c.r21, c.r22, v));
}

// geometry.py:138-157: vertex i of a box, bit k of i selects +0.5 on axis (2-k)
D3D_DEV v3 box_vertex(const Collider &c, int i) {
    v3 l = V3((i & 4) ? 0.5 * c.p0 : -0.5 * c.p0, (i & 2) ? 0.5 * c.p1 : -0.5 * c.p1,
              (i & 1) ? 0.5 * c.p2 : -0.5 * c.p2);
    return V3(c.tx + dot_blas(l, V3(c.r00, c.r01, c.r02)), c.ty + dot_blas(l, V3(c.r10, c.r11, c.r12)),
              c.tz + dot_blas(l, V3(c.r20, c.r21, c.r22)));
}

// First arg-max of V.dot(d) (colliders.py:132).  G lanes of a warp cooperate:
// lane `lane` scans vertices lane, lane+G, ...; the reduction prefers the larger
// value and, on ties, the lower index, so the result equals numpy's argmax.
template <int G>
D3D_DEV int argmax_dot(const double *V, int n, v3 d, int lane) {
    if (n == 1) return 0;
    if (G == 1) {
        int bi = 0;
        double best = gemv_row(__ldg(V), __ldg(V + 1), __ldg(V + 2), d);
        for (int i = 1; i < n; ++i) {
            double val = gemv_row(__ldg(V + 3 * i), __ldg(V + 3 * i + 1), __ldg(V + 3 * i + 2), d);
            if (val > best) { best = val; bi = i; }
        }
        return bi;
    }
    double best = 0.0;
    int bi = 0x7fffffff;
    bool have = false;
    for (int i = lane; i < n; i += G) {
        double val = gemv_row(__ldg(V + 3 * i), __ldg(V + 3 * i + 1), __ldg(V + 3 * i + 2), d);
        if (!have || val > best) { best = val; bi = i; have = true; }
    }
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, best, off, G);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off, G);
        bool ohave = __shfl_xor_sync(0xffffffffu, (int)have, off, G);
        if (ohave && (!have || ov > best || (ov == best && oi < bi))) {
            best = ov; bi = oi; have = true;
        }
    }
    return bi;
}

// utils.py:78-122
D3D_DEV void plane_basis(v3 n, v3 &x, v3 &y) {
    if (fabs(n.x) >= fabs(n.y)) {
        double len = sqrt(n.x * n.x + n.z * n.z);
        x = V3(-n.z / len, 0.0, n.x / len);
        y = V3(n.y * x.z, n.z * x.x - n.x * x.z, -n.y * x.x);
    } else {
        double len = sqrt(n.y * n.y + n.z * n.z);
        x = V3(0.0, n.z / len, -n.y / len);
        y = V3(n.y * x.z - n.z * x.y, -n.x * x.z, n.x * x.y);
    }
}

template <int G>
D3D_DEV v3 support_unmargined(const Collider &c, v3 d, int lane) {
    switch (c.type) {
    case D3D_SPHERE: {  // geometry.py:341-346
        double s = norm3(d);
        v3 ctr = V3(c.tx, c.ty, c.tz);
        if (s == 0.0) return ctr + V3(0.0, 0.0, c.p0);
        return ctr + (d / s) * c.p0;
    }
    case D3D_CAPSULE: {  // geometry.py:243-256
        v3 l = rot_t(c, d);
        double s = sqrt(l.x * l.x + l.y * l.y + l.z * l.z);
        v3 v;
        if (s == 0.0) v = V3(c.p0, 0.0, 0.0);
        else v = l * (c.p0 / s);
        if (l.z > 0.0) v.z += 0.5 * c.p1;
        else v.z -= 0.5 * c.p1;
        return xform(c, v);
    }
    case D3D_CYLINDER: {  // geometry.py:194-206
        v3 l = rot_t(c, d);
        double s = sqrt(l.x * l.x + l.y * l.y);
        double z = (l.z < 0.0) ? -0.5 * c.p1 : 0.5 * c.p1;
        v3 v;
        if (s == 0.0) v = V3(c.p0, 0.0, z);
        else { double k = c.p0 / s; v = V3(l.x * k, l.y * k, z); }
        return xform(c, v);
    }
    case D3D_ELLIPSOID: {  // geometry.py:282-284
        v3 r = V3(c.p0, c.p1, c.p2);
        v3 l = rot_t(c, d);
        return xform(c, vmul(normalized(vmul(l, r)), r));
    }
    case D3D_BOX: {  // colliders.py:132 over the 8 vertices of geometry.py:157
        if (G == 1) {
            v3 bestv = box_vertex(c, 0);
            double best = gemv_row(bestv.x, bestv.y, bestv.z, d);
#pragma unroll 1
            for (int i = 1; i < 8; ++i) {
                v3 v = box_vertex(c, i);
                double val = gemv_row(v.x, v.y, v.z, d);
                if (val > best) { best = val; bestv = v; }
            }
            return bestv;
        }
        return ld3(c.V + 3 * argmax_dot<G>(c.V, 8, d, lane));
    }
    case D3D_HULL:  // colliders.py:131-132
        return ld3(c.V + 3 * argmax_dot<G>(c.V, c.nv, d, lane));
    case D3D_MESH: {  // mesh.py:182-189 (arg-max form)
        v3 l = rot_t(c, d);
        return xform(c, ld3(c.V + 3 * argmax_dot<G>(c.V, c.nv, l, lane)));
    }
    case D3D_DISK: {  // geometry.py:375-383
        v3 ctr = V3(c.tx, c.ty, c.tz);
        v3 n = V3(c.r02, c.r12, c.r22);
        v3 x, y;
        plane_basis(n, x, y);
        v3 pt = V3(dot_blas(x, d), dot_blas(y, d), 0.0);
        double nrm = norm3(pt);
        if (nrm == 0.0) return ctr;
        pt = pt * (c.p0 / nrm);
        return V3(ctr.x + gemv_row(x.x, y.x, n.x, pt), ctr.y + gemv_row(x.y, y.y, n.y, pt),
                  ctr.z + gemv_row(x.z, y.z, n.z, pt));
    }
    case D3D_ELLIPSE: {  // geometry.py:412-414
        v3 a0 = V3(c.r00, c.r10, c.r20), a1 = V3(c.r01, c.r11, c.r21);
        double l0 = gemv_row(a0.x, a0.y, a0.z, d), l1 = gemv_row(a1.x, a1.y, a1.z, d);
        double w0 = c.p0 * l0, w1 = c.p1 * l1;
        double nrm = norm_dd(w0, w1, 0.0);
        if (nrm != 0.0) { w0 = w0 / nrm; w1 = w1 / nrm; }
        w0 *= c.p0; w1 *= c.p1;
        return V3(c.tx + fma(w1, a1.x, w0 * a0.x), c.ty + fma(w1, a1.y, w0 * a0.y),
                  c.tz + fma(w1, a1.z, w0 * a0.z));
    }
    case D3D_CONE: {  // geometry.py:443-454
        v3 l = rot_t(c, d);
        v3 dp = V3(l.x, l.y, 0.0);
        double nrm = norm3(dp);
        if (nrm == 0.0) dp = V3(0.0, 0.0, 0.0);
        else dp = dp * (c.p0 / nrm);
        v3 pt = (dot_blas(l, dp) >= l.z * c.p1) ? dp : V3(0.0, 0.0, c.p1);
        return xform(c, pt);
    }
    }
    return V3(0.0, 0.0, 0.0);
}

template <int G>
D3D_DEV v3 support(const Collider &c, v3 d, int lane) {
    v3 s = support_unmargined<G>(c, d, lane);
    if (c.margin != 0.0) s = s + normalized(d) * c.margin;  // colliders.py:629-631
    return s;
}

// colliders.py center(); hull / mesh means are sequential column sums (np.mean axis 0)
D3D_DEV v3 center_of(const Collider &c) {
    v3 t = V3(c.tx, c.ty, c.tz);
    if (c.type == D3D_HULL || c.type == D3D_MESH) {
        v3 s = V3(0.0, 0.0, 0.0);
        for (int k = 0; k < c.nv; ++k) s = s + ld3(c.V + 3 * k);
        s = s / (double)c.nv;
        return c.type == D3D_HULL ? s : xform(c, s);
    }
    if (c.type == D3D_CONE)
        return V3(t.x + 0.5 * c.p1 * c.r02, t.y + 0.5 * c.p1 * c.r12, t.z + 0.5 * c.p1 * c.r22);
    return t;
}
