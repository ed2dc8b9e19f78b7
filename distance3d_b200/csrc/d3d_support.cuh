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

// Register-resident collider record.  All support / centre code below is written
// against the accessor interface (r00() .. tz(), p0() .. p2(), margin(), type, nv, V)
// so the same code also runs on a record staged in shared memory (ColliderSmem).
struct Collider {
    int type;
    int nv;
    const double *V;  // vertex range in the pool (box / hull: world frame, mesh: local)
    real m_margin;
    real m[15];     // r00 r01 r02 tx  r10 r11 r12 ty  r20 r21 r22 tz  p0 p1 p2
    const int32_t *g;  // MeshGraph adjacency record (d3d_types.h) or nullptr
    mutable int cur;   // MeshGraph: vertex the last support call ended on (mesh.py:85)
    D3D_DEV const int32_t *graph() const { return g; }
    D3D_DEV int mesh_cur() const { return cur; }
    D3D_DEV void set_mesh_cur(int v, int) const { cur = v; }
    D3D_DEV real r00() const { return m[0]; }
    D3D_DEV real r01() const { return m[1]; }
    D3D_DEV real r02() const { return m[2]; }
    D3D_DEV real tx() const { return m[3]; }
    D3D_DEV real r10() const { return m[4]; }
    D3D_DEV real r11() const { return m[5]; }
    D3D_DEV real r12() const { return m[6]; }
    D3D_DEV real ty() const { return m[7]; }
    D3D_DEV real r20() const { return m[8]; }
    D3D_DEV real r21() const { return m[9]; }
    D3D_DEV real r22() const { return m[10]; }
    D3D_DEV real tz() const { return m[11]; }
    D3D_DEV real p0() const { return m[12]; }
    D3D_DEV real p1() const { return m[13]; }
    D3D_DEV real p2() const { return m[14]; }
    D3D_DEV real margin() const { return m_margin; }
};

// The same record staged in shared memory: field f of the owning thread lives at
// base[f * STRIDE] (STRIDE = threads per block: consecutive threads hit consecutive
// 8-byte words, conflict free; STRIDE = 1 for a warp-private record).
#define D3D_COLLIDER_FIELDS 16
template <int STRIDE>
struct ColliderSmem {
    int type;
    int nv;
    const double *V;
    const real *base;
    const int32_t *gpool;  // d3d_colliders::graph
    D3D_DEV real f(int i) const { return base[i * STRIDE]; }
    // MeshGraph keeps its hill-climbing state in the parameter fields it does not use:
    // field 12 = current vertex, field 13 = offset of the adjacency record (-1: none)
    D3D_DEV const int32_t *graph() const {
        int off = (int)f(13);
        return off < 0 ? nullptr : gpool + off;
    }
    D3D_DEV int mesh_cur() const { return (int)f(12); }
    D3D_DEV void set_mesh_cur(int v, int lane) const {
        if (STRIDE == 1) {  // warp-shared record: every lane has read the old value, lane 0 writes
            __syncwarp();
            if (lane == 0) const_cast<real *>(base)[12] = (real)v;
            __syncwarp();
        } else {
            const_cast<real *>(base)[12 * STRIDE] = (real)v;
        }
    }
    D3D_DEV real r00() const { return f(0); }
    D3D_DEV real r01() const { return f(1); }
    D3D_DEV real r02() const { return f(2); }
    D3D_DEV real tx() const { return f(3); }
    D3D_DEV real r10() const { return f(4); }
    D3D_DEV real r11() const { return f(5); }
    D3D_DEV real r12() const { return f(6); }
    D3D_DEV real ty() const { return f(7); }
    D3D_DEV real r20() const { return f(8); }
    D3D_DEV real r21() const { return f(9); }
    D3D_DEV real r22() const { return f(10); }
    D3D_DEV real tz() const { return f(11); }
    D3D_DEV real p0() const { return f(12); }
    D3D_DEV real p1() const { return f(13); }
    D3D_DEV real p2() const { return f(14); }
    D3D_DEV real margin() const { return f(15); }
};

// A hull / mesh without vertices (outside the contract of d3d_types.h) is read as this one point
// instead of whatever lies next to its empty range.
__device__ const double d3d_origin_vertex[3] = {0.0, 0.0, 0.0};
D3D_DEV void guard_empty_range(int &nv, const double *&V) {
    if (nv < 1) { nv = 1; V = d3d_origin_vertex; }
}

// MeshGraph: offset of the adjacency record and the vertex a pair starts from.
D3D_DEV void mesh_graph_of(const d3d_colliders &c, int64_t i, int &off, int &start) {
    off = c.graph_off ? __ldg(c.graph_off + i) : -1;
    start = c.mesh_start ? __ldg(c.mesh_start + i) : -1;
    if (start < 0) start = off >= 0 ? __ldg(c.graph + off) : 0;
}

// 128-bit vectorised loads of the 4x4 pose (rows 0..2) and the parameters.
D3D_DEV Collider load_collider(const d3d_colliders &c, int64_t i) {
    Collider o;
    o.type = __ldg(c.type + i);
    o.g = nullptr;
    o.cur = 0;
    if (o.type == D3D_MESH) {
        int off;
        mesh_graph_of(c, i, off, o.cur);
        if (off >= 0) o.g = c.graph + off;
    }
    o.nv = __ldg(c.vert_len + i);
    o.V = c.verts + 3 * (int64_t)__ldg(c.vert_off + i);
    guard_empty_range(o.nv, o.V);
    o.m_margin = c.margin ? __ldg(c.margin + i) : R(0.0);
    const double2 *T = reinterpret_cast<const double2 *>(c.pose + 16 * i);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double2 a = __ldg(T + k);
        o.m[2 * k] = a.x;
        o.m[2 * k + 1] = a.y;
    }
    const double *p = c.param + 3 * i;
    o.m[12] = __ldg(p); o.m[13] = __ldg(p + 1); o.m[14] = __ldg(p + 2);
    return o;
}

// Stage collider i into the calling thread's shared-memory record.
template <int STRIDE>
D3D_DEV ColliderSmem<STRIDE> stage_collider(const d3d_colliders &c, int64_t i, real *base) {
    ColliderSmem<STRIDE> o;
    o.type = __ldg(c.type + i);
    o.nv = __ldg(c.vert_len + i);
    o.V = c.verts + 3 * (int64_t)__ldg(c.vert_off + i);
    guard_empty_range(o.nv, o.V);
    o.base = base;
    o.gpool = c.graph;
    const double2 *T = reinterpret_cast<const double2 *>(c.pose + 16 * i);
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double2 a = __ldg(T + k);
        base[(2 * k) * STRIDE] = a.x;
        base[(2 * k + 1) * STRIDE] = a.y;
    }
    const double *p = c.param + 3 * i;
    base[12 * STRIDE] = __ldg(p);
    base[13 * STRIDE] = __ldg(p + 1);
    base[14 * STRIDE] = __ldg(p + 2);
    base[15 * STRIDE] = c.margin ? __ldg(c.margin + i) : R(0.0);
    if (o.type == D3D_MESH) {
        int off, start;
        mesh_graph_of(c, i, off, start);
        base[12 * STRIDE] = (real)start;
        base[13 * STRIDE] = (real)off;
    }
    return o;
}

// Warp-per-pair kernels: ONE record per warp in shared memory (STRIDE = 1).  Every lane reads
// the collider (same addresses, one broadcast transaction) and keeps type / nv / V in
// registers; lane 0 stores the record.  The caller synchronises the warp before the first use.
D3D_DEV ColliderSmem<1> stage_collider_warp(const d3d_colliders &c, int64_t i, real *base, int lane) {
    ColliderSmem<1> o;
    o.type = __ldg(c.type + i);
    o.nv = __ldg(c.vert_len + i);
    o.V = c.verts + 3 * (int64_t)__ldg(c.vert_off + i);
    guard_empty_range(o.nv, o.V);
    o.base = base;
    o.gpool = c.graph;
    const double2 *T = reinterpret_cast<const double2 *>(c.pose + 16 * i);
    const double *p = c.param + 3 * i;
    if (lane < 6) {
        double2 a = __ldg(T + lane);
        base[2 * lane] = a.x;
        base[2 * lane + 1] = a.y;
    } else if (lane < 9) {
        base[12 + (lane - 6)] = __ldg(p + (lane - 6));
    } else if (lane == 9) {
        base[15] = c.margin ? __ldg(c.margin + i) : R(0.0);
    }
    __syncwarp();
    if (o.type == D3D_MESH && lane == 0) {
        int off, start;
        mesh_graph_of(c, i, off, start);
        base[12] = (real)start;
        base[13] = (real)off;
    }
    return o;
}

// np.dot(pose[:3,:3].T, d)  (dgemv convention)
template <class C>
D3D_DEV v3 rot_t(const C &c, v3 d) {
    return V3(gemv_row(c.r00(), c.r10(), c.r20(), d), gemv_row(c.r01(), c.r11(), c.r21(), d),
              gemv_row(c.r02(), c.r12(), c.r22(), d));
}
// utils.py:143 transform_point
template <class C>
D3D_DEV v3 xform(const C &c, v3 v) {
    return V3(c.tx() + gemv_row(c.r00(), c.r01(), c.r02(), v), c.ty() + gemv_row(c.r10(), c.r11(), c.r12(), v),
              c.tz() + gemv_row(c.r20(), c.r21(), c.r22(), v));
}

// geometry.py:138-157: vertex i of a box, bit k of i selects +0.5 on axis (2-k)
template <class C>
D3D_DEV v3 box_vertex(const C &c, int i) {
    v3 l = V3((i & 4) ? R(0.5) * c.p0() : -R(0.5) * c.p0(), (i & 2) ? R(0.5) * c.p1() : -R(0.5) * c.p1(),
              (i & 1) ? R(0.5) * c.p2() : -R(0.5) * c.p2());
    return V3(c.tx() + dot_blas(l, V3(c.r00(), c.r01(), c.r02())), c.ty() + dot_blas(l, V3(c.r10(), c.r11(), c.r12())),
              c.tz() + dot_blas(l, V3(c.r20(), c.r21(), c.r22())));
}

// Index of the first maximum (MAX) / minimum of one (value, index) candidate per lane over a
// full warp, lowest index on ties - np.argmax / np.argmin semantics - with three integer warp
// reductions (REDUX) on an order-preserving key instead of a five-step shuffle butterfly on
// (double, int, flag).  +0.0 is added first so that -0.0 and +0.0 tie as they do for numpy;
// finite values only.  Lanes with valid == false never win.
template <bool MAX>
D3D_DEV int warp_first_extreme(double val, int idx, bool valid) {
    const unsigned FULL = 0xffffffffu;
    long long b = __double_as_longlong(val + 0.0);
    unsigned long long key = (unsigned long long)b ^ (b < 0 ? 0xffffffffffffffffull : 0x8000000000000000ull);
    if (!MAX) key = ~key;  // minimum of val = maximum of the complemented key
    unsigned hi = valid ? (unsigned)(key >> 32) : 0u, lo = (unsigned)key;
    unsigned mh = __reduce_max_sync(FULL, hi);
    bool c = valid && hi == mh;
    unsigned ml = __reduce_max_sync(FULL, c ? lo : 0u);
    c = c && lo == ml;
    return __reduce_min_sync(FULL, c ? idx : 0x7fffffff);
}

// First arg-max of V.dot(d) (colliders.py:132).  G lanes of a warp cooperate:
// lane `lane` scans vertices lane, lane+G, ...; the reduction prefers the larger
// value and, on ties, the lower index, so the result equals numpy's argmax.
template <int G>
D3D_DEV int argmax_dot(const double *V, int n, v3 d, int lane) {
    if (n == 1) return 0;
    if (G == 1) {
        int bi = 0;
        real best = gemv_row(__ldg(V), __ldg(V + 1), __ldg(V + 2), d);
        for (int i = 1; i < n; ++i) {
            real val = gemv_row(__ldg(V + 3 * i), __ldg(V + 3 * i + 1), __ldg(V + 3 * i + 2), d);
            if (val > best) { best = val; bi = i; }
        }
        return bi;
    }
    real best = R(0.0);
    int bi = 0x7fffffff;
    bool have = false;
    // four rounds per trip: the twelve loads of a lane are in flight together (the scan is
    // bound by the latency of the vertex loads, not by their number)
#pragma unroll 1
    for (int i0 = lane; i0 < n; i0 += 4 * G) {
        real x[4], y[4], z[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = min(i0 + u * G, n - 1);
            x[u] = __ldg(V + 3 * i); y[u] = __ldg(V + 3 * i + 1); z[u] = __ldg(V + 3 * i + 2);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * G;
            real val = gemv_row(x[u], y[u], z[u], d);
            if (i < n && (!have || val > best)) { best = val; bi = i; have = true; }
        }
    }
#ifndef D3D_F32
    if (G == 32) return warp_first_extreme<true>(best, bi, have);
#endif
#pragma unroll 1
    for (int off = G / 2; off > 0; off >>= 1) {
        real ov = __shfl_xor_sync(0xffffffffu, best, off, G);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off, G);
        bool ohave = __shfl_xor_sync(0xffffffffu, (int)have, off, G);
        if (ohave && (!have || ov > best || (ov == best && oi < bi))) {
            best = ov; bi = oi; have = true;
        }
    }
    return bi;
}

// mesh.py:90-139 hill_climb_mesh_extreme: from `start`, first the six axis-extreme shortcut
// vertices, then rounds over the neighbour list of the vertex the round started on; a
// candidate replaces the current best when ddot(d, v_c - v_best) > 10 EPS (mesh.py:9), the
// comparison base moving with every acceptance.  Strictly sequential - in the warp-per-pair
// kernels every lane runs it redundantly (broadcast loads).  The reference has no bound on
// the number of rounds; exact arithmetic needs at most nv, the cap only guards a GPU hang.
static __device__ __noinline__ int hill_climb(const double *V, int nv, const int32_t *g, real lx,
                                              real ly, real lz, int start) {
    const real eps = R(10.0) * D3D_EPS;
    const v3 l = V3(lx, ly, lz);
    int best = start;
    v3 vb = ld3(V + 3 * best);
#pragma unroll 1
    for (int k = 1; k <= 6; ++k) {
        int cidx = __ldg(g + k);
        v3 vc = ld3(V + 3 * cidx);
        if (dot_blas(l, vc - vb) > eps) { best = cidx; vb = vc; }
    }
    bool converged = false;
#pragma unroll 1
    for (int round = 0; !converged && round < 2 * nv + 16; ++round) {
        converged = true;
        const int lo = __ldg(g + 7 + best), hi = __ldg(g + 8 + best);
#pragma unroll 1
        for (int j = lo; j < hi; ++j) {
            int cidx = __ldg(g + j);
            v3 vc = ld3(V + 3 * cidx);
            if (dot_blas(l, vc - vb) > eps) { best = cidx; vb = vc; converged = false; }
        }
    }
    return best;
}

// utils.py:78-122
D3D_DEV void plane_basis(v3 n, v3 &x, v3 &y) {
    if (fabs(n.x) >= fabs(n.y)) {
        real len = sqrt(n.x * n.x + n.z * n.z);
        x = V3(-n.z / len, R(0.0), n.x / len);
        y = V3(n.y * x.z, n.z * x.x - n.x * x.z, -n.y * x.x);
    } else {
        real len = sqrt(n.y * n.y + n.z * n.z);
        x = V3(R(0.0), n.z / len, -n.y / len);
        y = V3(n.y * x.z - n.z * x.y, -n.x * x.z, n.x * x.y);
    }
}

// TM = bit mask of the type tags this instance can meet: cases outside it are compiled out.
// The GJK thread kernel has an instance for batches of analytic primitives only
// (D3D_PRIMITIVE_MASK): the kernel is instruction-fetch bound and unused cases cost run time
// (type-mask sweep of round 1: +6 % on the C1 mix, +13 % for a sphere-only instance).
#define D3D_ALL_TYPES_MASK 0x3ff
#define D3D_PRIMITIVE_MASK 0x1f  // sphere, capsule, box, ellipsoid, cylinder
#define D3D_VERTEX_MASK 0x64     // box, hull, mesh: arg-max over a vertex list
#define D3D_HAS(t) ((TM >> (t)) & 1)

// Types whose support is evaluated in the collider frame share ONE copy of the world->local
// rotation of d and of the local->world transform of the result (code size: the GJK kernels
// are instruction-fetch bound).
#define D3D_LOCAL_FRAME_MASK ((1 << D3D_CAPSULE) | (1 << D3D_CYLINDER) | (1 << D3D_ELLIPSOID) | \
                              (1 << D3D_MESH) | (1 << D3D_CONE))

template <int G, int TM = D3D_ALL_TYPES_MASK, class C>
D3D_DEV v3 support_unmargined(const C &c, v3 d, int lane) {
    const bool local_frame = (D3D_LOCAL_FRAME_MASK >> c.type) & 1;
    v3 l = d;
    if ((TM & D3D_LOCAL_FRAME_MASK) && local_frame) l = rot_t(c, d);
    v3 v = V3(R(0.0), R(0.0), R(0.0));
    switch (c.type) {
    case D3D_SPHERE: if (D3D_HAS(D3D_SPHERE)) {  // geometry.py:341-346
        real s = norm3(d);
        v3 ctr = V3(c.tx(), c.ty(), c.tz());
        if (s == R(0.0)) return ctr + V3(R(0.0), R(0.0), c.p0());
        return ctr + (d / s) * c.p0();
    }
    break;
    case D3D_CAPSULE: if (D3D_HAS(D3D_CAPSULE)) {  // geometry.py:243-256
        real s = dsqrt(l.x * l.x + l.y * l.y + l.z * l.z);
        if (s == R(0.0)) v = V3(c.p0(), R(0.0), R(0.0));
        else v = l * ddiv(c.p0(), s);
        if (l.z > R(0.0)) v.z += R(0.5) * c.p1();
        else v.z -= R(0.5) * c.p1();
    }
    break;
    case D3D_CYLINDER: if (D3D_HAS(D3D_CYLINDER)) {  // geometry.py:194-206
        real s = dsqrt(l.x * l.x + l.y * l.y);
        real z = (l.z < R(0.0)) ? -R(0.5) * c.p1() : R(0.5) * c.p1();
        if (s == R(0.0)) v = V3(c.p0(), R(0.0), z);
        else { real k = ddiv(c.p0(), s); v = V3(l.x * k, l.y * k, z); }
    }
    break;
    case D3D_ELLIPSOID: if (D3D_HAS(D3D_ELLIPSOID)) {  // geometry.py:282-284
        v3 r = V3(c.p0(), c.p1(), c.p2());
        v = vmul(normalized(vmul(l, r)), r);
    }
    break;
    case D3D_BOX: if (D3D_HAS(D3D_BOX)) {  // colliders.py:132 over the 8 vertices of geometry.py:157
        if (G == 1) {
            v3 bestv = box_vertex(c, 0);
            real best = gemv_row(bestv.x, bestv.y, bestv.z, d);
#pragma unroll 1
            for (int i = 1; i < 8; ++i) {
                v3 bv = box_vertex(c, i);
                real val = gemv_row(bv.x, bv.y, bv.z, d);
                if (val > best) { best = val; bestv = bv; }
            }
            return bestv;
        }
        return ld3(c.V + 3 * argmax_dot<G>(c.V, 8, d, lane));
    }
    break;
    case D3D_HULL:  // colliders.py:131-132
        if (D3D_HAS(D3D_HULL)) return ld3(c.V + 3 * argmax_dot<G>(c.V, c.nv, d, lane));
        break;
    case D3D_MESH: if (D3D_HAS(D3D_MESH)) {  // mesh.py:79-87; without a graph mesh.py:182-189
        const int32_t *g = c.graph();
        int idx;
        if (g) {
            idx = hill_climb(c.V, c.nv, g, l.x, l.y, l.z, c.mesh_cur());
            c.set_mesh_cur(idx, lane);
        } else {
            idx = argmax_dot<G>(c.V, c.nv, l, lane);
        }
        v = ld3(c.V + 3 * idx);
    }
    break;
    case D3D_DISK: if (D3D_HAS(D3D_DISK)) {  // geometry.py:375-383
        v3 ctr = V3(c.tx(), c.ty(), c.tz());
        v3 n = V3(c.r02(), c.r12(), c.r22());
        v3 x, y;
        plane_basis(n, x, y);
        v3 pt = V3(dot_blas(x, d), dot_blas(y, d), R(0.0));
        real nrm = norm3(pt);
        if (nrm == R(0.0)) return ctr;
        pt = pt * (c.p0() / nrm);
        return V3(ctr.x + gemv_row(x.x, y.x, n.x, pt), ctr.y + gemv_row(x.y, y.y, n.y, pt),
                  ctr.z + gemv_row(x.z, y.z, n.z, pt));
    }
    break;
    case D3D_ELLIPSE: if (D3D_HAS(D3D_ELLIPSE)) {  // geometry.py:412-414
        v3 a0 = V3(c.r00(), c.r10(), c.r20()), a1 = V3(c.r01(), c.r11(), c.r21());
        real l0 = gemv_row(a0.x, a0.y, a0.z, d), l1 = gemv_row(a1.x, a1.y, a1.z, d);
        real w0 = c.p0() * l0, w1 = c.p1() * l1;
        real nrm = norm_dd(w0, w1, R(0.0));
        if (nrm != R(0.0)) { w0 = w0 / nrm; w1 = w1 / nrm; }
        w0 *= c.p0(); w1 *= c.p1();
        return V3(c.tx() + fma(w1, a1.x, w0 * a0.x), c.ty() + fma(w1, a1.y, w0 * a0.y),
                  c.tz() + fma(w1, a1.z, w0 * a0.z));
    }
    break;
    case D3D_CONE: if (D3D_HAS(D3D_CONE)) {  // geometry.py:443-454
        v3 dp = V3(l.x, l.y, R(0.0));
        real nrm = norm3(dp);
        if (nrm == R(0.0)) dp = V3(R(0.0), R(0.0), R(0.0));
        else dp = dp * (c.p0() / nrm);
        v = (dot_blas(l, dp) >= l.z * c.p1()) ? dp : V3(R(0.0), R(0.0), c.p1());
    }
    break;
    }
    if ((TM & D3D_LOCAL_FRAME_MASK) && local_frame) return xform(c, v);
    return v;
}

template <int G, int TM = D3D_ALL_TYPES_MASK, class C>
D3D_DEV v3 support(const C &c, v3 d, int lane) {
    v3 s = support_unmargined<G, TM>(c, d, lane);
    if (c.margin() != R(0.0)) s = s + normalized(d) * c.margin();  // colliders.py:629-631
    return s;
}

// One shared out-of-line copy of the ten-way support switch for a kernel whose collider records
// live in shared memory (GJK thread / warp kernels, EPA): both colliders of a pair go through it.
template <int G, int STRIDE, int TM>
static __device__ __noinline__ v3 support_call(int type, int nv, const double *V, const real *base,
                                               const int32_t *gpool, real dx, real dy, real dz, int lane) {
    ColliderSmem<STRIDE> c;
    c.type = type; c.nv = nv; c.V = V; c.base = base; c.gpool = gpool;
    return support<G, TM>(c, V3(dx, dy, dz), lane);
}

// Out-of-line instance of the ten-way support switch for kernels that keep their collider
// records in registers / local memory (MPR): one copy of the code instead of one per call
// site (the MPR kernel shrank from 209 KB to a fraction of that; these kernels are
// instruction-fetch bound).
template <int G, int TM = D3D_ALL_TYPES_MASK>
static __device__ __noinline__ v3 support_ni(const Collider &c, real dx, real dy, real dz, int lane) {
    return support<G, TM>(c, V3(dx, dy, dz), lane);
}

// colliders.py center(); hull / mesh means are sequential column sums (np.mean axis 0)
template <class C>
D3D_DEV v3 center_of(const C &c) {
    v3 t = V3(c.tx(), c.ty(), c.tz());
    if (c.type == D3D_HULL || c.type == D3D_MESH) {
        v3 s = V3(R(0.0), R(0.0), R(0.0));
        for (int k = 0; k < c.nv; ++k) s = s + ld3(c.V + 3 * k);
        s = s / (real)c.nv;
        return c.type == D3D_HULL ? s : xform(c, s);
    }
    if (c.type == D3D_CONE)
        return V3(t.x + R(0.5) * c.p1() * c.r02(), t.y + R(0.5) * c.p1() * c.r12(), t.z + R(0.5) * c.p1() * c.r22());
    return t;
}
