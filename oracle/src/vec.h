/* TEST INFRASTRUCTURE (CPU oracle) -- not part of the product path.
 *
 * 3-vector helpers that pin down the floating-point conventions of the
 * reference's numba/numpy/OpenBLAS code, measured bit-for-bit in the build
 * container (see DESIGN.md "FP conventions", oracle/probe_fp_conventions.py):
 *   - njit scalar expressions and np.cross: no FMA contraction, left to right;
 *   - np.dot(vec3, vec3) (BLAS ddot):   fma(x2,y2, fma(x1,y1, x0*y0));
 *   - matrix*vector, >= 2 rows (dgemv): fma(x2,y2, fma(x0,y0, x1*y1)) per row;
 *   - (8x3)*(3x3) box vertices (dgemm): ddot convention;
 *   - np.linalg.norm inside njit (dnrm2): x87 80-bit accumulate + sqrt.
 * Compile with -ffp-contract=off.
 */
#ifndef D3D_ORACLE_VEC_H
#define D3D_ORACLE_VEC_H
#include <math.h>

#define D3D_EPS 2.220446049250313e-16
#define D3D_MAX_FLOAT 1.7976931348623157e308

typedef struct { double x, y, z; } v3;

static inline v3 V3(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vneg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline v3 vscale(v3 a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline v3 vdiv(v3 a, double s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline v3 vmul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
/* BLAS ddot convention */
static inline double vdot(v3 a, v3 b) {
    return __builtin_fma(a.z, b.z, __builtin_fma(a.y, b.y, a.x * b.x));
}
/* BLAS dgemv row convention (middle product first) */
static inline double gemv_row(double r0, double r1, double r2, v3 x) {
    return __builtin_fma(r2, x.z, __builtin_fma(r0, x.x, r1 * x.y));
}
/* numpy elementwise sum of products: (a0*b0 + a1*b1) + a2*b2 */
static inline double vdot_plain(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* dnrm2 as executed by OpenBLAS on x86-64: x87 extended precision */
static inline double vnorm_blas(v3 a) {
    long double x = a.x, y = a.y, z = a.z;
    return (double)sqrtl(x * x + y * y + z * z);
}
/* utils.py:12-30 norm_vector */
static inline v3 vnormalized(v3 a) {
    double n = vnorm_blas(a);
    if (n == 0.0) return a;
    return vdiv(a, n);
}
static inline v3 vload(const double *p) { return V3(p[0], p[1], p[2]); }
static inline void vstore(double *p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline int veq(v3 a, v3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

/* R^T d with R = pose[:3,:3] (np.dot(pose[:3,:3].T, d), dgemv) */
static inline v3 rot_t_apply(const double *T, v3 d) {
    return V3(gemv_row(T[0], T[4], T[8], d), gemv_row(T[1], T[5], T[9], d),
              gemv_row(T[2], T[6], T[10], d));
}
/* utils.py:143 transform_point: t + R v */
static inline v3 transform_point(const double *T, v3 v) {
    return V3(T[3] + gemv_row(T[0], T[1], T[2], v), T[7] + gemv_row(T[4], T[5], T[6], v),
              T[11] + gemv_row(T[8], T[9], T[10], v));
}
#endif
