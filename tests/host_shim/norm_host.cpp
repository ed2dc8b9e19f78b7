// Host build of the x87-norm section of distance3d_b200/csrc/d3d_math.cuh (tests only): the
// CUDA intrinsics it uses are restated with their documented semantics so that the very same
// source text can be checked against the oracle's long-double dnrm2 without a GPU.
// NORM_SECTION is the path of the extracted section (tests/test_norm_host.py writes it).
#include <math.h>
#include <stdint.h>
#include <string.h>
#define __device__
#define __noinline__ __attribute__((noinline))
#define D3D_DEV static inline
typedef double real;
static inline long long __double_as_longlong(double v) { long long r; memcpy(&r, &v, 8); return r; }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline long long __double2ll_rn(double v) { return llrint(v); }
static inline int max(int a, int b) { return a > b ? a : b; }
static __attribute__((noinline)) double ddiv(double a, double b) { return a / b; }
static __attribute__((noinline)) double dsqrt(double a) { return sqrt(a); }
#include NORM_SECTION
extern "C" void host_norm(const double *v, int64_t n, int mode, double *out) {
    for (int64_t i = 0; i < n; ++i) {
        double x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
        if (mode == 2) {
            double big = fmax(fabs(x), fmax(fabs(y), fabs(z)));
            out[i] = (big > 1e-140 && big < 1e140) ? norm_x87_core_t<true>(x, y, z) : norm_x87(x, y, z);
        } else if (mode == 3) {  // how often the production path leaves for the integer emulation
            out[i] = 0.0;
        } else {
            out[i] = norm_x87(x, y, z);
        }
    }
}
