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
// the device's reciprocal estimate is "about 2^-38"; the host stand-in is deliberately worse and
// biased (truncated to 24 bits, then pushed off by RCP_BIAS units of 2^-36) - the result of the
// norm must not depend on it
#ifndef RCP_BIAS
#define RCP_BIAS 0.0
#endif
static inline double rcp_rough(double x) {
    double y = __longlong_as_double(__double_as_longlong(1.0 / x) & ~0x1fffffffLL);  // 24 bits
    y = fma(y, fma(-x, y, 1.0), y);
    return y * (1.0 + RCP_BIAS * 1.4551915228366852e-11);
}
static __attribute__((noinline)) double ddiv(double a, double b) { return a / b; }
static __attribute__((noinline)) double dsqrt(double a) { return sqrt(a); }
#include NORM_SECTION
extern "C" void host_norm(const double *v, int64_t n, int mode, double *out) {
    for (int64_t i = 0; i < n; ++i) {
        double x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
        if (mode == 2) {
            double big = fmax(fabs(x), fmax(fabs(y), fabs(z)));
            out[i] = (big > 1e-140 && big < 1e140) ? norm_x87_core_t<true>(x, y, z) : norm_x87(x, y, z);
        } else if (mode == 1) {  // integer emulation alone, seeded by a plain double (off by up to
            // 2^11 units of the 64-bit mantissa: exercises the bit-by-bit rebuild of the root)
            double big = fmax(fabs(x), fmax(fabs(y), fabs(z)));
            out[i] = (big > 1e-140 && big < 1e140) ? norm_x87_exact(x, y, z, sqrt(x * x + y * y + z * z), 0.0)
                                                   : norm_x87(x, y, z);
        } else {
            out[i] = norm_x87(x, y, z);
        }
    }
}
