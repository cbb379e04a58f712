// ab_math.cuh -- the kernels' own FP64 exp / exp10 / log / log10 / atan / pow.
//
// Why not libdevice: ptxas builds every FP64 literal of CUDA's math functions with two 32-bit moves
// (UMOV / IMAD.MOV), which made ~35 % of the issued instructions of the flux kernels and a 188 KB
// code footprint (instruction-cache misses were the top stall in the round-1 ncu capture).  These
// versions read their coefficients from the constant bank (tools/gen_math_tables.py ->
// ab_math_tables.cuh; one LDCU.128 brings two coefficients), skip the denormal / NaN / huge-argument
// paths the physics never takes, and stay within ~2 ulp of the correctly rounded result -- five
// orders of magnitude inside the 1e-10 parity tolerance (tests/test_math_accuracy.py checks them
// against mpmath on the host, tests/test_gpu_parity.py end to end).
//
// Domain contract (asserted by the callers in ab_device.cuh):
//   dexp/dexp10 : any finite x; results below ~1e-304 flush to 0, x > 700 is not supported
//   dlog/dlog10 : x > 0 and normal
//   datan       : any finite x
//   dpowr       : x >= 0 (0**y = 0 for y > 0)
#pragma once

#ifdef ABM_HOST_TEST
// plain C++ build of the same source for the host-side accuracy tests
#include <cmath>
#include <cstdint>
#include <cstring>
#define ABM_FN static inline
#define ABM_TABLE static const
namespace abm {
static inline int hi_word(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int lo_word(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffff); }
static inline double make_double(int hi, int lo)
{
    int64_t b = ((int64_t)hi << 32) | (uint32_t)lo;
    double x; std::memcpy(&x, &b, 8); return x;
}
static inline double rcp_seed(double y) { return (double)(float)(1.0 / y); }   // ~24-bit seed like MUFU.RCP64H
}
#else
#include <cuda_runtime.h>
#define ABM_FN __device__ __forceinline__
namespace abm {
ABM_FN int hi_word(double x) { return __double2hiint(x); }
ABM_FN int lo_word(double x) { return __double2loint(x); }
ABM_FN double make_double(int hi, int lo) { return __hiloint2double(hi, lo); }
ABM_FN double rcp_seed(double y)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));   // MUFU.RCP64H, rel. error <= 2^-23
    return r;
}
}
#endif

#include "ab_math_tables.cuh"

namespace abm {

// 1/y for normal y: seed + two Newton steps (2^-23 -> 2^-46 -> rounding), no special cases
ABM_FN double fast_rcp(double y)
{
    double r = rcp_seed(y);
    double e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    e = fma(-y, r, 1.0);
    return fma(r, e, r);
}

// p(r) ~ exp(r) on |r| <= ln2/2, then * 2^k through the exponent field
ABM_FN double exp_core(double r, int k)
{
    double p = EXP_C[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = fma(p, r, EXP_C[i]);
    return make_double(hi_word(p) + (k << 20), lo_word(p));
}

ABM_FN double dexp(double x)
{
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to nearest integer
    const double t = fma(x, MATH_K[K_L2E], MAGIC);
    const int k = lo_word(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -MATH_K[K_LN2_HI], x);
    r = fma(kf, -MATH_K[K_LN2_LO], r);
    const double v = exp_core(r, k);
    return (x < -700.) ? 0. : v;
}

ABM_FN double dexp10(double x)
{
    const double MAGIC = 6755399441055744.0;
    const double t = fma(x, MATH_K[K_L2T], MAGIC);
    const int k = lo_word(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -MATH_K[K_LG2_HI], x);
    r = fma(kf, -MATH_K[K_LG2_LO], r);
    const double v = exp_core(r * MATH_K[K_LN10], k);
    return (x < -304.) ? 0. : v;
}

// log(x) = k ln2 + log(m), m in [sqrt(2)/2, sqrt(2)); log(m) = 2 atanh(s), s = (m-1)/(m+1)
ABM_FN double dlog(double x)
{
    int hx = hi_word(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    hx |= (i ^ 0x3ff00000);
    k += (i >> 20);
    const double f = make_double(hx, lo_word(x)) - 1.0;
    const double s = f * fast_rcp(2.0 + f);
    const double z = s * s;
    double p = LOG_C[6];
#pragma unroll
    for (int j = 5; j >= 0; --j) p = fma(p, z, LOG_C[j]);
    const double R = z * p;
    const double hfsq = 0.5 * f * f;
    const double dk = (double)k;
    return dk * MATH_K[K_LN2_HI] - ((hfsq - fma(s, hfsq + R, dk * MATH_K[K_LN2_LO])) - f);
}

ABM_FN double dlog10(double x) { return dlog(x) * MATH_K[K_LOG10E]; }

// atan: |x| <= tan(pi/8): poly; <= tan(3pi/8): pi/4 + atan((x-1)/(x+1)); else pi/2 - atan(1/x)
ABM_FN double datan(double x)
{
    const double ax = fabs(x);
    double num = ax, den = 1.0, bhi = 0., blo = 0.;
    if (ax > 2.414213562373095) {
        num = -1.0; den = ax; bhi = MATH_K[K_PIO2_HI]; blo = MATH_K[K_PIO2_LO];
    } else if (ax > 0.4142135623730950) {
        num = ax - 1.0; den = ax + 1.0; bhi = MATH_K[K_PIO4_HI]; blo = MATH_K[K_PIO4_LO];
    }
    const double t = num * fast_rcp(den);
    const double z = t * t;
    double q = ATAN_C[10];
#pragma unroll
    for (int j = 9; j >= 0; --j) q = fma(q, z, ATAN_C[j]);
    const double a = fma(t * z, q, t);          // atan(t)
    return copysign(bhi + (a + blo), x);
}

// x**y, x >= 0
ABM_FN double dpowr(double x, double y) { return (x > 0.) ? dexp(y * dlog(x)) : 0.; }

}  // namespace abm
