// dcbf_math.cuh -- lean FP64 elementary functions for the solver kernels.
//
// The CUDA library versions of sincos / atan2 / division carry argument-range machinery the solver never needs (huge
// angles, denormals, infinities) and cost 150-250 executed instructions per call; together they were 20 % of the warp
// kernel's instruction stream (profiles/r02_*).  The arguments here are tame -- headings of a few radians, goal offsets of
// metres, slack gaps between 1e-12 and 1e2 -- so plain Cody-Waite reduction + fdlibm/Cephes kernels are enough:
//   fsincos  : 2-term pi/2 reduction (|a| < 1e5), fdlibm __kernel_sin / __kernel_cos            < 1.5 ulp
//   fatan2   : one division (the octant reduction picks numerator and denominator first),
//              Cephes atan rational P4/Q5 on |t| <= 0.66                                         < 2e-16 absolute
//   frcp/fdiv: MUFU.RCP64H seed + two Newton steps (+ one residual step for the quotient)       <= 1 ulp (normal range)
//   frsqrt   : MUFU.RSQ64H seed + two Newton steps                                              <= 2 ulp (pivots of the Cholesky)
//   flog     : fdlibm log kernel without the special cases (barrier sums of the warp kernels)    < 1 ulp
// The file compiles for the host too (tests/hostsim, tests/test_math_cpu.py checks every function against libm).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define DCBF_MHD __host__ __device__ __forceinline__
#else
#define DCBF_MHD inline
#endif

namespace dcbf {

DCBF_MHD double frcp(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}

DCBF_MHD double fdiv(double a, double b) {
#if defined(__CUDA_ARCH__)
    const double r = frcp(b);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
#else
    return a / b;
#endif
}

// 1 / sqrt(x) for normal positive x: MUFU.RSQ64H seed + two Newton steps
DCBF_MHD double frsqrt(double x) {
#if defined(__CUDA_ARCH__)
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
    r = fma(r, fma(-hx * r, r, 0.5), r);   // r (1.5 - 0.5 x r^2)
    return fma(r, fma(-hx * r, r, 0.5), r);
#else
    return 1.0 / sqrt(x);
#endif
}

// natural logarithm of a normal positive x (fdlibm __ieee754_log without its special cases: the arguments are products of slack gaps,
// 1e-40 ... 1e6; a non-positive or non-finite argument returns a meaningless finite number -- the callers reject such trial points on
// the sign of the gaps).  x = 2^k m, m in [sqrt(1/2), sqrt(2)), log m = 2 s + s R(s^2), s = f / (2 + f), f = m - 1;  < 1 ulp
DCBF_MHD double flog(double x) {
    int64_t bits;
#if defined(__CUDA_ARCH__)
    bits = __double_as_longlong(x);
#else
    memcpy(&bits, &x, sizeof(bits));
#endif
    int hx = (int)(bits >> 32);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;                  // mantissa >= sqrt(2): use m / 2 and k + 1
    k += i >> 20;
    bits = ((int64_t)(hx | (i ^ 0x3ff00000)) << 32) | (bits & 0xffffffffll);
    double m;
#if defined(__CUDA_ARCH__)
    m = __longlong_as_double(bits);
#else
    memcpy(&m, &bits, sizeof(m));
#endif
    const double f = m - 1.0, dk = (double)k;
    const double s = fdiv(f, 2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01), 3.999999999940941908e-01);
    const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01, 1.818357216161805012e-01), 2.857142874366239149e-01), 6.666666666666735130e-01);
    const double R = t2 + t1, hfsq = 0.5 * f * f;
    return dk * 6.93147180369123816490e-01 - ((hfsq - fma(s, hfsq + R, dk * 1.90821492927058770002e-10)) - f);
}

// sin and cos of a (|a| < 1e5)
DCBF_MHD void fsincos(double a, double *sn, double *cs) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: round to nearest integer in the low word
    const double t = fma(a, 0.63661977236758134308, MAGIC);   // a * 2/pi
    int64_t bits;
#if defined(__CUDA_ARCH__)
    bits = __double_as_longlong(t);
#else
    memcpy(&bits, &t, sizeof(bits));
#endif
    const int k = (int)(uint32_t)bits;
    const double q = t - MAGIC;
    double r = fma(-q, 1.57079632679489655800e+00, a);        // pi/2, high part
    r = fma(-q, 6.12323399573676603587e-17, r);               // pi/2, low part
    const double z = r * r;
    // fdlibm __kernel_sin / __kernel_cos on |r| <= pi/4
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double s = fma(r * z, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double s1 = (k & 1) ? c : s, c1 = (k & 1) ? s : c;
    *sn = (k & 2) ? -s1 : s1;
    *cs = ((k + 1) & 2) ? -c1 : c1;
}

// atan2(y, x) for finite arguments, not both zero
DCBF_MHD double fatan2(double y, double x) {
    const double ax = fabs(x), ay = fabs(y);
    const double mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
    // atan(mn / mx) in [0, pi/4]; above tan(pi/8)-ish use atan(t) = pi/4 + atan((t - 1) / (t + 1)) -- one division either way
    const bool hi = mn > 0.66 * mx;
    const double num = hi ? mn - mx : mn, den = hi ? mn + mx : mx;
    const double t = den > 0.0 ? fdiv(num, den) : 0.0;   // atan2(0, 0) = 0 like libm / numpy
    const double z = t * t;
    // Cephes atan: z * P4(z) / Q5(z)
    double p = fma(z, -8.750608600031904122785e-1, -1.615753718733365076637e1);
    p = fma(z, p, -7.500855792314704667340e1);
    p = fma(z, p, -1.228866684490136173410e2);
    p = fma(z, p, -6.485021904942025371773e1);
    double q = z + 2.485846490142306297962e1;
    q = fma(z, q, 1.650270098316988542046e2);
    q = fma(z, q, 4.328810604912902668951e2);
    q = fma(z, q, 4.853903996359136964868e2);
    q = fma(z, q, 1.945506571482613964425e2);
    const double w = fdiv(z * p, q);
    double r = fma(t, w, t);
    if (hi) r = (r + 3.061616997868383017329e-17) + 7.85398163397448309616e-1;
    if (ay > ax) r = 1.57079632679489661923 - r;
    if (x < 0.0) r = 3.14159265358979323846 - r;
    return y < 0.0 ? -r : r;
}

}  // namespace dcbf
