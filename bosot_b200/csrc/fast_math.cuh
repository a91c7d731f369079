// Device math shared by the pair kernels.
#pragma once

namespace bosot {

// exp(x) for x <= 0 (the Gram argument -d / l^2): Cody-Waite reduction by ln 2, degree-13 Taylor
// polynomial on |r| <= ln2 / 2 (truncation 4e-18 relative), scaling by 2^n in two exact steps so
// that results in the denormal range underflow gradually without a slow path.  The coefficients
// live in constant memory: as immediates every one of them costs two extra UMOV per use.
static __constant__ double c_exp[16] = {
    1.4426950408889634074,       // [0] log2(e)
    6755399441055744.0,          // [1] 1.5 * 2^52: rounds to nearest integer in the low mantissa bits
    -6.93147180369123816490e-01, // [2] -ln2 (high part, 32 trailing zero bits)
    -1.90821492927058770002e-10, // [3] -ln2 (low part)
    1.6059043836821613e-10,      // [4] 1/13!
    2.08767569878681e-09,        // [5] 1/12!
    2.505210838544172e-08,       // [6] 1/11!
    2.755731922398589e-07,       // [7] 1/10!
    2.7557319223985893e-06,      // [8] 1/9!
    2.48015873015873e-05,        // [9] 1/8!
    1.984126984126984e-04,       // [10] 1/7!
    1.388888888888889e-03,       // [11] 1/6!
    8.333333333333333e-03,       // [12] 1/5!
    4.1666666666666664e-02,      // [13] 1/4!
    1.6666666666666666e-01,      // [14] 1/3!
    -1400.0};                    // [15] clamp: exp(-1400) == 0 in fp64, keeps both scale factors normal

__device__ __forceinline__ double exp_nonpos(double x) {
  x = x < c_exp[15] ? c_exp[15] : x;  // NaN stays NaN
  const double t = fma(x, c_exp[0], c_exp[1]);
  const int n = __double2loint(t);
  const double nf = t - c_exp[1];
  double r = fma(nf, c_exp[2], x);
  r = fma(nf, c_exp[3], r);
  double q = c_exp[4];
#pragma unroll
  for (int k = 5; k <= 14; ++k) q = fma(q, r, c_exp[k]);
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  const int n1 = n >> 1, n2 = n - n1;  // 2^n = 2^n1 * 2^n2, both factors normal for n >= -2044
  const double s1 = __hiloint2double((1023 + n1) << 20, 0);
  const double s2 = __hiloint2double((1023 + n2) << 20, 0);
  return (q * s1) * s2;
}


// 1 / d: hardware seed (about 20 bits) + two Newton steps; NaN for d = 0
__device__ __forceinline__ double fast_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  return fma(x, e, x);
}

// log(y) for finite y >= 1: y = 2^e m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m - 1) / (m + 1),
// |s| <= 0.1716, odd series up to s^21 (truncation 2e-17).  Absolute error about 2e-16 (1 + e): what the
// rational-quadratic kernel needs for (1 + x)^-alpha = exp(-alpha log(1 + x)).  NaN in, NaN out.
static __constant__ double c_log[12] = {
    1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0,
    1.41421356237309504880,   // [10] sqrt(2)
    0.69314718055994530942};  // [11] ln 2
__device__ __forceinline__ double log_ge1(double y) {
  const int hi = __double2hiint(y);
  int e = (hi >> 20) - 1023;
  double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(y));
  if (m > c_log[10]) { m *= 0.5; ++e; }
  const double s = (m - 1.0) * fast_rcp(m + 1.0), s2 = s * s;
  double q = c_log[0];
#pragma unroll
  for (int k = 1; k <= 9; ++k) q = fma(q, s2, c_log[k]);
  q = fma(q, s2, 1.0);
  const double r = fma((double)e, c_log[11], 2.0 * s * q);
  return y == y ? r : y;
}

}  // namespace bosot
