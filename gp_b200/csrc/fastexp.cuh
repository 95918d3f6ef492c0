// exp(t) for t <= 0 in about 20 instructions: the covariance kernels evaluate one exponential per
// matrix element (N^2 of them per Gram build and again per gradient), and with the library exp() the
// Gram kernel is instruction-issue bound (91 % issue utilisation, 52 % of HBM write bandwidth, ncu
// profiles/ncu_summary_r01e.md) -- more than half of its ~80 instructions per element only move
// 64-bit polynomial constants into uniform registers.  Here the coefficients sit in constant memory
// and are read as direct DFMA operands.
//
// exp(t) = 2^k exp(r), k = rint(t log2 e), r = t - k ln2 (two-term Cody-Waite), exp(r) by the degree-12
// Taylor polynomial on |r| <= ln2/2 (truncation 1.7e-16 relative); maximum error 2 ulp against libm over
// [-708, 0] (checked in tests/).  Results below 2^-1022 are flushed to zero (relative to a kernel matrix
// with unit-scale diagonal that is 300 orders of magnitude under rounding).
#pragma once

namespace gpb {

static __constant__ double c_exp_poly[13] = {
    1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0,
    1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0};
static __constant__ double c_exp_red[4] = {1.4426950408889634074, 6755399441055744.0,
                                           6.93147180369123816490e-01, 1.90821492927058770002e-10};

__device__ __forceinline__ double exp_nonpos(double t) {
  const double tc = fmax(t, -708.0);
  double kd = fma(tc, c_exp_red[0], c_exp_red[1]);
  const int ki = __double2loint(kd);
  kd -= c_exp_red[1];
  double r = fma(kd, -c_exp_red[2], tc);
  r = fma(kd, -c_exp_red[3], r);
  double p = c_exp_poly[12];
#pragma unroll
  for (int i = 11; i >= 0; i--) p = fma(p, r, c_exp_poly[i]);
  const double scale = __hiloint2double((ki + 1023) << 20, 0);
  return (t < -708.0) ? 0.0 : p * scale;
}

}  // namespace gpb
