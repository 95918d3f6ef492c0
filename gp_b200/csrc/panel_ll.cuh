// Device helpers shared by the shared-memory (left-looking) panel kernels: panel.cu and small.cu.
//
// Latencies that shape them, measured on B200 (tools/fp64_latency.cu): DFMA / DMUL 8.7 / 8.2 cycles dependent issue,
// MUFU.RSQ64H 17, rsqrt(double) 52, DMMA.8x8x4 26 dependent and one per 16 cycles per sub-partition, LDS ~30,
// SHFL.f64 26.  A 128-wide Cholesky panel is a serial chain of 128 pivots; everything here is about keeping that
// chain short (two pivots per reciprocal square root) and about keeping loads off it (operands of the next DMMA
// step are fetched before the current DMMAs issue: a kernel with one or two warps per sub-partition has nobody else
// to hide a 30-cycle shared-memory load behind).
#pragma once
#include "common.cuh"

namespace gpb {

__device__ __forceinline__ void dmma884v(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// 1 / sqrt(x) for positive normal x: MUFU.RSQ64H seed (2^-22) and one third-order correction -- the library's
// sequence without its special-case branch (a non-positive pivot is reported through `bad`, its garbage never used).
__device__ __forceinline__ double rsqrt_pos(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);
  const double p = fma(e, 0.375, 0.5);
  return fma(p * e, y, y);
}

// Cholesky of an 8x8 block held (lower part) in registers, every lane the same values.  Two columns per step:
// for the leading 2x2 block [a b; b c] the second pivot is (ac - b^2) / a, so rsqrt(a) and rsqrt(ac - b^2) do not
// wait for each other and the dependent chain per PAIR of columns is one rsqrt, not two (94 cycles against 160).
// inv[k] = 1 / L[k][k]; bad = first k with a non-positive (or NaN) pivot, -1 if none.
__device__ __forceinline__ void factor8_pairs(double (&d)[8][8], double (&inv)[8], int &bad) {
  bad = -1;
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const double a = d[k][k], b = d[k + 1][k], c = d[k + 1][k + 1];
    const double det = fma(a, c, -(b * b));
    if (bad < 0) {
      if (!(a > 0.0)) bad = k;
      else if (!(det > 0.0)) bad = k + 1;
    }
    const double r1 = rsqrt_pos(a), rd = rsqrt_pos(det);
    const double s1 = a * r1;          // sqrt(a)
    const double i2 = rd * s1;         // 1 / L[k+1][k+1]
    const double l10 = b * r1;
    inv[k] = r1;
    inv[k + 1] = i2;
    d[k][k] = s1;
    d[k + 1][k] = l10;
    d[k + 1][k + 1] = (det * rd) * r1; // sqrt(det / a)
#pragma unroll
    for (int i = k + 2; i < 8; i++) {
      const double l0 = d[i][k] * r1;
      d[i][k] = l0;
      d[i][k + 1] = fma(-l0, l10, d[i][k + 1]) * i2;
    }
#pragma unroll
    for (int j = k + 2; j < 8; j++)
#pragma unroll
      for (int i = j; i < 8; i++) d[i][j] = fma(-d[i][k + 1], d[j][k + 1], fma(-d[i][k], d[j][k], d[i][j]));
  }
}

// x <- x D^-T for one row x[0..8) against an 8x8 lower-triangular block given ROW-MAJOR (dd[c * 8 + cp], cp <= c)
// with its reciprocal diagonal: true substitution (backward stable), 16 dependent FP64 operations.
__device__ __forceinline__ void solve_row8(double (&x)[8], const double (&dl)[8][8], const double (&inv)[8]) {
#pragma unroll
  for (int c = 0; c < 8; c++) {
    double sv = x[c];
#pragma unroll
    for (int cp = 0; cp < c; cp++) sv = fma(-x[cp], dl[c][cp], sv);
    x[c] = sv * inv[c];
  }
}

// loads the lower part of a row-major 8x8 block and its reciprocal diagonal with 16-byte shared-memory loads
__device__ __forceinline__ void load_block8(const double *dd, const double *iv, double (&dl)[8][8], double (&inv)[8]) {
#pragma unroll
  for (int c = 1; c < 8; c++)
#pragma unroll
    for (int cp = 0; cp < c; cp += 2) {
      const double2 v = *reinterpret_cast<const double2 *>(dd + c * 8 + cp);
      dl[c][cp] = v.x;
      if (cp + 1 < c) dl[c][cp + 1] = v.y;
    }
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    const double2 v = *reinterpret_cast<const double2 *>(iv + c);
    inv[c] = v.x;
    inv[c + 1] = v.y;
  }
}

// s[mt][ch] += sum over nkb 8-wide k-blocks of A(rows of m-tile mt) * B^T with the fragments of step kb + 1 loaded
// before the DMMAs of step kb issue.  pa / pb point at this lane's element of k-block 0 (row-tile 0 for A);
// a_kb / b_kb = distance between k-blocks, a_ch / b_ch = distance between the two 4-wide halves of a k-block,
// a_mt = distance between the m-tiles of A (all in doubles).  act[mt] switches an m-tile off (warp-uniform).
template <int MT>
__device__ __forceinline__ void ll_accumulate(double (&s)[MT][2][2], const double *pa, const double *pb, int nkb, int a_kb,
                                              int b_kb, int a_ch, int b_ch, int a_mt, const bool (&act)[MT]) {
  if (nkb <= 0) return;
  double a[MT][2], b[2];
#pragma unroll
  for (int ch = 0; ch < 2; ch++) {
    b[ch] = pb[ch * b_ch];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) a[mt][ch] = pa[ch * a_ch + mt * a_mt];
  }
  for (int kb = 0; kb < nkb; kb++) {
    double an[MT][2] = {}, bn[2] = {};
    pa += a_kb;
    pb += b_kb;
    if (kb + 1 < nkb) {
#pragma unroll
      for (int ch = 0; ch < 2; ch++) {
        bn[ch] = pb[ch * b_ch];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) an[mt][ch] = pa[ch * a_ch + mt * a_mt];
      }
    }
#pragma unroll
    for (int ch = 0; ch < 2; ch++)
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
        if (act[mt]) dmma884v(s[mt][ch], a[mt][ch], b[ch]);
#pragma unroll
    for (int ch = 0; ch < 2; ch++) {
      b[ch] = bn[ch];
#pragma unroll
      for (int mt = 0; mt < MT; mt++) a[mt][ch] = an[mt][ch];
    }
  }
}

}  // namespace gpb
