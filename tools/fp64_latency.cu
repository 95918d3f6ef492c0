// Dependent-issue latencies that bound the panel kernels' serial chains on B200 (one warp, clock64 around N dependent ops):
// DFMA, DMUL, rsqrt(double) (MUFU.RSQ64H + Newton), DMMA.8x8x4 accumulate chain, LDS.64, SHFL (64-bit = 2 x 32).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 2048;
__global__ void k(double *out, long long *cyc, double seed) {
  __shared__ double sm[64];
  const int lane = threadIdx.x;
  sm[lane] = seed + lane; sm[lane + 32] = seed;
  __syncwarp();
  double a = seed, b = 1.0000001, c = 1e-9;
  long long t0, t1;
  // DFMA
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = fma(a, b, c);
  t1 = clock64(); if (lane == 0) cyc[0] = t1 - t0;
  // DMUL
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) a = a * b;
  t1 = clock64(); if (lane == 0) cyc[1] = t1 - t0;
  // rsqrt chain
  double r = fabs(a) + 1.0;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) r = rsqrt(r) + 1.0;
  t1 = clock64(); if (lane == 0) cyc[2] = t1 - t0;
  // raw MUFU.RSQ64H approx chain
  double q = r;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q)); q = y; }
  t1 = clock64(); if (lane == 0) cyc[3] = t1 - t0;
  // DMMA accumulate chain
  double acc[2] = {0.0, 0.0};
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[0]), "+d"(acc[1]) : "d"(b), "d"(c));
  t1 = clock64(); if (lane == 0) cyc[4] = t1 - t0;
  // LDS pointer chase (64-bit)
  int idx = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) { double v = sm[idx & 63]; idx = (int)(v - seed) + (idx & 31); idx &= 63; }
  t1 = clock64(); if (lane == 0) cyc[5] = t1 - t0;
  // SHFL double chain
  double s = a;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; i++) s = __shfl_sync(0xffffffffu, s, (lane + 1) & 31);
  t1 = clock64(); if (lane == 0) cyc[6] = t1 - t0;
  // two independent DMMA chains interleaved (issue rate for one warp)
  double a2[2] = {0, 0}, a3[2] = {0, 0}, a4[2] = {0, 0}, a5[2] = {0, 0};
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N / 4; i++) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(a2[0]), "+d"(a2[1]) : "d"(b), "d"(c));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(a3[0]), "+d"(a3[1]) : "d"(b), "d"(c));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(a4[0]), "+d"(a4[1]) : "d"(b), "d"(c));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(a5[0]), "+d"(a5[1]) : "d"(b), "d"(c));
  }
  t1 = clock64(); if (lane == 0) cyc[7] = t1 - t0;
  // division chain
  double dv = fabs(a) + 2.0;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; i++) dv = 1.0 / dv + 1.5;
  t1 = clock64(); if (lane == 0) cyc[8] = t1 - t0;
  out[lane] = a + r + q + acc[0] + acc[1] + idx + s + a2[0] + a3[1] + a4[0] + a5[1] + dv;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
  for (int rep = 0; rep < 2; rep++) k<<<1, 32>>>(out, cyc, 1.5);
  long long h[16];
  cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
  const char *names[] = {"DFMA", "DMUL", "rsqrt(double)+DADD", "MUFU.RSQ64H", "DMMA.8x8x4 chain", "LDS.64 chase(+cvt)", "SHFL.f64", "DMMA 4 chains (per op)", "1/x + DADD"};
  printf("{");
  for (int i = 0; i < 9; i++) printf("\"%s\": %.1f%s", names[i], (double)h[i] / N, i < 8 ? ", " : "");
  printf("}\n");
  return cudaDeviceSynchronize() != cudaSuccess;
}
