// How fast can ONE warp issue DMMA.8x8x4 when (a) it is alone on the SM, (b) one warp runs on each of the four
// sub-partitions, (c) operands come from shared memory with the next step's loads issued first (the panel kernels' loop)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_issue tools/dmma_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 512;
#define DMMA(acc, a, b) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[0]), "+d"(acc[1]) : "d"(a), "d"(b))
__global__ void k(double *out, long long *cyc, int mode) {
  __shared__ double sm[8][N + 64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = lane; i < N + 64; i += 32) sm[warp][i] = 1.0 + 1e-9 * i;
  __syncthreads();
  double a0[2] = {0, 0}, a1[2] = {0, 0}, a2[2] = {0, 0}, a3[2] = {0, 0};
  const double *p = sm[warp] + lane;
  long long t0 = clock64();
  if (mode == 0) {  // register operands, 4 chains
    double x = 1.5, y = 0.5;
#pragma unroll 4
    for (int i = 0; i < N / 4; i++) { DMMA(a0, x, y); DMMA(a1, x, y); DMMA(a2, x, y); DMMA(a3, x, y); }
  } else if (mode == 1) {  // 2 chains, register operands
    double x = 1.5, y = 0.5;
#pragma unroll 4
    for (int i = 0; i < N / 2; i++) { DMMA(a0, x, y); DMMA(a1, x, y); }
  } else {  // 2 chains, operands from shared memory, next step's loads issued before this step's DMMAs
    double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
#pragma unroll 4
    for (int i = 0; i < N / 2; i++) {
      p += 2;
      const double nx0 = p[0], ny0 = p[1], nx1 = p[2], ny1 = p[3];
      DMMA(a0, x0, y0); DMMA(a1, x1, y1);
      x0 = nx0; y0 = ny0; x1 = nx1; y1 = ny1;
    }
  }
  long long t1 = clock64();
  if (lane == 0) cyc[warp] = t1 - t0;
  out[threadIdx.x] = a0[0] + a1[0] + a2[1] + a3[1];
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64 * 8);
  printf("{");
  for (int mode = 0; mode < 3; mode++)
    for (int warps = 1; warps <= 8; warps *= 2) {
      for (int rep = 0; rep < 2; rep++) k<<<1, 32 * warps>>>(out, cyc, mode);
      long long h[8];
      cudaMemcpy(h, cyc, 8 * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int w = 0; w < warps; w++) mx = h[w] > mx ? h[w] : mx;
      printf("\"mode%d_warps%d\": %.1f, ", mode, warps, (double)mx / N);
    }
  printf("\"unit\": \"cycles per DMMA per warp\"}\n");
  return cudaDeviceSynchronize() != cudaSuccess;
}
