// Does the DMMA.8x8x4 issue rate depend on the operand register pattern?  (a) one A/B register pair
// reused by every instruction (the peak microbenchmark), (b) the GEMM kernel's pattern: 32
// accumulators, 8 distinct "m" fragments x 4 distinct "n" fragments per k-step, (c) the same with the
// fragments re-read from shared memory every k-step like the real main loop.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double *out, int iters, const double *in) {
  __shared__ double sm[16 * 132 * 2];
  for (int i = threadIdx.x; i < 16 * 132 * 2; i += 256) sm[i] = in[i & 255];
  __syncthreads();
  double acc[4][8][2];
  for (int n = 0; n < 4; n++) for (int m = 0; m < 8; m++) acc[n][m][0] = acc[n][m][1] = 0.0;
  double af[8], bf[4];
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, warp = threadIdx.x >> 5;
  const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
  for (int m = 0; m < 8; m++) af[m] = in[m + lane];
  for (int n = 0; n < 4; n++) bf[n] = in[32 + n + lane];
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      if (MODE == 2) {
#pragma unroll
        for (int m = 0; m < 8; m++) af[m] = sm[(kk * 4 + t) * 132 + wm + m * 8 + g];
#pragma unroll
        for (int n = 0; n < 4; n++) bf[n] = sm[16 * 132 + (kk * 4 + t) * 132 + wn + n * 8 + g];
      }
#pragma unroll
      for (int n = 0; n < 4; n++)
#pragma unroll
        for (int m = 0; m < 8; m++) {
          if (MODE == 0) dmma(acc[n][m], af[0], bf[0]);
          else dmma(acc[n][m], bf[n], af[m]);
        }
    }
  }
  double s = 0;
  for (int n = 0; n < 4; n++) for (int m = 0; m < 8; m++) s += acc[n][m][0] + acc[n][m][1];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int MODE> double run(double *out, const double *in, int sms) {
  const int iters = 4096;
  k<MODE><<<sms, 256>>>(out, iters, in); CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < 5; r++) { CK(cudaEventRecord(e0)); k<MODE><<<sms, 256>>>(out, iters, in); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  return 2.0 * 256 * 128 * iters * 8.0 * sms / best * 1e-9;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  double *out, *in; CK(cudaMalloc(&out, 8 * 256 * p.multiProcessorCount)); CK(cudaMalloc(&in, 8 * 512)); CK(cudaMemset(in, 0, 8 * 512));
  printf("{\"same_operands_tflops\": %.2f, ", run<0>(out, in, p.multiProcessorCount));
  printf("\"gemm_register_pattern_tflops\": %.2f, ", run<1>(out, in, p.multiProcessorCount));
  printf("\"gemm_pattern_with_lds_tflops\": %.2f}\n", run<2>(out, in, p.multiProcessorCount));
  return 0;
}
