// FP64 peak microbenchmark for B200 (sm_100a): measures the issue rate of
//   (1) mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4) with register-resident operands
//   (2) plain DFMA
// so that every "fraction of FP64 peak" in this repo has a MEASURED denominator
// (MEASURED_PEAKS.json carries no FP64 entry).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dmma(double *out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) k_dfma(double *out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char **argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
  int first = 1;
  // DMMA: sweep warps per SM (threads per block, 1 block/SM .. 2 blocks/SM)
  int tpbs[] = {128, 256, 512, 1024};
  for (int t = 0; t < 4; t++) {
    for (int bps = 1; bps <= 2; bps++) {
      int tpb = tpbs[t]; if (tpb * bps > 2048) continue;
      int grid = sms * bps;
      double ms = time_ms([&] { k_dmma<8><<<grid, tpb>>>(out, iters, 1.0, 1.0); }, 5);
      double flops = 2.0 * 256.0 * 8 * iters * (double)(tpb / 32) * grid;
      printf("%s{\"kind\": \"dmma884\", \"nacc\": 8, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}",
             first ? "" : ",\n", tpb / 32 * bps, ms, flops / ms * 1e-9);
      first = 0;
    }
  }
  {
    int tpb = 256, grid = sms;
    double ms = time_ms([&] { k_dmma<32><<<grid, tpb>>>(out, iters, 1.0, 1.0); }, 5);
    double flops = 2.0 * 256.0 * 32 * iters * (double)(tpb / 32) * grid;
    printf(",\n{\"kind\": \"dmma884\", \"nacc\": 32, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}", tpb / 32, ms, flops / ms * 1e-9);
    ms = time_ms([&] { k_dmma<1><<<grid, 128>>>(out, iters, 1.0, 1.0); }, 5);
    // dependent chain: latency in ns per DMMA
    printf(",\n{\"kind\": \"dmma884_dependent_chain\", \"ns_per_mma\": %.3f}", ms * 1e6 / iters);
  }
  for (int t = 0; t < 4; t++) {
    int tpb = tpbs[t]; int grid = sms * (tpb == 1024 ? 2 : 1);
    double ms = time_ms([&] { k_dfma<16><<<grid, tpb>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double flops = 2.0 * 16 * iters * (double)tpb * grid;
    printf(",\n{\"kind\": \"dfma\", \"nacc\": 16, \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}",
           tpb / 32 * (tpb == 1024 ? 2 : 1), ms, flops / ms * 1e-9);
  }
  // sustained DMMA: ~3 s loop
  {
    int tpb = 512, grid = sms * 2;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0; double flops = 0;
    for (; n < 400; n++) { k_dmma<8><<<grid, tpb>>>(out, iters * 4, 1.0, 1.0); flops += 2.0 * 256 * 8 * iters * 4 * (double)(tpb / 32) * grid; }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf(",\n{\"kind\": \"dmma884_sustained\", \"seconds\": %.3f, \"tflops\": %.3f}", ms * 1e-3, flops / ms * 1e-9);
  }
  printf("\n]}\n");
  return 0;
}
