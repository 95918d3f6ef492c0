// Batched FP64 tensor-core (DMMA) tile GEMM for sm_100a.
//
// One CTA computes one 128x128 output tile of one matrix of the batch from a task list
// (blockIdx.x = task, blockIdx.y = batch item).  The contraction streams 16-wide k-chunks of both
// operands through a 4-stage cp.async (LDGSTS) pipeline into padded shared memory; fragments are
// read conflict-free (leading dimensions == 4 mod 16 doubles) and fed to
// mma.sync.aligned.m8n8k4.f64 (SASS DMMA.8x8x4 -- tcgen05 has no FP64 kind, so this IS the FP64
// tensor path on B200; measured peak 37.0 TFLOP/s, profiles/fp64_peak_r01.json).
// 8 warps, warp tile 64(m) x 32(n): 32 DMMAs per 12 shared-memory fragment loads.
//
// The mma is used "transposed" (mma rows <-> n, mma cols <-> m) so that each thread's two
// accumulator values are adjacent in the column-major output: 16-byte global accesses.
//
// This kernel carries every O(N^3) stage of the path:
//   NT  C = C0 - A B^T          left-looking Cholesky block-column update (replaces the inside of
//                               Eigen LLT under cholesky_decompose, fit_hyperparameters.stan:25)
//   NN  T = L21 W11, W21 = -W22 T   recursive triangular inverse (K^-1 for the gradient)
//   TN  G = W^T W  + fused trace epilogue  0.5 tr((a a^T - K^-1) dK/dtheta)  (the reverse sweep
//                               of multi_normal_cholesky -> cholesky_decompose -> cov_exp_quad)
#include <algorithm>

#include "common.cuh"

namespace gpb {

#ifndef GPB_KC
#define GPB_KC 16
#endif
#ifndef GPB_STAGES
#define GPB_STAGES 4
#endif
constexpr int KC = GPB_KC;          // k-chunk per pipeline stage
constexpr int STAGES = GPB_STAGES;  // cp.async stages in flight
#ifndef GPB_WARPS_M
#define GPB_WARPS_M 2
#endif
#ifndef GPB_WARPS_N
#define GPB_WARPS_N 4
#endif
constexpr int WARPS_M = GPB_WARPS_M, WARPS_N = GPB_WARPS_N;   // warp grid over the 128x128 CTA tile
constexpr int WM = TILE / WARPS_M, WN = TILE / WARPS_N;       // warp tile
constexpr int MI = WM / 8, NI = WN / 8;                       // 8x8 mma tiles per warp
constexpr int NTHREADS = 32 * WARPS_M * WARPS_N;
constexpr int LD_MC = TILE + 4;  // operand with the tile index contiguous: stage[KC][132]
constexpr int LD_KC = KC + 4;    // operand with k contiguous:             stage[128][KC+4]
constexpr int STAGE_DOUBLES = (TILE * LD_KC > KC * LD_MC) ? TILE * LD_KC : KC * LD_MC;
constexpr int CPR_K = KC / 2;                       // 16-byte chunks per row of a k-contiguous stage
constexpr int ROWS_PER_R_K = NTHREADS / CPR_K;      // rows covered by one sweep of the CTA (k-contiguous)
constexpr int NCOPY = KC * TILE / 2 / NTHREADS;     // 16-byte copies per thread per operand per stage
static_assert(LD_KC % 16 == 4 && LD_MC % 16 == 4, "conflict-free fragment loads need ld == 4 mod 16");
constexpr int GEMM_SMEM_BYTES = STAGES * 2 * STAGE_DOUBLES * (int)sizeof(double);  // 163840

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c[0]), "+d"(c[1])
      : "d"(a), "d"(b));
}

// scratch of the trace epilogue lives behind the pipeline stages so that the next tile's prefetch can
// already be in flight while the epilogue of the current tile runs
constexpr int EPI_SCRATCH_DOUBLES = 4 * TILE + 3 * (NTHREADS / 32) + 8;
constexpr int GEMM_SMEM_TOTAL = GEMM_SMEM_BYTES + EPI_SCRATCH_DOUBLES * (int)sizeof(double);

template <bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tile_kernel(const GemmParams p, const long long total_work) {
  extern __shared__ __align__(16) double smem[];
  const long long lda = p.A.ld, ldb = p.B.ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp % WARPS_M) * WM, wn = (warp / WARPS_M) * WN;

  // per-thread copy geometry (identical for every tile): 4 x 16B for A and for B per stage; the r-th
  // copy of a thread is `rstep` further along the operand's non-contiguous dimension
  int dstA[NCOPY], dstB[NCOPY];
  long long thrA, thrB;  // this thread's element offset inside an operand tile chunk
  {
    if (!A_KC) { const int k = tid >> 6, m2 = tid & 63; thrA = 2 * m2 + (long long)k * lda; }
    else { const int m = tid / CPR_K, k2 = tid % CPR_K; thrA = 2 * k2 + (long long)m * lda; }
    if (!B_KC) { const int k = tid >> 6, n2 = tid & 63; thrB = 2 * n2 + (long long)k * ldb; }
    else { const int n = tid / CPR_K, k2 = tid % CPR_K; thrB = 2 * k2 + (long long)n * ldb; }
#pragma unroll
    for (int r = 0; r < NCOPY; r++) {
      const int idx = tid + NTHREADS * r;
      dstA[r] = A_KC ? (idx / CPR_K) * LD_KC + 2 * (idx % CPR_K) : (idx >> 6) * LD_MC + 2 * (idx & 63);
      dstB[r] = B_KC ? (idx / CPR_K) * LD_KC + 2 * (idx % CPR_K) : (idx >> 6) * LD_MC + 2 * (idx & 63);
    }
  }
  const long long rstepA = (A_KC ? ROWS_PER_R_K : NTHREADS / 64) * lda, rstepB = (B_KC ? ROWS_PER_R_K : NTHREADS / 64) * ldb;
  const long long stepA = A_KC ? (long long)KC : (long long)KC * lda;
  const long long stepB = B_KC ? (long long)KC : (long long)KC * ldb;

  // state of the tile being computed / prefetched
  const double *srcA = nullptr, *srcB = nullptr;
  auto setup = [&](long long w, TileTask &task, long long &b) {
    task = p.tasks[w % p.ntasks];
    b = w / p.ntasks;
    srcA = p.A.p + b * p.A.stride + task.a_r + (long long)task.a_c * lda + thrA;
    srcB = p.B.p + b * p.B.stride + task.b_r + (long long)task.b_c * ldb + thrB;
  };
  // source addresses are recomputed from constant bases for every chunk: incrementing registers an
  // in-flight LDGSTS still reads costs a long-scoreboard (WAR) stall per chunk
  auto load_stage = [&](int stage, int chunk) {
    double *sA = smem + stage * 2 * STAGE_DOUBLES;
    double *sB = sA + STAGE_DOUBLES;
    const double *a0 = srcA + (long long)chunk * stepA, *b0 = srcB + (long long)chunk * stepB;
#pragma unroll
    for (int r = 0; r < NCOPY; r++) cp_async16(sA + dstA[r], a0 + r * rstepA);
#pragma unroll
    for (int r = 0; r < NCOPY; r++) cp_async16(sB + dstB[r], b0 + r * rstepB);
  };
  auto load_frags = [&](const double *sA, const double *sB, int kk, double (&af)[MI], double (&bf)[NI]) {
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
      af[mi] = A_KC ? sA[(wm + mi * 8 + g) * LD_KC + kk * 4 + t] : sA[(kk * 4 + t) * LD_MC + wm + mi * 8 + g];
#pragma unroll
    for (int ni = 0; ni < NI; ni++)
      bf[ni] = B_KC ? sB[(wn + ni * 8 + g) * LD_KC + kk * 4 + t] : sB[(kk * 4 + t) * LD_MC + wn + ni * 8 + g];
  };
  auto prologue = [&](int nk) {  // all STAGES stages in flight
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      if (s < nk) load_stage(s, s);
      cp_async_commit();
    }
  };

  long long w = blockIdx.x;
  if (w >= total_work) return;
  TileTask task;
  long long b;
  setup(w, task, b);
  prologue(task.k_len / KC);

  double acc[NI][MI][2];
  while (true) {
    const int nk = task.k_len / KC;
#pragma unroll
    for (int ni = 0; ni < NI; ni++)
#pragma unroll
      for (int mi = 0; mi < MI; mi++) acc[ni][mi][0] = acc[ni][mi][1] = 0.0;

    cp_async_wait<STAGES - 1>();
    __syncthreads();
    double af[2][MI], bf[2][NI];
    load_frags(smem, smem + STAGE_DOUBLES, 0, af[0], bf[0]);

    for (int kc = 0; kc < nk; kc++) {
      const double *sA = smem + (kc % STAGES) * 2 * STAGE_DOUBLES;
      const double *sB = sA + STAGE_DOUBLES;
#pragma unroll
      for (int kk = 0; kk < KC / 4; kk++) {
        const int cur = kk & 1, nxt = cur ^ 1;
        if (kk < KC / 4 - 1) {
          load_frags(sA, sB, kk + 1, af[nxt], bf[nxt]);
        } else {
          // chunk transition, hidden behind the DMMAs of this last k-step: everybody has finished
          // reading stage kc (its last fragments are in registers), chunk kc+1 has landed
          cp_async_wait<STAGES - 2>();
          __syncthreads();
          if (kc + STAGES < nk) load_stage(kc % STAGES, kc + STAGES);
          cp_async_commit();
          if (kc + 1 < nk) {
            const double *nA = smem + ((kc + 1) % STAGES) * 2 * STAGE_DOUBLES;
            load_frags(nA, nA + STAGE_DOUBLES, 0, af[nxt], bf[nxt]);
          }
        }
#pragma unroll
        for (int ni = 0; ni < NI; ni++)
#pragma unroll
          for (int mi = 0; mi < MI; mi++) dmma884(acc[ni][mi], bf[cur][ni], af[cur][mi]);
      }
    }

    // ---- tile boundary: every stage is free (the barrier of the last transition ordered all reads
    // of this tile's data); start the next tile's loads, then run this tile's epilogue under them
    const TileTask cur_task = task;
    const long long cur_b = b;
    w += gridDim.x;
    const bool more = w < total_work;
    if (more) {
      setup(w, task, b);
      prologue(task.k_len / KC);
    }

    if (EPI == EPI_AXPBY) {
      double *__restrict__ C = p.C.p + cur_b * p.C.stride;
      const double *__restrict__ C0 = p.C0.p ? p.C0.p + cur_b * p.C0.stride : nullptr;
      const double alpha = p.alpha, beta = p.beta;
#pragma unroll
      for (int ni = 0; ni < NI; ni++) {
        const int n = cur_task.c_c + wn + ni * 8 + g;
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
          const int m = cur_task.c_r + wm + mi * 8 + 2 * t;
          double2 v = make_double2(alpha * acc[ni][mi][0], alpha * acc[ni][mi][1]);
          if (C0) {
            const double2 c0 = *reinterpret_cast<const double2 *>(C0 + m + (long long)n * p.C0.ld);
            v.x = fma(beta, c0.x, v.x);
            v.y = fma(beta, c0.y, v.y);
          }
          *reinterpret_cast<double2 *>(C + m + (long long)n * p.C.ld) = v;
        }
      }
    } else {
      // ---- fused trace epilogue: this tile of G = K^-1 never has to reach HBM ---------------
      double *scr = smem + STAGES * 2 * STAGE_DOUBLES;
      double *xr = scr, *xc = scr + TILE, *ar = scr + 2 * TILE, *ac = scr + 3 * TILE;
      double *red = scr + 4 * TILE;
      const double *x = p.x + cur_b * p.x_stride;
      const double *av = p.avec + cur_b * p.a_stride;
      __syncthreads();  // the previous tile's readers of the scratch are done
      if (tid < TILE) {
        const int i = cur_task.c_r + tid;
        xr[tid] = (i < p.n) ? x[i] : 0.0;
        ar[tid] = (i < p.n) ? av[i] : 0.0;
      } else if (tid < 2 * TILE) {
        const int j = cur_task.c_c + tid - TILE;
        xc[tid - TILE] = (j < p.n) ? x[j] : 0.0;
        ac[tid - TILE] = (j < p.n) ? av[j] : 0.0;
      }
      __syncthreads();
      const double rho = p.theta[cur_b * 3 + 1];
      const double nh = -0.5 / (rho * rho);
      const bool diag_tile = (cur_task.flags & 1) != 0;
      double s_se = 0.0, s_d2 = 0.0, s_tr = 0.0;
      double *__restrict__ C = p.C.p ? p.C.p + cur_b * p.C.stride : nullptr;
#pragma unroll
      for (int ni = 0; ni < NI; ni++) {
        const int nl = wn + ni * 8 + g;
        const int j = cur_task.c_c + nl;
#pragma unroll
        for (int mi = 0; mi < MI; mi++) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int ml = wm + mi * 8 + 2 * t + e;
            const int i = cur_task.c_r + ml;
            const double G = acc[ni][mi][e];
            if (i < p.n && j < p.n) {
              const double d = xr[ml] - xc[nl];
              const double d2 = d * d;
              const double ek = exp(d2 * nh);
              const double M = ar[ml] * ac[nl] - G;
              s_se += M * ek;
              s_d2 += M * ek * d2;
              if (diag_tile && i == j) s_tr += G;
            }
          }
          if (C) {
            const int m = cur_task.c_r + wm + mi * 8 + 2 * t;
            *reinterpret_cast<double2 *>(C + m + (long long)j * p.C.ld) =
                make_double2(acc[ni][mi][0], acc[ni][mi][1]);
          }
        }
      }
      const double wgt = diag_tile ? 1.0 : 2.0;
      s_se *= wgt;
      s_d2 *= wgt;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s_se += __shfl_xor_sync(0xffffffffu, s_se, o);
        s_d2 += __shfl_xor_sync(0xffffffffu, s_d2, o);
        s_tr += __shfl_xor_sync(0xffffffffu, s_tr, o);
      }
      if (lane == 0) {
        red[warp * 3 + 0] = s_se;
        red[warp * 3 + 1] = s_d2;
        red[warp * 3 + 2] = s_tr;
      }
      __syncthreads();
      if (tid == 0) {
        double r0 = 0, r1 = 0, r2 = 0;
        for (int w8 = 0; w8 < NTHREADS / 32; w8++) {
          r0 += red[w8 * 3 + 0];
          r1 += red[w8 * 3 + 1];
          r2 += red[w8 * 3 + 2];
        }
        const long long widx = (w - gridDim.x) % p.ntasks;
        double *o = p.partial + (cur_b * p.ntasks + widx) * 4;
        o[0] = r0;
        o[1] = r1;
        o[2] = r2;
        o[3] = 0.0;
      }
    }
    if (!more) break;
  }
  cp_async_wait<0>();
}

template <bool A_KC, bool B_KC, int EPI>
static int launch_one(Handle *h, const GemmParams &p_in, int ntasks, int batch) {
  static bool configured = false;
  static int num_sms = 0;
  auto kern = gemm_tile_kernel<A_KC, B_KC, EPI>;
  if (!configured) {
    GPB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_TOTAL));
    GPB_CUDA(h, cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, h->device));
    configured = true;
  }
  GemmParams p = p_in;
  p.ntasks = ntasks;
  const long long total = (long long)ntasks * batch;
  const int grid = (int)std::min<long long>(total, num_sms);
  ProfScope ps__(h, PC_GEMM);
  kern<<<grid, NTHREADS, GEMM_SMEM_TOTAL, h->stream>>>(p, total);
  GPB_LAUNCH_CHECK(h);
  return 0;
}

int launch_gemm(Handle *h, GemmLayout layout, GemmEpi epi, const GemmParams &p, int ntasks, int batch) {
  if (ntasks <= 0 || batch <= 0) return 0;
  if (epi == EPI_AXPBY) {
    switch (layout) {
      case LAYOUT_NT: return launch_one<false, false, EPI_AXPBY>(h, p, ntasks, batch);
      case LAYOUT_TN: return launch_one<true, true, EPI_AXPBY>(h, p, ntasks, batch);
      case LAYOUT_NN: return launch_one<false, true, EPI_AXPBY>(h, p, ntasks, batch);
    }
  } else {
    if (layout == LAYOUT_TN) return launch_one<true, true, EPI_TRACE>(h, p, ntasks, batch);
  }
  snprintf(h->err, sizeof(h->err), "launch_gemm: unsupported layout/epilogue %d/%d", (int)layout, (int)epi);
  return -2;
}

}  // namespace gpb
