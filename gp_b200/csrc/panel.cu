// Panel kernels of the tiled FP64 Cholesky (sm_100a): the latency-critical 128x128 pieces that sit
// between the DMMA GEMM launches.  They replace the unblocked inner part of Eigen's LLT that runs
// under Stan Math's cholesky_decompose (models/fit_hyperparameters.stan:25, covariance.cpp:29).
//
//   potrf_tile_kernel   in-register right-looking Cholesky of one 128x128 diagonal tile per CTA
//                       (2-D cyclic ownership, one barrier per column, pivot check -> info)
//   trsm_tile_kernel    X = C L^-T for a 128x128 tile below the diagonal: true substitution on 8x8
//                       diagonal blocks with quad shuffles, DMMA (mma.sync m8n8k4 f64) for the
//                       rank-8 updates.  The same kernel with C = I and a transposed store
//                       produces the inverse of a diagonal tile (seed of the recursive TRTRI).
#include "common.cuh"

namespace gpb {

__device__ __forceinline__ void dmma884p(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------
// POTRF of one 128x128 tile, blocked (v2): right-looking over sixteen 8-column blocks with the tile
// register-resident in mma accumulator layout (warp w owns rows 16w..16w+15, like the TRSM kernel):
//   1. the warp that owns the 8x8 diagonal block factors it with warp shuffles (8 dependent steps)
//   2. every warp solves its rows of the 8-column panel against that block (quad shuffles)
//   3. rank-8 SYRK update of the trailing lower triangle with DMMA (mma.sync m8n8k4 f64); the panel
//      goes through shared memory once per block (B operand), the A operand comes from registers
// Two block barriers per 8 columns instead of one per column, and the O(n^3) part on the tensor pipe.
// ------------------------------------------------------------------------------------------------
constexpr int LD_P = 12;  // panel row stride in doubles: conflict-free B-operand reads ((12g + t) mod 16 distinct)

__global__ void __launch_bounds__(256, 1)
potrf_tile_kernel_v2(double *__restrict__ Lbase, long long ld, long long stride, long long diag_off, int index_base,
                     int n, int *info) {
  __shared__ __align__(16) double Dsm[2][64];         // Dsm[c*8 + r] = D[r][c] (factored diagonal block)
  __shared__ double dinv[2][8];
  __shared__ __align__(16) double Psm[2][TILE * LD_P];  // solved panel rows: Psm[row*LD_P + c]
  __shared__ int s_info;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const unsigned FULL = 0xffffffffu;
  const int qbase = lane & ~3;
  double *T = Lbase + (long long)blockIdx.x * stride + diag_off;
  const int r0 = warp * 16;
  if (tid == 0) s_info = 0;

  double acc[2][16][2];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < 16; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int r = r0 + mt * 8 + g, c = nt * 8 + 2 * t + e;
        acc[mt][nt][e] = (nt <= 2 * warp + mt) ? T[r + (long long)c * ld] : 0.0;  // tiles right of the diagonal are never used
      }
  __syncthreads();

#pragma unroll
  for (int cb = 0; cb < 16; cb++) {
    const int buf = cb & 1;
    // ---- 1. diagonal 8x8 block: rows 8cb..8cb+7 live in warp cb/2, m-tile cb%2 ------------------
    if (warp == (cb >> 1)) {
#pragma unroll
      for (int mo = 0; mo < 2; mo++) {
        if (mo == (cb & 1)) {
          double a0 = acc[mo][cb][0], a1 = acc[mo][cb][1];  // element (row g, cols 2t, 2t+1)
#pragma unroll
          for (int k = 0; k < 8; k++) {
            const double ak = (k & 1) ? a1 : a0;               // this lane's element in column k (valid if t == k/2)
            const double piv = __shfl_sync(FULL, ak, 4 * k + (k >> 1));
            if (!(piv > 0.0) && lane == 0 && s_info == 0) s_info = index_base + cb * 8 + k + 1;
            const double inv = rsqrt(piv), dgl = piv * inv;  // shortest dependent chain on the critical path
            const double lrow = __shfl_sync(FULL, ak, 4 * g + (k >> 1)) * inv;            // l[g][k]
            const double lc0 = __shfl_sync(FULL, ak, 4 * (2 * t) + (k >> 1)) * inv;       // l[2t][k]
            const double lc1 = __shfl_sync(FULL, ak, 4 * (2 * t + 1) + (k >> 1)) * inv;   // l[2t+1][k]
            if (2 * t > k && g >= 2 * t) a0 = fma(-lrow, lc0, a0);
            if (2 * t + 1 > k && g >= 2 * t + 1) a1 = fma(-lrow, lc1, a1);
            if (t == (k >> 1)) {  // write the finished column k
              const double v = (g > k) ? lrow : ((g == k) ? dgl : 0.0);
              if (k & 1) a1 = v; else a0 = v;
            }
          }
          // strict upper part of the block -> exact zeros
          if (2 * t > g) a0 = 0.0;
          if (2 * t + 1 > g) a1 = 0.0;
          acc[mo][cb][0] = a0;
          acc[mo][cb][1] = a1;
          Dsm[buf][(2 * t) * 8 + g] = a0;
          Dsm[buf][(2 * t + 1) * 8 + g] = a1;
          if (g == 2 * t) dinv[buf][g] = 1.0 / a0;
          if (g == 2 * t + 1) dinv[buf][g] = 1.0 / a1;
        }
      }
    }
    __syncthreads();
    // ---- 2. panel solve for the m-tiles strictly below the diagonal block ------------------------
    double af[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
      const bool below = (2 * warp + mt) > cb;  // warp-uniform
      if (below) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
          double xv = acc[mt][cb][c & 1] * dinv[buf][c];
          if (t == (c >> 1)) acc[mt][cb][c & 1] = xv;
          xv = __shfl_sync(FULL, xv, qbase + (c >> 1));
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int cp = 2 * t + e;
            if (cp > c) acc[mt][cb][e] = fma(-xv, Dsm[buf][c * 8 + cp], acc[mt][cb][e]);
          }
        }
        *reinterpret_cast<double2 *>(&Psm[buf][(r0 + mt * 8 + g) * LD_P + 2 * t]) =
            make_double2(acc[mt][cb][0], acc[mt][cb][1]);
      }
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
        const int src = qbase + 2 * ks + (t >> 1);
        const double v0 = __shfl_sync(FULL, acc[mt][cb][0], src);
        const double v1 = __shfl_sync(FULL, acc[mt][cb][1], src);
        af[mt][ks] = (t & 1) ? v1 : v0;
      }
    }
    __syncthreads();
    // ---- 3. trailing update, lower triangle only: C[i][j] -= X[i] . X[j], 8cb+7 < j <= i ---------
#pragma unroll
    for (int nt = 0; nt < 16; nt++) {
      if (nt > cb && nt <= 2 * warp + 1) {
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          const double bfv = -Psm[buf][(nt * 8 + g) * LD_P + ks * 4 + t];
          if (nt <= 2 * warp) dmma884p(acc[0][nt], af[0][ks], bfv);
          dmma884p(acc[1][nt], af[1][ks], bfv);
        }
      }
    }
  }

  // write back: lower triangle = L, strict upper = 0 (Eigen matrixL() convention)
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < 16; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int r = r0 + mt * 8 + g, c = nt * 8 + 2 * t + e;
        T[r + (long long)c * ld] = (r >= c) ? acc[mt][nt][e] : 0.0;
      }
  __syncthreads();
  if (tid == 0 && s_info != 0 && s_info <= n) {
    if (info[blockIdx.x] == 0) info[blockIdx.x] = s_info;
  }
}

// ------------------------------------------------------------------------------------------------
// TRSM tile: X L^T = C with L the 128x128 lower-triangular diagonal tile, C a 128x128 tile.
// 8 warps; warp w owns rows 16w..16w+15 as two m8 mma row tiles x sixteen n8 column tiles.
// MODE 0: C read from / X written to the tile (in place).  MODE 1: C = I, X^T written to Wout
// (Wout = L^-1, lower triangular, strict upper zero).
// ------------------------------------------------------------------------------------------------
constexpr int LD_L = TILE + 4;
constexpr int TRSM_SMEM_BYTES = (TILE * LD_L + TILE) * (int)sizeof(double);

template <int MODE>
__global__ void __launch_bounds__(256, 1)
trsm_tile_kernel(const double *Ldiag_base, double *Cbase, long long ld, long long stride, long long diag_off,
                 long long diag_step, long long c_off, long long c_step) {
  extern __shared__ __align__(16) double sm[];
  double *Ls = sm;                     // Ls[k*LD_L + n] = L[n][k]
  double *invd = sm + TILE * LD_L;     // 1 / L[n][n]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int item = blockIdx.y;
  const double *Ld = Ldiag_base + (long long)item * stride + diag_off + (long long)blockIdx.x * diag_step;
  double *Ct = Cbase + (long long)item * stride + c_off + (long long)blockIdx.x * c_step;

  // stage the diagonal tile with cp.async (column-major copy, 16B vectors): all 32 copies of a thread are
  // in flight at once, together with the C-tile loads below (both tiles are 128 KB and this kernel runs
  // one CTA per SM, so exposed load latency is what bounds it)
  for (int idx = tid; idx < TILE * TILE / 2; idx += 256) {
    const int k = idx >> 6, n2 = idx & 63;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(Ls + k * LD_L + 2 * n2);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(Ld + 2 * n2 + (long long)k * ld) : "memory");
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");

  // the C tile goes straight to registers
  double acc[2][16][2];
  const int r0 = warp * 16;
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < 16; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int r = r0 + mt * 8 + g, c = nt * 8 + 2 * t + e;
        acc[mt][nt][e] = (MODE == 0) ? Ct[r + (long long)c * ld] : ((r == c) ? 1.0 : 0.0);
      }

  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  if (tid < TILE) invd[tid] = 1.0 / Ls[tid * LD_L + tid];
  __syncthreads();

  const unsigned FULL = 0xffffffffu;
  const int qbase = lane & ~3;
#pragma unroll
  for (int cb = 0; cb < 16; cb++) {
    // (i) substitution inside the 8x8 diagonal block; a row's 8 values live in one quad
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int owner = qbase + (c >> 1);
      const int n = cb * 8 + c;
      const double dinv = invd[n];
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        double xv = acc[mt][cb][c & 1] * dinv;
        if (t == (c >> 1)) acc[mt][cb][c & 1] = xv;
        xv = __shfl_sync(FULL, xv, owner);
        // remaining columns c' > c of this block held by this lane: c' = 2t, 2t+1
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int cp = 2 * t + e;
          if (cp > c) acc[mt][cb][e] = fma(-xv, Ls[n * LD_L + cb * 8 + cp], acc[mt][cb][e]);
        }
      }
    }
    // (ii) re-layout the solved 8 columns from accumulator layout to mma A-operand layout
    double af[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
        const int src = qbase + 2 * ks + (t >> 1);
        const double v0 = __shfl_sync(FULL, acc[mt][cb][0], src);
        const double v1 = __shfl_sync(FULL, acc[mt][cb][1], src);
        af[mt][ks] = (t & 1) ? v1 : v0;
      }
    // (iii) rank-8 update of the columns to the right:  C[:, n] -= X[:, cb] * L[n, cb]^T
#pragma unroll
    for (int nt = 0; nt < 16; nt++) {
      if (nt > cb) {
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          const double bfv = -Ls[(cb * 8 + ks * 4 + t) * LD_L + nt * 8 + g];
#pragma unroll
          for (int mt = 0; mt < 2; mt++) dmma884p(acc[mt][nt], af[mt][ks], bfv);
        }
      }
    }
  }

#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nt = 0; nt < 16; nt++) {
      const int r = r0 + mt * 8 + g, c = nt * 8 + 2 * t;
      if (MODE == 0) {
        Ct[r + (long long)c * ld] = acc[mt][nt][0];
        Ct[r + (long long)(c + 1) * ld] = acc[mt][nt][1];
      } else {
        // W[c][r] = X[r][c]; X = L^-T is upper triangular (r <= c); emit exact zeros elsewhere
        const double v0 = (r <= c) ? acc[mt][nt][0] : 0.0;
        const double v1 = (r <= c + 1) ? acc[mt][nt][1] : 0.0;
        *reinterpret_cast<double2 *>(Ct + c + (long long)r * ld) = make_double2(v0, v1);
      }
    }
}

// ------------------------------------------------------------------------------------------------
// Pipelined TRSM over several row tiles of one block column (MODE 0 semantics).  Timing the kernel above
// with its arithmetic removed showed 7.5 of its 18.1 ms per N=4096 step are exposed global-memory time
// (one CTA per SM: load 256 KB, compute, store 128 KB, nothing overlaps).  Here one CTA keeps the diagonal
// tile in shared memory for up to `tiles_per_cta` row tiles -- packed (only the lower triangle, in groups of
// 8 columns: 72 KB instead of 135 KB) so that a full 128x128 prefetch buffer fits next to it -- and the
// cp.async prefetch of the next C tile runs under the substitution + DMMA work of the current one.
// The arithmetic per element is that of trsm_tile_kernel<0>.  What is left (14.4 of 17.3 ms with loads and stores
// removed) is the substitution chain itself: its DMUL/DFMA steps queue behind the other warps' DMMAs on the one
// FP64 pipe (ncu: 44.6 % DMMA-active, DFMA stalls 70 % short-scoreboard + 23 % math-pipe); 16 warps of 8 rows
// instead of 8 warps of 16 rows measured the same.
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int lpk_ld(int cb) { return TILE + 4 - 8 * cb; }   // rows 8cb..127 (+4 pad): == 4 or 12 mod 16
__host__ __device__ constexpr int lpk_off(int cb) { return 8 * (cb * (TILE + 4) - 4 * cb * (cb - 1)); }  // sum_{c<cb} 8*lpk_ld(c)
constexpr int LPK_DOUBLES = lpk_off(16);          // 9216
constexpr int LD_CB = TILE + 2;                   // prefetch buffer: column-major, 130 (conflict-free 8-byte reads)
constexpr int TRSM_PIPE_SMEM_BYTES = (LPK_DOUBLES + TILE + TILE * LD_CB) * (int)sizeof(double);

template <int MT>  // m8 row tiles per warp: 128 / (8 MT) warps per CTA
__global__ void __launch_bounds__(32 * (16 / MT), 1)
trsm_tiles_pipelined_kernel(const double *Ldiag_base, double *Cbase, long long ld, long long stride, long long diag_off,
                            long long c_off, int ntiles, int tiles_per_cta) {
  extern __shared__ __align__(16) double sm[];
  double *Lp = sm;                      // packed: Lp[lpk_off(cb) + kk*lpk_ld(cb) + (n - 8cb)] = L[n][8cb + kk], n >= 8cb
  double *invd = sm + LPK_DOUBLES;      // 1 / L[n][n]
  double *Cb = invd + TILE;             // Cb[c*LD_CB + r] = C[r][c] of the tile being fetched
  constexpr int NTH = 32 * (16 / MT);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int item = blockIdx.y;
  const double *Ld = Ldiag_base + (long long)item * stride + diag_off;
  const int t0 = blockIdx.x * tiles_per_cta;
  const int t1 = min(ntiles, t0 + tiles_per_cta);
  double *Ct = Cbase + (long long)item * stride + c_off + (long long)t0 * TILE;

  auto cp16 = [](double *dst, const double *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
  };
  auto prefetch_c = [&](const double *src) {
#pragma unroll 8
    for (int idx = tid; idx < TILE * TILE / 2; idx += NTH) {
      const int c = idx >> 6, r2 = idx & 63;
      cp16(Cb + c * LD_CB + 2 * r2, src + 2 * r2 + (long long)c * ld);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

  // diagonal tile, packed by groups of 8 columns
#pragma unroll
  for (int cb = 0; cb < 16; cb++) {
    const int rows2 = (TILE - 8 * cb) / 2;          // 16-byte chunks per column of this group
    for (int idx = tid; idx < 8 * rows2; idx += NTH) {
      const int kk = idx / rows2, r2 = idx - kk * rows2;
      cp16(Lp + lpk_off(cb) + kk * lpk_ld(cb) + 2 * r2, Ld + (8 * cb + 2 * r2) + (long long)(8 * cb + kk) * ld);
    }
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  prefetch_c(Ct);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  if (tid < TILE) invd[tid] = 1.0 / Lp[lpk_off(tid >> 3) + (tid & 7) * lpk_ld(tid >> 3) + (tid & 7)];
  __syncthreads();

  const unsigned FULL = 0xffffffffu;
  const int qbase = lane & ~3;
  const int r0 = warp * 8 * MT;
  for (int ti = t0; ti < t1; ti++) {
    double acc[MT][16][2];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
      for (int nt = 0; nt < 16; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) acc[mt][nt][e] = Cb[(nt * 8 + 2 * t + e) * LD_CB + r0 + mt * 8 + g];
    __syncthreads();  // everybody has taken its rows out of the buffer
    if (ti + 1 < t1) prefetch_c(Ct + TILE);  // next row tile, in flight during the arithmetic below

#pragma unroll
    for (int cb = 0; cb < 16; cb++) {
      const double *Lg = Lp + lpk_off(cb);  // group cb: Lg[kk*ld + (n - 8cb)]
      const int ldg = lpk_ld(cb);
      // (i) substitution inside the 8x8 diagonal block; a row's 8 values live in one quad
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const int owner = qbase + (c >> 1);
        const double dinv = invd[cb * 8 + c];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
          double xv = acc[mt][cb][c & 1] * dinv;
          if (t == (c >> 1)) acc[mt][cb][c & 1] = xv;
          xv = __shfl_sync(FULL, xv, owner);
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int cp = 2 * t + e;
            if (cp > c) acc[mt][cb][e] = fma(-xv, Lg[c * ldg + cp], acc[mt][cb][e]);
          }
        }
      }
      // (ii) re-layout the solved 8 columns from accumulator layout to mma A-operand layout
      double af[MT][2];
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int ks = 0; ks < 2; ks++) {
          const int src = qbase + 2 * ks + (t >> 1);
          const double v0 = __shfl_sync(FULL, acc[mt][cb][0], src);
          const double v1 = __shfl_sync(FULL, acc[mt][cb][1], src);
          af[mt][ks] = (t & 1) ? v1 : v0;
        }
      // (iii) rank-8 update of the columns to the right:  C[:, n] -= X[:, cb] * L[n, cb]^T
      // (k-step outermost: the two DMMAs that accumulate into the same tile are a whole sweep apart)
#pragma unroll
      for (int ks = 0; ks < 2; ks++) {
#pragma unroll
        for (int nt = 0; nt < 16; nt++) {
          if (nt > cb) {
            const double bfv = -Lg[(ks * 4 + t) * ldg + (nt - cb) * 8 + g];
#pragma unroll
            for (int mt = 0; mt < MT; mt++) dmma884p(acc[mt][nt], af[mt][ks], bfv);
          }
        }
      }
    }

#pragma unroll
    for (int mt = 0; mt < MT; mt++)
#pragma unroll
      for (int nt = 0; nt < 16; nt++) {
        const int r = r0 + mt * 8 + g, c = nt * 8 + 2 * t;
        Ct[r + (long long)c * ld] = acc[mt][nt][0];
        Ct[r + (long long)(c + 1) * ld] = acc[mt][nt][1];
      }
    Ct += TILE;
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();  // the next tile has landed for everybody
  }
}

// ------------------------------------------------------------------------------------------------
// Round-2 panel kernels: left-looking over 8-column blocks with the tile resident in SHARED memory.
//
// The round-1 kernels keep the 128x128 tile in registers in mma accumulator layout, which forces the
// sixteen block steps to be fully unrolled (register arrays need static indices): potrf_tile_kernel_v2
// is 28 000 SASS instructions, each executed once -- it runs at instruction-fetch speed (46 us per tile,
// 5 600 cycles per block step against ~1 500 of dependent arithmetic).  With the tile in shared memory
// every index is dynamic, the block loop stays rolled (a few hundred instructions), and the left-looking
// order needs no register-resident trailing matrix at all:
//
//   for cb = 0..15:   P  = T[:, 8cb:8cb+8] - T[:, 0:8cb] * T[8cb:8cb+8, 0:8cb]^T     DMMA, own rows
//                     D  = chol(P[8cb:8cb+8, :])   every warp redundantly, in registers, no shuffles
//                     T[below, 8cb:8cb+8] = P[below] D^-T    true substitution, one row per lane
//
// potrf_tile_ll_kernel: two block barriers per step (the diagonal block must be complete before it is
// factored; the solved panel must be visible before the next step's update).
// trsm_ll_kernel: the rows of X = C L^-T are independent, so each warp runs its own rows through all
// sixteen steps with no block barrier at all; a CTA takes 32, 64 or 128 rows so that a single large
// matrix (batch 1) still spreads one block column over the whole GPU.
// ------------------------------------------------------------------------------------------------
constexpr int LD_T = TILE + 4;  // column-major tile in shared memory, 132: conflict-free mma fragment reads
constexpr int POTRF_LL_SMEM_BYTES = TILE * LD_T * (int)sizeof(double);

// Cholesky of an 8x8 block held (lower part) in registers; every lane computes the same thing.
// inv[k] = 1 / L[k][k]; bad = first k with a non-positive (or NaN) pivot, -1 if none.
__device__ __forceinline__ void factor8_regs(double (&d)[8][8], double (&inv)[8], int &bad) {
  bad = -1;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const double piv = d[k][k];
    if (!(piv > 0.0) && bad < 0) bad = k;
    const double r = rsqrt(piv);
    inv[k] = r;
    d[k][k] = piv * r;
#pragma unroll
    for (int i = k + 1; i < 8; i++) d[i][k] *= r;
#pragma unroll
    for (int j = k + 1; j < 8; j++)
#pragma unroll
      for (int i = j; i < 8; i++) d[i][j] = fma(-d[i][k], d[j][k], d[i][j]);
  }
}

__global__ void __launch_bounds__(256, 1)
potrf_tile_ll_kernel(double *__restrict__ Lbase, long long ld, long long stride, long long diag_off, int index_base,
                     int n, int *info) {
  extern __shared__ __align__(16) double sm[];
  double *Ts = sm;  // Ts[c * LD_T + r] = T[r][c]
  __shared__ double Dsm[64];  // the updated diagonal block of the current step, Dsm[c * 8 + r]
  __shared__ int s_info;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double *T = Lbase + (long long)blockIdx.x * stride + diag_off;
  const int r0 = warp * 16;
  if (tid == 0) s_info = 0;

  for (int idx = tid; idx < TILE * TILE / 2; idx += 256) {
    const int c = idx >> 6, r2 = idx & 63;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(Ts + c * LD_T + 2 * r2);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(T + 2 * r2 + (long long)c * ld) : "memory");
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();

#pragma unroll 1
  for (int cb = 0; cb < 16; cb++) {
    const int c0 = 8 * cb;
    const bool active = (2 * warp + 1) >= cb;  // this warp still has rows at or below the diagonal block
    // ---- 1. left-looking update of the warp's rows of block column cb -------------------------------
    if (active && cb > 0) {
      double s[2][2][2];
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int ch = 0; ch < 2; ch++) s[mt][ch][0] = s[mt][ch][1] = 0.0;
      const bool m0 = (2 * warp) >= cb;  // the upper m-tile is still below / on the diagonal block
      const double *pa = Ts + t * LD_T + r0 + g;
      const double *pb = Ts + t * LD_T + c0 + g;
      for (int kb = 0; kb < cb; kb++) {
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
          const int ko = (8 * kb + 4 * ch) * LD_T;
          const double b = pb[ko];
          if (m0) dmma884p(s[0][ch], pa[ko], b);
          dmma884p(s[1][ch], pa[ko + 8], b);
        }
      }
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        if (mt == 0 && !m0) continue;
#pragma unroll
        for (int e = 0; e < 2; e++) {
          double *p = Ts + (c0 + 2 * t + e) * LD_T + r0 + mt * 8 + g;
          *p = (*p - s[mt][0][e]) - s[mt][1][e];
        }
      }
    }
    if (warp == (cb >> 1)) {
      // publish the updated diagonal block apart from the tile: its rows are overwritten in place by their
      // owner in step 2 while the other warps may still be reading it
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 2; e++) Dsm[(2 * t + e) * 8 + g] = Ts[(c0 + 2 * t + e) * LD_T + c0 + g];
    }
    __syncthreads();
    // ---- 2. diagonal block (redundantly per lane) + substitution, one row per lane -------------------
    if (active) {
      double d[8][8], inv[8];
#pragma unroll
      for (int j = 0; j < 8; j++)
#pragma unroll
        for (int i = j; i < 8; i++) d[i][j] = Dsm[j * 8 + i];
      int bad;
      factor8_regs(d, inv, bad);
      if (bad >= 0 && warp == (cb >> 1) && lane == 0 && s_info == 0) s_info = index_base + c0 + bad + 1;
      const int r = r0 + lane;
      if (lane < 16 && r >= c0) {
        double x[8];
#pragma unroll
        for (int c = 0; c < 8; c++) x[c] = Ts[(c0 + c) * LD_T + r];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          double sv = x[c];
#pragma unroll
          for (int cp = 0; cp < c; cp++) sv = fma(-x[cp], d[c][cp], sv);
          x[c] = sv * inv[c];
        }
        const int i = r - c0;  // < 8: a row of the diagonal block itself (the formula above reproduces L_D)
#pragma unroll
        for (int c = 0; c < 8; c++) Ts[(c0 + c) * LD_T + r] = (i < 8 && c > i) ? 0.0 : x[c];
      }
    }
    __syncthreads();
  }

  // write back: lower triangle = L, strict upper = 0 (Eigen matrixL() convention)
  for (int idx = tid; idx < TILE * TILE / 2; idx += 256) {
    const int c = idx >> 6, r = 2 * (idx & 63);
    const double2 v = *reinterpret_cast<const double2 *>(Ts + c * LD_T + r);
    *reinterpret_cast<double2 *>(T + r + (long long)c * ld) = make_double2(r >= c ? v.x : 0.0, r + 1 >= c ? v.y : 0.0);
  }
  if (tid == 0 && s_info != 0 && s_info <= n) {
    if (info[blockIdx.x] == 0) info[blockIdx.x] = s_info;
  }
}

// X L^T = C for ROWS = 32 MT consecutive rows of a block column (MODE-0 semantics of trsm_tile_kernel),
// four warps of 8 MT rows each.  blockIdx.x = row chunk, blockIdx.y = batch item.
template <int MT>
struct TrsmLL {
  static constexpr int ROWS = 32 * MT;
  static constexpr int LD_C = ROWS + 4;
  static constexpr int SMEM_BYTES = (LPK_DOUBLES + TILE + TILE * LD_C) * (int)sizeof(double);
};

template <int MT>
__global__ void __launch_bounds__(128)
trsm_ll_kernel(const double *Ldiag_base, double *Cbase, long long ld, long long stride, long long diag_off, long long c_off) {
  constexpr int ROWS = TrsmLL<MT>::ROWS, LD_C = TrsmLL<MT>::LD_C;
  extern __shared__ __align__(16) double sm[];
  double *Lp = sm;                  // packed lower triangle of the diagonal tile (see lpk_off / lpk_ld)
  double *invd = sm + LPK_DOUBLES;  // 1 / L[n][n]
  double *Cs = invd + TILE;         // Cs[c * LD_C + r] = C[r][c]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const double *Ld = Ldiag_base + (long long)blockIdx.y * stride + diag_off;
  double *Ct = Cbase + (long long)blockIdx.y * stride + c_off + (long long)blockIdx.x * ROWS;

  auto cp16 = [](double *dst, const double *src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
  };
#pragma unroll 1
  for (int cb = 0; cb < 16; cb++) {
    const int rows2 = (TILE - 8 * cb) / 2;
    for (int idx = tid; idx < 8 * rows2; idx += 128) {
      const int kk = idx / rows2, r2 = idx - kk * rows2;
      cp16(Lp + lpk_off(cb) + kk * lpk_ld(cb) + 2 * r2, Ld + (8 * cb + 2 * r2) + (long long)(8 * cb + kk) * ld);
    }
  }
  for (int idx = tid; idx < TILE * ROWS / 2; idx += 128) {
    const int c = idx / (ROWS / 2), r2 = idx - c * (ROWS / 2);
    cp16(Cs + c * LD_C + 2 * r2, Ct + 2 * r2 + (long long)c * ld);
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  if (tid < TILE) invd[tid] = 1.0 / Lp[lpk_off(tid >> 3) + (tid & 7) * lpk_ld(tid >> 3) + (tid & 7)];
  __syncthreads();

  const int r0 = warp * 8 * MT;
#pragma unroll 1
  for (int cb = 0; cb < 16; cb++) {
    const int c0 = 8 * cb;
    const double *Lg = Lp + lpk_off(cb);
    const int ldg = lpk_ld(cb);
    if (cb > 0) {
      // S = X[rows, 0:c0] * L[c0:c0+8, 0:c0]^T, two independent accumulation chains per m-tile
      double s[MT][2][2];
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int ch = 0; ch < 2; ch++) s[mt][ch][0] = s[mt][ch][1] = 0.0;
      const double *pa = Cs + t * LD_C + r0 + g;
      for (int kb = 0; kb < cb; kb++) {
        const double *Lk = Lp + lpk_off(kb) + (c0 + g - 8 * kb);
        const int ldk = lpk_ld(kb);
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
          const double b = Lk[(4 * ch + t) * ldk];
          const double *a = pa + (8 * kb + 4 * ch) * LD_C;
#pragma unroll
          for (int mt = 0; mt < MT; mt++) dmma884p(s[mt][ch], a[8 * mt], b);
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; mt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          double *p = Cs + (c0 + 2 * t + e) * LD_C + r0 + mt * 8 + g;
          *p = (*p - s[mt][0][e]) - s[mt][1][e];
        }
      __syncwarp();
    }
    // substitution against the 8x8 diagonal block, one row per lane
    if (lane < 8 * MT) {
      const int r = r0 + lane;
      double x[8];
#pragma unroll
      for (int c = 0; c < 8; c++) x[c] = Cs[(c0 + c) * LD_C + r];
#pragma unroll
      for (int c = 0; c < 8; c++) {
        double sv = x[c];
#pragma unroll
        for (int cp = 0; cp < c; cp++) sv = fma(-x[cp], Lg[cp * ldg + c], sv);
        x[c] = sv * invd[c0 + c];
      }
#pragma unroll
      for (int c = 0; c < 8; c++) Cs[(c0 + c) * LD_C + r] = x[c];
    }
    __syncwarp();
  }
  __syncthreads();
  for (int idx = tid; idx < TILE * ROWS / 2; idx += 128) {
    const int c = idx / (ROWS / 2), r2 = idx - c * (ROWS / 2);
    *reinterpret_cast<double2 *>(Ct + 2 * r2 + (long long)c * ld) = *reinterpret_cast<const double2 *>(Cs + c * LD_C + 2 * r2);
  }
}

int panel_smem_setup(Handle *h) {
  GPB_CUDA(h, cudaFuncSetAttribute(trsm_tiles_pipelined_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_PIPE_SMEM_BYTES));
  GPB_CUDA(h, cudaFuncSetAttribute(trsm_tile_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM_BYTES));
  GPB_CUDA(h, cudaFuncSetAttribute(trsm_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM_BYTES));
  GPB_CUDA(h, cudaFuncSetAttribute(potrf_tile_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTRF_LL_SMEM_BYTES));
  GPB_CUDA(h, cudaFuncSetAttribute(trsm_ll_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TrsmLL<1>::SMEM_BYTES));
  GPB_CUDA(h, cudaFuncSetAttribute(trsm_ll_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TrsmLL<2>::SMEM_BYTES));
  GPB_CUDA(h, cudaFuncSetAttribute(trsm_ll_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TrsmLL<4>::SMEM_BYTES));
  return 0;
}

int launch_potrf_tile_at(Handle *h, double *L, long long ld, long long stride, long long diag_off, int index_base,
                         int n, int batch, int *info) {
  ProfScope ps__(h, PC_POTRF);
  if (h->panel_impl == 1)
    potrf_tile_kernel_v2<<<batch, 256, 0, h->stream>>>(L, ld, stride, diag_off, index_base, n, info);
  else
    potrf_tile_ll_kernel<<<batch, 256, POTRF_LL_SMEM_BYTES, h->stream>>>(L, ld, stride, diag_off, index_base, n, info);
  GPB_LAUNCH_CHECK(h);
  return 0;
}

int launch_potrf_tile(Handle *h, double *L, long long ld, long long stride, int tile_idx, int n, int batch, int *info) {
  return launch_potrf_tile_at(h, L, ld, stride, (long long)tile_idx * TILE * (ld + 1), tile_idx * TILE, n, batch, info);
}

// X = C L^-T for `ntiles` consecutive 128-row tiles starting at element offset c_off (tile step = 128
// rows), against the diagonal tile at element offset diag_off.
int launch_trsm_tiles_at(Handle *h, double *L, long long ld, long long stride, long long diag_off, long long c_off,
                         int ntiles, int batch) {
  if (ntiles <= 0) return 0;
  const long long total = (long long)ntiles * batch;
  ProfScope ps__(h, PC_TRSM);
  if (h->panel_impl == 1) {
    // round-1 kernels: tiles per CTA as many as still leave every SM several CTAs
    const int tpc = total >= 148 * 16 ? 4 : (total >= 148 * 6 ? 2 : 1);
    if (tpc == 1 || h->trsm_pipelined == 0) {
      dim3 grid(ntiles, batch);
      trsm_tile_kernel<0><<<grid, 256, TRSM_SMEM_BYTES, h->stream>>>(L, L, ld, stride, diag_off, 0, c_off, TILE);
    } else {
      dim3 grid((ntiles + tpc - 1) / tpc, batch);
      trsm_tiles_pipelined_kernel<2><<<grid, 256, TRSM_PIPE_SMEM_BYTES, h->stream>>>(L, L, ld, stride, diag_off, c_off, ntiles, tpc);
    }
  } else {
    // rows per CTA: 32 while that still fits one wave of two CTAs per SM (a single large matrix), else 128
    int mt = h->trsm_mt_override;
    if (mt != 1 && mt != 2 && mt != 4) mt = (total * 4 <= 2 * 148) ? 1 : ((total * 2 <= 148) ? 2 : 4);
    dim3 grid(ntiles * (4 / mt), batch);
    if (mt == 1) trsm_ll_kernel<1><<<grid, 128, TrsmLL<1>::SMEM_BYTES, h->stream>>>(L, L, ld, stride, diag_off, c_off);
    else if (mt == 2) trsm_ll_kernel<2><<<grid, 128, TrsmLL<2>::SMEM_BYTES, h->stream>>>(L, L, ld, stride, diag_off, c_off);
    else trsm_ll_kernel<4><<<grid, 128, TrsmLL<4>::SMEM_BYTES, h->stream>>>(L, L, ld, stride, diag_off, c_off);
  }
  GPB_LAUNCH_CHECK(h);
  return 0;
}

int launch_trsm_tiles(Handle *h, double *L, long long ld, long long stride, int tile_col, int ntiles_below, int batch) {
  return launch_trsm_tiles_at(h, L, ld, stride, (long long)tile_col * TILE * (ld + 1),
                              (long long)(tile_col + 1) * TILE + (long long)tile_col * TILE * ld, ntiles_below, batch);
}

// inverse of `ntiles` diagonal tiles: L tiles at l_off + t*l_step, W tiles at w_off + t*w_step
int launch_tile_inverse_at(Handle *h, const double *L, long long ld, long long l_off, long long l_step, double *W,
                           long long w_off, long long w_step, long long stride, int ntiles, int batch) {
  dim3 grid(ntiles, batch);
  ProfScope ps__(h, PC_TRSM);
  trsm_tile_kernel<1><<<grid, 256, TRSM_SMEM_BYTES, h->stream>>>(L, W, ld, stride, l_off, l_step, w_off, w_step);
  GPB_LAUNCH_CHECK(h);
  return 0;
}

int launch_tile_inverse(Handle *h, const double *L, double *W, long long ld, long long stride, int ntiles, int batch) {
  const long long step = (long long)TILE * (ld + 1);
  return launch_tile_inverse_at(h, L, ld, 0, step, W, 0, step, stride, ntiles, batch);
}

}  // namespace gpb
