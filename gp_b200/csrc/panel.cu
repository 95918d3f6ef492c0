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
// POTRF of one 128x128 tile.  256 threads = 16x16 grid; thread (ti,tj) owns A[ti+16a][tj+16b].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
potrf_tile_kernel(double *__restrict__ Lbase, long long ld, long long stride, long long diag_off, int index_base,
                  int n, int *info) {
  __shared__ double col[2][TILE];
  __shared__ int s_info;
  const int tid = threadIdx.x;
  const int ti = tid & 15, tj = tid >> 4;
  double *T = Lbase + (long long)blockIdx.x * stride + diag_off;
  double r[8][8];
#pragma unroll
  for (int b = 0; b < 8; b++)
#pragma unroll
    for (int a = 0; a < 8; a++) r[a][b] = T[(ti + 16 * a) + (long long)(tj + 16 * b) * ld];
  if (tid == 0) s_info = 0;

  for (int k = 0; k < TILE; k++) {
    const int kb = k >> 4, kt = k & 15;
    double *cbuf = col[k & 1];
    if (tj == kt) {
      // this thread column owns column k: publish it (rows >= k are meaningful)
#pragma unroll
      for (int b = 0; b < 8; b++)
        if (b == kb) {
#pragma unroll
          for (int a = 0; a < 8; a++) cbuf[ti + 16 * a] = r[a][b];
        }
    }
    __syncthreads();
    const double piv = cbuf[k];
    if (!(piv > 0.0)) {
      // non-positive (or NaN) pivot: record the first one, keep going so the CTA stays in step
      if (tid == 0 && s_info == 0) s_info = index_base + k + 1;
    }
    const double dgl = sqrt(piv);
    const double inv = 1.0 / dgl;
    double li[8], lj[8];
#pragma unroll
    for (int a = 0; a < 8; a++) li[a] = cbuf[ti + 16 * a] * inv;
#pragma unroll
    for (int b = 0; b < 8; b++) lj[b] = cbuf[tj + 16 * b] * inv;
#pragma unroll
    for (int b = 0; b < 8; b++) {
      if (b >= kb) {
        const int j = tj + 16 * b;
#pragma unroll
        for (int a = 0; a < 8; a++) {
          if (a >= b) {  // i >= j can only hold when a >= b (ti,tj < 16), refined below
            const int i = ti + 16 * a;
            if (j > k && i >= j) r[a][b] = fma(-li[a], lj[b], r[a][b]);
          }
        }
      }
    }
    if (tj == kt) {
#pragma unroll
      for (int b = 0; b < 8; b++)
        if (b == kb) {
#pragma unroll
          for (int a = 0; a < 8; a++) {
            const int i = ti + 16 * a;
            if (i > k) r[a][b] = li[a];
            else if (i == k) r[a][b] = dgl;
          }
        }
    }
  }
  // write back: lower triangle = L, strict upper = 0 (Eigen matrixL() convention)
#pragma unroll
  for (int b = 0; b < 8; b++)
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const int i = ti + 16 * a, j = tj + 16 * b;
      T[i + (long long)j * ld] = (i >= j) ? r[a][b] : 0.0;
    }
  __syncthreads();
  if (tid == 0 && s_info != 0 && s_info <= n) {
    // keep the smallest index over tiles (tiles are processed in increasing order, so first wins)
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

  // stage the diagonal tile (column-major copy, 16B vectors)
  for (int idx = tid; idx < TILE * TILE / 2; idx += 256) {
    const int k = idx >> 6, n2 = idx & 63;
    const double2 v = *reinterpret_cast<const double2 *>(Ld + 2 * n2 + (long long)k * ld);
    *reinterpret_cast<double2 *>(Ls + k * LD_L + 2 * n2) = v;
  }
  __syncthreads();
  if (tid < TILE) invd[tid] = 1.0 / Ls[tid * LD_L + tid];
  __syncthreads();

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

int panel_smem_setup(Handle *h) {
  GPB_CUDA(h, cudaFuncSetAttribute(trsm_tile_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM_BYTES));
  GPB_CUDA(h, cudaFuncSetAttribute(trsm_tile_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM_BYTES));
  return 0;
}

int launch_potrf_tile_at(Handle *h, double *L, long long ld, long long stride, long long diag_off, int index_base,
                         int n, int batch, int *info) {
  ProfScope ps__(h, PC_POTRF);
  potrf_tile_kernel<<<batch, 256, 0, h->stream>>>(L, ld, stride, diag_off, index_base, n, info);
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
  dim3 grid(ntiles, batch);
  ProfScope ps__(h, PC_TRSM);
  trsm_tile_kernel<0><<<grid, 256, TRSM_SMEM_BYTES, h->stream>>>(L, L, ld, stride, diag_off, 0, c_off, TILE);
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
