// C-ABI entry points for the NON-CENTRED LATENT models (SURVEY CS-E, row f-2): models/exact_gp.stan:17-25,
// fit_full_gp.stan:18-26, westbrook_exact.stan:17-24 (one mat-vec f = L z) and heteroscedastic.stan:23-32 (two
// mat-vecs on one L).  Stan differentiates THROUGH cholesky_decompose; this is that reverse sweep as one pass:
//
//   forward   K = cov_exp_quad(x, alpha, rho) + diag_add I ;  L = chol(K) ;  f_m = L z_m
//   backward  given fbar_m:   zbar_m = u_m = L^T fbar_m
//             Lbar = sum_m tril(fbar_m z_m^T)                         (adjoint of the mat-vecs)
//             P    = Phi(L^T Lbar) = Phi(sum_m u_m z_m^T)             (Phi: lower triangle, diagonal halved)
//             Kbar = L^-T P L^-1 = W^T (P W),  W = L^-1               (adjoint of the Cholesky)
//             thetabar = sum_ij Kbar_ij dK_ij/dtheta                  (adjoint of cov_exp_quad)
// P W needs no GEMM: (P W)_ik = sum_m u_m,i * (sum_{k<=j<i} z_m,j W_jk + z_m,i W_ik / 2) is a running sum down each column
// of diag(z) W.  So the pass is one triangular inverse (N^3/3) and ONE tile GEMM W^T (P W) whose tiles go straight from
// the accumulators into the contraction with dK/dtheta (recomputed from x): N^3 in total for ALL parameters, where the
// forward-mode route (gpb200_se_chol_tangent) costs 5 N^3 / 3 per parameter.
#include "host.cuh"

using namespace gpb;

namespace {
enum { TK_LATENT_G = 60 };

unsigned long long hash_doubles(const double *x, int n) {   // FNV-1a over the bytes of a HOST array
  unsigned long long hsh = 1469598103934665603ULL;
  const unsigned char *b = reinterpret_cast<const unsigned char *>(x);
  for (size_t i = 0; i < (size_t)n * 8; i++) { hsh ^= b[i]; hsh *= 1099511628211ULL; }
  return hsh;
}

// T[i][k] (+)= u_i * (sum_{k <= j < i} z_j W[j][k] + z_i W[i][k] / 2) for i >= k; zero above the diagonal.
// One warp per column k; 32 rows at a time with a shuffle prefix sum: reads of W are contiguous along the column.
__global__ void __launch_bounds__(256) latent_pw_kernel(int np, const double *__restrict__ W, const double *__restrict__ u,
                                                        const double *__restrict__ z, int accumulate, double *__restrict__ T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + warp;
  if (k >= np) return;
  const double *wc = W + (long long)k * np;
  double *tc = T + (long long)k * np;
  const int i0 = k & ~31;
  for (int i = lane; i < i0; i += 32)
    if (!accumulate) tc[i] = 0.0;
  double carry = 0.0;
  for (int base = i0; base < np; base += 32) {
    const int i = base + lane;
    const double v = (i >= k) ? z[i] * wc[i] : 0.0;
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const double val = (i >= k) ? u[i] * (carry + (incl - v) + 0.5 * v) : 0.0;
    tc[i] = accumulate ? tc[i] + val : val;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
}

__global__ void __launch_bounds__(256) latent_sum_kernel(long long nrec, const double *__restrict__ partial, double *__restrict__ out2) {
  __shared__ double red[8][2];
  double s0 = 0, s1 = 0;
  for (long long r = threadIdx.x; r < nrec; r += 256) { s0 += partial[r * 4]; s1 += partial[r * 4 + 1]; }
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s0; red[threadIdx.x >> 5][1] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; w++) { a += red[w][0]; b += red[w][1]; }
    out2[0] = a; out2[1] = b;
  }
}

// builds (or finds cached) the padded factor of K = cov_exp_quad(x, alpha, rho) + diag_add I; returns LAPACK info
int latent_factor(Handle *h, int n, const double *x, double alpha, double rho, double diag_add, int *info_out) {
  const int np = round_up(n, TILE);
  const size_t bytes = (size_t)np * np * 8;
  const unsigned long long xh = h->device_ptrs ? (unsigned long long)(uintptr_t)x : hash_doubles(x, n);
  if (h->latent_L && h->latent_n == n && h->latent_key[0] == alpha && h->latent_key[1] == rho && h->latent_key[2] == diag_add &&
      h->latent_xhash == xh && !h->device_ptrs) {
    *info_out = 0;
    return 0;
  }
  if (h->latent_bytes < bytes) {
    if (h->latent_L) { GPB_CUDA(h, cudaStreamSynchronize(h->stream)); GPB_CUDA(h, cudaFree(h->latent_L)); h->latent_L = nullptr; }
    GPB_CUDA(h, cudaMalloc(&h->latent_L, bytes));
    h->latent_bytes = bytes;
  }
  h->latent_n = 0;
  Arena a;
  RC(ws_reserve(h, pad256(n * 8) + 1024, &a));
  double *dx = a.take<double>(n), *dth = a.take<double>(3);
  int *info = a.take<int>(1);
  RC(to_device(h, x, dx, n));
  const double th[3] = {alpha, rho, 0.0};
  GPB_CUDA(h, cudaMemcpyAsync(dth, th, sizeof(th), cudaMemcpyHostToDevice, h->stream));
  GPB_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int), h->stream));
  GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  RC(launch_gram_se_batched(h, n, np, dx, 0, dth, diag_add, 1, h->latent_L, 0, 1));
  RC(chol_batched(h, h->latent_L, np, (long long)np * np, n, 1, info));
  RC(read_info(h, info, info_out));
  if (*info_out == 0) {
    h->latent_n = n;
    h->latent_key[0] = alpha; h->latent_key[1] = rho; h->latent_key[2] = diag_add;
    h->latent_xhash = xh;
  }
  return 0;
}
}  // namespace

// f_m = L z_m for m = 0..nvec-1 (z, f: nvec vectors of length n, consecutive).  L = chol(cov_exp_quad(x, alpha, rho) +
// diag_add I) is kept on the device for gpb200_latent_backward.
extern "C" int gpb200_latent_forward(gpb200_handle_t h, int n, const double *x, double alpha, double rho, double diag_add, int nvec,
                                     const double *z, double *f) {
  CHECK_H(h);
  if (n < 1 || nvec < 1) BAD_ARG(h, 2, "latent_forward: n and nvec must be >= 1");
  int info = 0;
  RC(latent_factor(h, n, x, alpha, rho, diag_add, &info));
  if (info) return info;
  const int np = round_up(n, TILE);
  Arena a;
  RC(ws_reserve(h, 2 * pad256((size_t)np * 8 * nvec) + 1024, &a));
  double *dz = a.take<double>((size_t)np * nvec), *df = a.take<double>((size_t)np * nvec);
  for (int m = 0; m < nvec; m++) {
    RC(to_device(h, z + (size_t)m * n, dz + (size_t)m * np, n));
    RC(launch_trmv_lower_n(h, np, h->latent_L, 0, dz + (size_t)m * np, 0, n, df + (size_t)m * np, 0, 1));
    RC(from_device(h, df + (size_t)m * np, f + (size_t)m * n, n * sizeof(double)));
  }
  return finish(h);
}

// Reverse sweep: given fbar_m = d lp / d f_m, returns zbar_m = L^T fbar_m (nvec x n) and
// theta_bar = (d lp / d alpha, d lp / d rho) THROUGH f_m = L z_m and the Cholesky.
extern "C" int gpb200_latent_backward(gpb200_handle_t h, int n, const double *x, double alpha, double rho, double diag_add, int nvec,
                                      const double *z, const double *fbar, double *theta_bar, double *zbar) {
  CHECK_H(h);
  if (n < 1 || nvec < 1) BAD_ARG(h, 2, "latent_backward: n and nvec must be >= 1");
  int info = 0;
  RC(latent_factor(h, n, x, alpha, rho, diag_add, &info));
  if (info) return info;
  const int np = round_up(n, TILE), nt = np / TILE;
  const size_t mat = (size_t)np * np;
  TaskList tl;
  const long long key = tkey(TK_LATENT_G, nt);
  if (!cached(h, key, &tl)) {
    std::vector<TileTask> t;
    for (int I = 0; I < nt; I++)
      for (int J = 0; J < nt; J++) {
        const int k0 = std::max(I, J);   // W[k][I] = 0 for k < I, (P W)[k][J] = 0 for k < J
        t.push_back({k0 * TILE, I * TILE, k0 * TILE, J * TILE, I * TILE, J * TILE, (nt - k0) * TILE,
                     TF_FULL_WEIGHT | (I >= J ? TF_A_TRI_FIRST : 0) | (J >= I ? TF_B_TRI_FIRST : 0)});
      }
    sort_desc(t, 0);
    std::vector<int> off = {0, (int)t.size()};
    RC(upload_tasks(h, key, t, off, &tl));
  }
  const int ntasks = tl.count(0);
  Arena a;
  RC(ws_reserve(h, 3 * pad256(mat * 8) + (3 * (size_t)nvec + 2) * pad256((size_t)np * 8) + pad256((size_t)ntasks * 4 * 4 * 8) + 4096, &a));
  double *W = a.take<double>(mat), *S = a.take<double>(mat), *T = a.take<double>(mat);
  double *dz = a.take<double>((size_t)np * nvec), *dfb = a.take<double>((size_t)np * nvec), *du = a.take<double>((size_t)np * nvec);
  double *dx = a.take<double>(np), *zero = a.take<double>(np), *dth = a.take<double>(3), *out2 = a.take<double>(2);
  double *partial = a.take<double>((size_t)ntasks * 4 * 4);
  if (!partial) BAD_ARG(h, 1002, "latent_backward: workspace exhausted");
  GPB_CUDA(h, cudaMemsetAsync(dz, 0, sizeof(double) * np * nvec, h->stream));
  GPB_CUDA(h, cudaMemsetAsync(dfb, 0, sizeof(double) * np * nvec, h->stream));
  GPB_CUDA(h, cudaMemsetAsync(zero, 0, sizeof(double) * np, h->stream));
  GPB_CUDA(h, cudaMemsetAsync(dx, 0, sizeof(double) * np, h->stream));
  RC(to_device(h, x, dx, n));
  const double th[3] = {alpha, rho, 0.0};
  GPB_CUDA(h, cudaMemcpyAsync(dth, th, sizeof(th), cudaMemcpyHostToDevice, h->stream));
  GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  GPB_CUDA(h, cudaMemcpyAsync(W, h->latent_L, mat * 8, cudaMemcpyDeviceToDevice, h->stream));
  RC(trtri_batched(h, W, S, np, (long long)mat, 1));
  for (int m = 0; m < nvec; m++) {
    RC(to_device(h, z + (size_t)m * n, dz + (size_t)m * np, n));
    RC(to_device(h, fbar + (size_t)m * n, dfb + (size_t)m * np, n));
    RC(launch_trmv_lower_t(h, np, h->latent_L, 0, dfb + (size_t)m * np, 0, du + (size_t)m * np, 0, 1));   // u = zbar = L^T fbar
    ProfScope ps__(h, PC_SOLVE);
    latent_pw_kernel<<<(np + 7) / 8, 256, 0, h->stream>>>(np, W, du + (size_t)m * np, dz + (size_t)m * np, m > 0, T);
    GPB_LAUNCH_CHECK(h);
    RC(from_device(h, du + (size_t)m * np, zbar + (size_t)m * n, n * sizeof(double)));
  }
  GemmParams p{};
  p.A = mref(W, np, 0);
  p.B = mref(T, np, 0);
  p.C = mref(nullptr, np, 0);
  p.tasks = tl.at(0);
  p.x = dx; p.x_stride = 0;
  p.avec = zero; p.a_stride = 0;
  p.theta = dth;
  p.partial = partial;
  p.n = n;
  p.ntasks = ntasks;
  RC(launch_gemm(h, LAYOUT_TN, EPI_TRACE, p, ntasks, 1));
  {
    ProfScope ps__(h, PC_OTHER);
    latent_sum_kernel<<<1, 256, 0, h->stream>>>((long long)ntasks * gemm_nsplit(h, ntasks, 1), partial, out2);
    GPB_LAUNCH_CHECK(h);
  }
  double r[2];
  GPB_CUDA(h, cudaMemcpyAsync(r, out2, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
  GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  // the epilogue accumulated M = 0 - G: Kbar contracted with e and e d^2 is minus those sums
  const double tb[2] = {-2.0 * alpha * r[0], -alpha * alpha * r[1] / (rho * rho * rho)};
  if (h->device_ptrs) {
    GPB_CUDA(h, cudaMemcpyAsync(theta_bar, tb, sizeof(tb), cudaMemcpyHostToDevice, h->stream));
    GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  } else {
    theta_bar[0] = tb[0];
    theta_bar[1] = tb[1];
  }
  return 0;
}
