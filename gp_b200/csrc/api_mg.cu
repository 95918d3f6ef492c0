// C-ABI entry points: per-rank building blocks of the block-cyclic multi-GPU Cholesky (config 5).
#include "host.cuh"

using namespace gpb;

// =================================================================================================
// =================================================================================================
// (e) building blocks of the block-cyclic multi-GPU Cholesky (config 5).  DEVICE pointers only.
// A "panel" is a block column of the padded matrix stored compactly: rows [col0, np) x ncols
// columns, leading dimension ldp >= np - col0.  The collective (panel broadcast over NCCL) lives
// in gp_b200/block_cyclic.py; these calls are the per-rank compute between collectives.
// =================================================================================================
namespace {
long long mgkey(int kind, int a, int b, int c, int d) {
  return ((long long)kind << 52) | ((long long)(a & 0x1fff) << 39) | ((long long)(b & 0x1fff) << 26) |
         ((long long)(c & 0x1fff) << 13) | (long long)(d & 0x1fff);
}
enum { TK_MG_FACTOR = 40, TK_MG_UPDATE = 41 };

int mg_check_panel(Handle *h, int n, int col0, int ncols, long long ldp, int *np_out) {
  const int np = round_up(n, TILE);
  if (n < 1 || col0 < 0 || ncols < TILE || (col0 % TILE) || (ncols % TILE) || col0 + ncols > np)
    BAD_ARG(h, 3, "mg: panel must be tile aligned and inside the padded matrix");
  if (ldp < np - col0 || (ldp & 1)) BAD_ARG(h, 6, "mg: ldp must be even and >= np - col0");
  *np_out = np;
  return 0;
}
}  // namespace

extern "C" int gpb200_mg_gram_panel(gpb200_handle_t h, int n, const double *x, double alpha, double rho,
                                    double diag_add, int col0, int ncols, double *P, long long ldp) {
  CHECK_H(h);
  int np;
  RC(mg_check_panel(h, n, col0, ncols, ldp, &np));
  return launch_gram_se_panel(h, n, np, x, alpha, rho, diag_add, col0, ncols, P, ldp);
}

extern "C" int gpb200_mg_panel_factor(gpb200_handle_t h, int n, int col0, int ncols, double *P, long long ldp,
                                      int *info_dev) {
  CHECK_H(h);
  int np;
  RC(mg_check_panel(h, n, col0, ncols, ldp, &np));
  const int ntp = ncols / TILE, nrt = (np - col0) / TILE;
  TaskList tl;
  const long long key = mgkey(TK_MG_FACTOR, nrt, ntp, 0, 0);
  if (!cached(h, key, &tl)) {
    std::vector<TileTask> t;
    std::vector<int> off(1, 0);
    for (int jl = 0; jl < ntp; jl++) {
      if (jl > 0)
        for (int i = jl; i < nrt; i++) t.push_back({i * TILE, 0, jl * TILE, 0, i * TILE, jl * TILE, jl * TILE, i == jl});
      off.push_back((int)t.size());
    }
    RC(upload_tasks(h, key, t, off, &tl));
  }
  GemmParams p{};
  p.A = mref(P, ldp, 0);
  p.B = mref(P, ldp, 0);
  p.C = mref(P, ldp, 0);
  p.C0 = mref(P, ldp, 0);
  p.alpha = -1.0;
  p.beta = 1.0;
  for (int jl = 0; jl < ntp; jl++) {
    if (tl.count(jl) > 0) {
      p.tasks = tl.at(jl);
      RC(launch_gemm(h, LAYOUT_NT, EPI_AXPBY, p, tl.count(jl), 1));
    }
    const long long doff = (long long)jl * TILE * (ldp + 1);
    RC(launch_potrf_tile_at(h, P, ldp, 0, doff, col0 + jl * TILE, n, 1, info_dev));
    RC(launch_trsm_tiles_at(h, P, ldp, 0, doff, doff + TILE, nrt - 1 - jl, 1));
  }
  return 0;
}

extern "C" int gpb200_mg_panel_update(gpb200_handle_t h, int n, int pcol0, int pncols, const double *P, long long ldp,
                                      int ccol0, int cncols, double *Cp, long long ldc) {
  CHECK_H(h);
  int np;
  RC(mg_check_panel(h, n, pcol0, pncols, ldp, &np));
  RC(mg_check_panel(h, n, ccol0, cncols, ldc, &np));
  if (ccol0 < pcol0 + pncols) BAD_ARG(h, 7, "mg_panel_update: the target panel must lie right of the source panel");
  const int nt = np / TILE, d = (ccol0 - pcol0) / TILE, cnt = cncols / TILE, crt = nt - ccol0 / TILE, pk = pncols / TILE;
  TaskList tl;
  const long long key = mgkey(TK_MG_UPDATE, d, cnt, crt, pk);
  if (!cached(h, key, &tl)) {
    std::vector<TileTask> t;
    for (int jl = 0; jl < cnt; jl++)
      for (int il = jl; il < crt; il++)  // il, jl: tile coordinates local to the target panel
        t.push_back({(il + d) * TILE, 0, (jl + d) * TILE, 0, il * TILE, jl * TILE, pk * TILE, il == jl});
    std::vector<int> off = {0, (int)t.size()};
    RC(upload_tasks(h, key, t, off, &tl));
  }
  GemmParams p{};
  p.A = mref(const_cast<double *>(P), ldp, 0);
  p.B = mref(const_cast<double *>(P), ldp, 0);
  p.C = mref(Cp, ldc, 0);
  p.C0 = mref(Cp, ldc, 0);
  p.alpha = -1.0;
  p.beta = 1.0;
  p.tasks = tl.at(0);
  return launch_gemm(h, LAYOUT_NT, EPI_AXPBY, p, tl.count(0), 1);
}

// forward substitution through one factored panel: z[pcol0 .. +ncols) = solve, acc[rows below] +=
// L z.  y, acc, z are replicated device vectors of length np; wscratch holds one inverted tile
// (ldp x 128 doubles).
extern "C" int gpb200_mg_panel_trsv(gpb200_handle_t h, int n, int col0, int ncols, const double *P, long long ldp,
                                    const double *y, double *acc, double *z, double *wscratch) {
  CHECK_H(h);
  int np;
  RC(mg_check_panel(h, n, col0, ncols, ldp, &np));
  const int ntp = ncols / TILE, nrt = (np - col0) / TILE;
  for (int jl = 0; jl < ntp; jl++) {
    const long long doff = (long long)jl * TILE * (ldp + 1);
    RC(launch_tile_inverse_at(h, P, ldp, doff, 0, wscratch, 0, 0, 0, 1, 1));
    RC(launch_trsv_diag(h, ldp, 0, col0 + jl * TILE, wscratch, 0, y, 0, nullptr, n, acc, z, 0, 1));
    RC(launch_trsv_update(h, ldp, doff + TILE, col0 + jl * TILE, col0 + (jl + 1) * TILE, nrt - 1 - jl, P, 0, z, acc, 0, 1));
  }
  return 0;
}

namespace {
__global__ void panel_logdiag_kernel(int n, int col0, int ncols, const double *__restrict__ P, long long ldp,
                                     double *__restrict__ out) {
  __shared__ double red[8];
  double s = 0.0;
  for (int c = threadIdx.x; c < ncols; c += 256)
    if (col0 + c < n) s += log(P[c + (long long)c * ldp]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; w++) t += red[w];
    out[0] += t;
  }
}
}  // namespace

// out[0] += sum_{i in panel, i < n} log L_ii  (device scalar, accumulated across this rank's panels)
extern "C" int gpb200_mg_panel_logdiag(gpb200_handle_t h, int n, int col0, int ncols, const double *P, long long ldp,
                                       double *out) {
  CHECK_H(h);
  int np;
  RC(mg_check_panel(h, n, col0, ncols, ldp, &np));
  ProfScope ps__(h, PC_OTHER);
  panel_logdiag_kernel<<<1, 256, 0, h->stream>>>(n, col0, ncols, P, ldp, out);
  GPB_LAUNCH_CHECK(h);
  return 0;
}
