// C-ABI entry points: per-rank building blocks of the block-cyclic multi-GPU Cholesky (config 5).
#include <dlfcn.h>

#include "host.cuh"
#include "fastexp.cuh"

using namespace gpb;

// =================================================================================================
// NCCL, loaded at run time: the library a host process already carries (torch bundles one) is reused through
// RTLD_NOLOAD; a plain C / R host falls back to the system libnccl.so.2.  Only six entry points are needed.
// =================================================================================================
namespace {
struct NcclId { char b[128]; };
struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(void *) = nullptr;
  int (*CommInitRank)(void **, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void *) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int nccl_load(Handle *h) {
  if (g_nccl.lib) return 0;
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  const char *env = getenv("GPB200_NCCL_LIB");
  if (!lib && env) lib = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) {
    snprintf(h->err, sizeof(h->err), "mg: libnccl.so.2 not found (%s); set GPB200_NCCL_LIB", dlerror());
    return -1004;
  }
  *(void **)&g_nccl.GetUniqueId = dlsym(lib, "ncclGetUniqueId");
  *(void **)&g_nccl.CommInitRank = dlsym(lib, "ncclCommInitRank");
  *(void **)&g_nccl.CommDestroy = dlsym(lib, "ncclCommDestroy");
  *(void **)&g_nccl.Broadcast = dlsym(lib, "ncclBroadcast");
  *(void **)&g_nccl.AllReduce = dlsym(lib, "ncclAllReduce");
  *(void **)&g_nccl.GetErrorString = dlsym(lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.Broadcast || !g_nccl.AllReduce) {
    snprintf(h->err, sizeof(h->err), "mg: libnccl.so.2 lacks a required symbol");
    return -1004;
  }
  g_nccl.lib = lib;
  return 0;
}
#define GPB_NCCL(h, call)                                                                              \
  do {                                                                                                 \
    int r__ = (call);                                                                                  \
    if (r__ != 0) {                                                                                    \
      snprintf((h)->err, sizeof((h)->err), "%s failed: %s", #call,                                     \
               g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error");                     \
      return -1005;                                                                                    \
    }                                                                                                  \
  } while (0)
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0, NCCL_MAX = 2, NCCL_INT32 = 2;
constexpr int MG_EVENT_RING = 64;
}  // namespace

extern "C" int gpb200_mg_comm_id(gpb200_handle_t h, void *id128) {
  CHECK_H(h);
  RC(nccl_load(h));
  GPB_NCCL(h, g_nccl.GetUniqueId(id128));
  return 0;
}

extern "C" int gpb200_mg_comm_init(gpb200_handle_t h, const void *id128, int rank, int world) {
  CHECK_H(h);
  if (world < 1 || rank < 0 || rank >= world) BAD_ARG(h, 3, "mg_comm_init: bad rank / world");
  if (h->nccl_comm) BAD_ARG(h, 1, "mg_comm_init: the handle already owns a communicator");
  h->mg_rank = rank;
  h->mg_world = world;
  if (!h->cstream) {
    int lo = 0, hi = 0;
    GPB_CUDA(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    GPB_CUDA(h, cudaStreamCreateWithPriority(&h->cstream, cudaStreamNonBlocking, hi));
    for (int i = 0; i < MG_EVENT_RING; i++) {
      cudaEvent_t e;
      GPB_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->mg_events.push_back(e);
    }
  }
  if (world == 1) return 0;
  RC(nccl_load(h));
  NcclId id;
  memcpy(&id, id128, 128);
  GPB_NCCL(h, g_nccl.CommInitRank(&h->nccl_comm, world, id, rank));
  return 0;
}

extern "C" int gpb200_mg_comm_destroy(gpb200_handle_t h) {
  CHECK_H(h);
  if (h->cstream) GPB_CUDA(h, cudaStreamSynchronize(h->cstream));
  if (h->nccl_comm) {
    g_nccl.CommDestroy(h->nccl_comm);
    h->nccl_comm = nullptr;
  }
  h->mg_world = 1;
  h->mg_rank = 0;
  return 0;
}

// Broadcast of `count` doubles from `root`, enqueued on the handle's communication stream behind everything
// enqueued so far on the compute stream (the root's data is final, a receiver's buffer is no longer read).
// Returns a ticket; gpb200_mg_wait makes the compute stream wait for that collective.  With one rank there is no
// collective and the ticket stands for the point the compute stream had reached.
extern "C" int gpb200_mg_bcast(gpb200_handle_t h, double *buf, long long count, int root, long long *ticket) {
  CHECK_H(h);
  if (!h->cstream) BAD_ARG(h, 1, "mg_bcast: gpb200_mg_comm_init first");
  const long long t = h->mg_tickets++;
  if (ticket) *ticket = t;
  cudaEvent_t e = h->mg_events[(size_t)(t % MG_EVENT_RING)];
  GPB_CUDA(h, cudaEventRecord(e, h->stream));
  // one rank: no collective, but the ticket still orders "data final on the stream that produced it" before a wait on
  // another stream (callers run the panel chain and the trailing updates on different streams)
  if (h->mg_world == 1) return 0;
  GPB_CUDA(h, cudaStreamWaitEvent(h->cstream, e, 0));
  GPB_NCCL(h, g_nccl.Broadcast(buf, buf, (size_t)count, NCCL_DOUBLE, root, h->nccl_comm, h->cstream));
  GPB_CUDA(h, cudaEventRecord(e, h->cstream));
  return 0;
}

extern "C" int gpb200_mg_wait(gpb200_handle_t h, long long ticket) {
  CHECK_H(h);
  if (ticket < 0 || ticket >= h->mg_tickets || h->mg_tickets - ticket > MG_EVENT_RING)
    BAD_ARG(h, 2, "mg_wait: unknown or expired ticket");
  GPB_CUDA(h, cudaStreamWaitEvent(h->stream, h->mg_events[(size_t)(ticket % MG_EVENT_RING)], 0));
  return 0;
}

// In-place all-reduce on the compute stream (op 0 = sum, 1 = max); is_int selects int32 data.
extern "C" int gpb200_mg_allreduce(gpb200_handle_t h, void *buf, long long count, int op, int is_int) {
  CHECK_H(h);
  if (h->mg_world == 1) return 0;
  if (!h->nccl_comm) BAD_ARG(h, 1, "mg_allreduce: gpb200_mg_comm_init first");
  GPB_NCCL(h, g_nccl.AllReduce(buf, buf, (size_t)count, is_int ? NCCL_INT32 : NCCL_DOUBLE, op == 1 ? NCCL_MAX : NCCL_SUM,
                               h->nccl_comm, h->stream));
  return 0;
}

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

namespace {
// one tile column jl of the in-panel left-looking factorisation: update with the panel's columns to its left, POTRF, TRSM
int mg_factor_col(Handle *h, int n, int col0, int ncols, double *P, long long ldp, int jl, int *info_dev) {
  const int np = round_up(n, TILE);
  const int ntp = ncols / TILE, nrt = (np - col0) / TILE;
  TaskList tl;
  const long long key = mgkey(TK_MG_FACTOR, nrt, ntp, 0, 0);
  if (!cached(h, key, &tl)) {
    std::vector<TileTask> t;
    std::vector<int> off(1, 0);
    for (int j = 0; j < ntp; j++) {
      if (j > 0)
        for (int i = j; i < nrt; i++) t.push_back({i * TILE, 0, j * TILE, 0, i * TILE, j * TILE, j * TILE, i == j});
      off.push_back((int)t.size());
    }
    RC(upload_tasks(h, key, t, off, &tl));
  }
  if (tl.count(jl) > 0) {
    GemmParams p{};
    p.A = mref(P, ldp, 0);
    p.B = mref(P, ldp, 0);
    p.C = mref(P, ldp, 0);
    p.C0 = mref(P, ldp, 0);
    p.alpha = -1.0;
    p.beta = 1.0;
    p.tasks = tl.at(jl);
    RC(launch_gemm(h, LAYOUT_NT, EPI_AXPBY, p, tl.count(jl), 1));
  }
  const long long doff = (long long)jl * TILE * (ldp + 1);
  return launch_potrf_trsm_at(h, P, ldp, 0, doff, col0 + jl * TILE, n, doff + TILE, nrt - 1 - jl, 1, info_dev);
}
}  // namespace

extern "C" int gpb200_mg_panel_factor(gpb200_handle_t h, int n, int col0, int ncols, double *P, long long ldp,
                                      int *info_dev) {
  CHECK_H(h);
  int np;
  RC(mg_check_panel(h, n, col0, ncols, ldp, &np));
  for (int jl = 0; jl < ncols / TILE; jl++) RC(mg_factor_col(h, n, col0, ncols, P, ldp, jl, info_dev));
  return 0;
}

// The same, one 128-wide tile column at a time (jl = 0 .. ncols/128 - 1, in order): a finished tile column can be
// handed to gpb200_mg_bcast while the next one is still being factored.
extern "C" int gpb200_mg_panel_factor_col(gpb200_handle_t h, int n, int col0, int ncols, double *P, long long ldp, int jl,
                                          int *info_dev) {
  CHECK_H(h);
  int np;
  RC(mg_check_panel(h, n, col0, ncols, ldp, &np));
  if (jl < 0 || jl >= ncols / TILE) BAD_ARG(h, 7, "mg_panel_factor_col: tile column outside the panel");
  return mg_factor_col(h, n, col0, ncols, P, ldp, jl, info_dev);
}

// Apply a factored panel to tile columns [jl0, jl1) of a panel right of it: C -= P_rows * P_cols^T (NT, K = pncols).
// The column range lets the owner of the next panel update its first tile column alone on the panel chain and the other
// ones beside the chain (block_cyclic.py::_factor_native).
extern "C" int gpb200_mg_panel_update_cols(gpb200_handle_t h, int n, int pcol0, int pncols, const double *P, long long ldp,
                                           int ccol0, int cncols, double *Cp, long long ldc, int jl0, int jl1) {
  CHECK_H(h);
  int np;
  RC(mg_check_panel(h, n, pcol0, pncols, ldp, &np));
  RC(mg_check_panel(h, n, ccol0, cncols, ldc, &np));
  if (ccol0 < pcol0 + pncols) BAD_ARG(h, 7, "mg_panel_update: the target panel must lie right of the source panel");
  const int nt = np / TILE, d = (ccol0 - pcol0) / TILE, cnt = cncols / TILE, crt = nt - ccol0 / TILE, pk = pncols / TILE;
  if (jl0 < 0 || jl1 > cnt || jl0 >= jl1) BAD_ARG(h, 11, "mg_panel_update_cols: empty or out-of-range tile-column range");
  TaskList tl;
  const long long key = mgkey(TK_MG_UPDATE, d, cnt, crt, pk);
  if (!cached(h, key, &tl)) {
    std::vector<TileTask> t;
    std::vector<int> off(1, 0);
    for (int jl = 0; jl < cnt; jl++) {
      for (int il = jl; il < crt; il++)  // il, jl: tile coordinates local to the target panel
        t.push_back({(il + d) * TILE, 0, (jl + d) * TILE, 0, il * TILE, jl * TILE, pk * TILE, il == jl});
      off.push_back((int)t.size());
    }
    RC(upload_tasks(h, key, t, off, &tl));
  }
  int ntasks = 0;
  for (int jl = jl0; jl < jl1; jl++) ntasks += tl.count(jl);
  if (ntasks == 0) return 0;
  GemmParams p{};
  p.A = mref(const_cast<double *>(P), ldp, 0);
  p.B = mref(const_cast<double *>(P), ldp, 0);
  p.C = mref(Cp, ldc, 0);
  p.C0 = mref(Cp, ldc, 0);
  p.alpha = -1.0;
  p.beta = 1.0;
  p.tasks = tl.at(jl0);
  return launch_gemm(h, LAYOUT_NT, EPI_AXPBY, p, ntasks, 1);
}

extern "C" int gpb200_mg_panel_update(gpb200_handle_t h, int n, int pcol0, int pncols, const double *P, long long ldp,
                                      int ccol0, int cncols, double *Cp, long long ldc) {
  return gpb200_mg_panel_update_cols(h, n, pcol0, pncols, P, ldp, ccol0, cncols, Cp, ldc, 0, cncols / TILE);
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

// =================================================================================================
// (e) distributed GRADIENT of config 5: per-rank building blocks, no communication inside.
//
// After the block-cyclic factorisation every rank holds every panel of L (its own and the broadcast copies).
// The gradient needs K^-1 = W^T W with W = L^-1, but only through  tr(K^-1 dK) = sum_k w_k dK w_k^T  over the ROWS w_k
// of W -- and row k of W is column k of X = L^-T, which depends on L alone.  So rank r computes the columns of X that
// belong to ITS panels (a backward substitution with identity right-hand side, N^3 / 3P flops, no exchange), and the
// symmetric product G_r = X_r X_r^T of those columns feeds the fused trace epilogue tile by tile (another N^3 / 3P);
// the partial sums of all ranks add up to the sums of the single-GPU path.  Likewise z_k = x_k^T y for the rank's own k
// and a = sum_r X_r z_r.  The only collectives are two small all-reduces (the caller's: gpb200_mg_allreduce).
//
//   Lsq  np x np (ld = np)   the full lower factor, assembled panel by panel with gpb200_mg_panel_to_square
//   Xp   np x nmine (ld = np) the rank's columns of X, packed panel after panel (nmine = sum of its panels' widths)
//   S    pc x nmine,  Wd  2 x npanels x pc x pc   scratch (one solved row panel; the inverted diagonal blocks)
// =================================================================================================
namespace {
enum { TK_MG_XSOLVE = 42, TK_MG_XUPDATE = 43, TK_MG_XTRACE = 44 };

// Which columns of X a rank computes is free (every rank holds all of L): panels are dealt out in SNAKE order
// (0..P-1, P-1..0, 0..P-1, ...) because the cost of a column panel grows with the square of its index -- the plain
// cyclic deal of the factorisation gives the last rank 39 % more inverse work than the first at 8 ranks x 64 panels.
__host__ __device__ inline int mg_snake_panel(int q, int rank, int world) { return q * world + ((q & 1) ? world - 1 - rank : rank); }

struct MgGeom {
  int np, pc, npanels, rank, world, nq;
  long long nmine;
  int ncols(int p) const { return std::min(pc, np - p * pc); }
  int panel_of(int q) const { return mg_snake_panel(q, rank, world); }
  int first_q_at_or_after(int P) const {   // smallest q with panel_of(q) >= P (panel_of is increasing in q)
    int q = P / world;
    if (panel_of(q) < P) q++;
    return q;
  }
};

int mg_geom(Handle *h, int n, int pc, int rank, int world, MgGeom *g) {
  if (n < 1 || pc < TILE || (pc % TILE) || world < 1 || rank < 0 || rank >= world) BAD_ARG(h, 2, "mg: bad n / panel width / rank / world");
  g->np = round_up(n, TILE);
  g->pc = std::min(pc, g->np);
  g->npanels = (g->np + g->pc - 1) / g->pc;
  g->rank = rank;
  g->world = world;
  g->nq = 0;
  g->nmine = 0;
  for (int q = 0; g->panel_of(q) < g->npanels; q++) { g->nq++; g->nmine += g->ncols(g->panel_of(q)); }
  return 0;
}

__global__ void mg_x_init_kernel(int np, int pc, int rank, int world, long long nmine, double *__restrict__ Xp) {
  const long long total = (long long)np * nmine;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long kk = e / np;
    const int i = (int)(e - kk * np);
    const int q = (int)(kk / pc);
    const long long gcol = (long long)mg_snake_panel(q, rank, world) * pc + (kk - (long long)q * pc);
    Xp[e] = (i == gcol) ? 1.0 : 0.0;
  }
}

// out[i] (+)= sum_{kk in chunk} V[i + kk * ldv] * z[kk]: blockIdx.x = 128-row strip, blockIdx.y = column chunk
__global__ void __launch_bounds__(128) mg_gemv_n_kernel(int rows, long long cols, const double *__restrict__ V, long long ldv,
                                                        const double *__restrict__ z, double *__restrict__ part) {
  const int i = blockIdx.x * 128 + threadIdx.x;
  const long long per = (cols + gridDim.y - 1) / gridDim.y;
  const long long k0 = blockIdx.y * per, k1 = min(cols, k0 + per);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  long long k = k0;
  if (i < rows) {
    const double *vp = V + i + k0 * ldv;
    for (; k + 3 < k1; k += 4) {
      s0 = fma(vp[0], z[k], s0);
      s1 = fma(vp[ldv], z[k + 1], s1);
      s2 = fma(vp[2 * ldv], z[k + 2], s2);
      s3 = fma(vp[3 * ldv], z[k + 3], s3);
      vp += 4 * ldv;
    }
    for (; k < k1; k++) { s0 = fma(vp[0], z[k], s0); vp += ldv; }
    part[(long long)blockIdx.y * rows + i] = (s0 + s1) + (s2 + s3);
  }
}
__global__ void mg_sum_parts_kernel(int rows, int nparts, const double *__restrict__ part, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  double s = 0.0;
  for (int c = 0; c < nparts; c++) s += part[(long long)c * rows + i];
  out[i] = s;
}
// sums[0] = sum_k z_k^2 ; sums[1] = sum_{i < n} log L[i][i]   (one CTA, deterministic)
__global__ void __launch_bounds__(256) mg_qf_logdet_kernel(int n, int np, long long nz, const double *__restrict__ z,
                                                           const double *__restrict__ Lsq, double *__restrict__ sums) {
  __shared__ double red[8][2];
  double q = 0.0, ld = 0.0;
  for (long long k = threadIdx.x; k < nz; k += 256) q = fma(z[k], z[k], q);
  for (int i = threadIdx.x; i < n; i += 256) ld += log(Lsq[i + (long long)i * np]);
  for (int o = 16; o > 0; o >>= 1) { q += __shfl_xor_sync(0xffffffffu, q, o); ld += __shfl_xor_sync(0xffffffffu, ld, o); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = q; red[threadIdx.x >> 5][1] = ld; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < 8; w++) { a += red[w][0]; b += red[w][1]; }
    sums[0] = a; sums[1] = b;
  }
}
// part[blockIdx.x * 2 + {0,1}] = sum_{i in this CTA's 128 rows} sum_{j < n} a_i a_j e_ij {1, d_ij^2},  e = exp(-d^2 / 2 rho^2)
__global__ void __launch_bounds__(256) mg_quadform_kernel(int n, int row_lo, int row_hi, const double *__restrict__ x,
                                                          const double *__restrict__ a, const double *__restrict__ theta3,
                                                          double *__restrict__ part) {
  __shared__ double xs[128], as_[128], red[8][2];
  const int i0 = row_lo + blockIdx.x * 128, tid = threadIdx.x;
  if (tid < 128) {
    const int i = i0 + tid;
    xs[tid] = (i < row_hi) ? x[i] : 0.0;
    as_[tid] = (i < row_hi) ? a[i] : 0.0;
  }
  __syncthreads();
  const double rho = theta3[1], nh = -0.5 / (rho * rho);
  double s0 = 0.0, s1 = 0.0;
  for (int j = tid; j < n; j += 256) {
    const double xj = x[j], aj = a[j];
    double t0 = 0.0, t1 = 0.0;
#pragma unroll 4
    for (int r = 0; r < 128; r++) {
      const double d = xs[r] - xj, d2 = d * d;
      const double w = as_[r] * exp_nonpos(d2 * nh);
      t0 += w;
      t1 = fma(w, d2, t1);
    }
    s0 = fma(aj, t0, s0);
    s1 = fma(aj, t1, s1);
  }
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  if ((tid & 31) == 0) { red[tid >> 5][0] = s0; red[tid >> 5][1] = s1; }
  __syncthreads();
  if (tid == 0) {
    double u = 0, v = 0;
    for (int w = 0; w < 8; w++) { u += red[w][0]; v += red[w][1]; }
    part[blockIdx.x * 2] = u;
    part[blockIdx.x * 2 + 1] = v;
  }
}
__global__ void __launch_bounds__(256) mg_sum2_kernel(int nrec, const double *__restrict__ part, double *__restrict__ out2) {
  __shared__ double red[8][2];
  double s0 = 0, s1 = 0;
  for (int r = threadIdx.x; r < nrec; r += 256) { s0 += part[2 * r]; s1 += part[2 * r + 1]; }
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = s0; red[threadIdx.x >> 5][1] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double u = 0, v = 0;
    for (int w = 0; w < 8; w++) { u += red[w][0]; v += red[w][1]; }
    out2[0] = u; out2[1] = v;
  }
}
// sums3[c] = sum over records of partial[rec * 4 + c], c = 0..2 (one CTA, fixed order)
__global__ void __launch_bounds__(256) mg_trace_sum_kernel(long long nrec, const double *__restrict__ partial, double *__restrict__ sums3) {
  __shared__ double red[8][3];
  double s[3] = {0, 0, 0};
  for (long long r = threadIdx.x; r < nrec; r += 256)
    for (int c = 0; c < 3; c++) s[c] += partial[r * 4 + c];
  for (int c = 0; c < 3; c++)
    for (int o = 16; o > 0; o >>= 1) s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
  if ((threadIdx.x & 31) == 0) for (int c = 0; c < 3; c++) red[threadIdx.x >> 5][c] = s[c];
  __syncthreads();
  if (threadIdx.x < 3) {
    double a = 0;
    for (int w = 0; w < 8; w++) a += red[w][threadIdx.x];
    sums3[threadIdx.x] = a;
  }
}
}  // namespace

extern "C" int gpb200_mg_panel_to_square(gpb200_handle_t h, int n, int col0, int ncols, const double *P, long long ldp, double *Lsq) {
  CHECK_H(h);
  int np;
  RC(mg_check_panel(h, n, col0, ncols, ldp, &np));
  GPB_CUDA(h, cudaMemcpy2DAsync(Lsq + col0 + (long long)col0 * np, (size_t)np * 8, P, (size_t)ldp * 8, (size_t)(np - col0) * 8, ncols,
                               cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}

extern "C" long long gpb200_mg_my_columns(int n, int pc, int rank, int world) {
  gpb200_handle_s tmp;
  MgGeom g;
  if (mg_geom(&tmp, n, pc, rank, world, &g)) return -1;
  return g.nmine;
}

extern "C" int gpb200_mg_inverse_rows(gpb200_handle_t h, int n, int pc, int rank, int world, const double *Lsq, double *Xp,
                                      double *S, double *Wd) {
  CHECK_H(h);
  MgGeom g;
  RC(mg_geom(h, n, pc, rank, world, &g));
  if (g.nmine == 0) return 0;
  const int np = g.np;
  pc = g.pc;
  const long long blk = (long long)pc * pc;
  // inverted diagonal blocks of every panel (each rank needs all of them; N pc^2 / 3 flops in total)
  for (int K = 0; K < g.npanels; K++) {
    const int bk = g.ncols(K);
    RC(launch_pack(h, bk, bk, Lsq + (long long)K * pc * (np + 1), np, pc, pc, Wd + K * blk, 2, 0.0));
  }
  RC(trtri_batched(h, Wd, Wd + (long long)g.npanels * blk, pc, blk, g.npanels));
  {
    ProfScope ps__(h, PC_OTHER);
    mg_x_init_kernel<<<148 * 8, 256, 0, h->stream>>>(np, pc, rank, world, g.nmine, Xp);
    GPB_LAUNCH_CHECK(h);
  }
  const int pt = pc / TILE;
  for (int K = g.npanels - 1; K >= 0; K--) {
    const int qK = g.first_q_at_or_after(K);
    const long long nact = g.nmine - (long long)qK * pc;
    if (nact <= 0) continue;
    const int bk = g.ncols(K), mt = bk / TILE, ntl = (int)(nact / TILE);
    const long long cK = (long long)K * pc;
    double *Xact = Xp + (long long)qK * pc * np;  // first active column
    {  // S = Wd_K^T R[K rows]   (TN; Wd_K lower triangular: k >= m)
      TaskList tl;
      const long long key = mgkey(TK_MG_XSOLVE, mt, ntl & 0x1fff, ntl >> 13, 0);
      if (!cached(h, key, &tl)) {
        std::vector<TileTask> t;
        for (int mi = 0; mi < mt; mi++)
          for (int nj = 0; nj < ntl; nj++)
            t.push_back({mi * TILE, mi * TILE, mi * TILE, nj * TILE, mi * TILE, nj * TILE, (mt - mi) * TILE, TF_A_TRI_FIRST});
        sort_desc(t, 0);
        std::vector<int> off = {0, (int)t.size()};
        RC(upload_tasks(h, key, t, off, &tl));
      }
      GemmParams p{};
      p.A = mref(Wd + K * blk, pc, 0);
      p.B = mref(Xact + cK, np, 0);
      p.C = mref(S, pc, 0);
      p.alpha = 1.0;
      p.tasks = tl.at(0);
      RC(launch_gemm(h, LAYOUT_TN, EPI_AXPBY, p, tl.count(0), 1));
    }
    GPB_CUDA(h, cudaMemcpy2DAsync(Xact + cK, (size_t)np * 8, S, (size_t)pc * 8, (size_t)bk * 8, (size_t)nact, cudaMemcpyDeviceToDevice,
                                 h->stream));
    if (K > 0) {  // R[rows above] -= L[K rows, cols above]^T X[K rows]   (TN, contraction over the panel's rows)
      const int it = (int)(cK / TILE);
      TaskList tl;
      const long long key = mgkey(TK_MG_XUPDATE, it, ntl & 0x1fff, mt, ntl >> 13);
      if (!cached(h, key, &tl)) {
        std::vector<TileTask> t;
        t.reserve((size_t)it * ntl);
        for (int i = 0; i < it; i++)
          for (int nj = 0; nj < ntl; nj++) t.push_back({0, i * TILE, 0, nj * TILE, i * TILE, nj * TILE, bk, 0});
        std::vector<int> off = {0, (int)t.size()};
        RC(upload_tasks(h, key, t, off, &tl));
      }
      GemmParams p{};
      p.A = mref(const_cast<double *>(Lsq) + cK, np, 0);
      p.B = mref(S, pc, 0);
      p.C = mref(Xact, np, 0);
      p.C0 = mref(Xact, np, 0);
      p.alpha = -1.0;
      p.beta = 1.0;
      p.tasks = tl.at(0);
      RC(launch_gemm(h, LAYOUT_TN, EPI_AXPBY, p, tl.count(0), 1));
    }
  }
  (void)pt;
  return 0;
}

// z_mine = X_r^T y (length nmine), sums = (sum z_mine^2, log det L), a_part = X_r z_mine (length np).
// ypad: y followed by zeros up to np.  part: scratch of 8 * np doubles.
extern "C" int gpb200_mg_solve_partials(gpb200_handle_t h, int n, int pc, int rank, int world, const double *Lsq, const double *Xp,
                                        const double *ypad, double *z_mine, double *a_part, double *sums2, double *part) {
  CHECK_H(h);
  MgGeom g;
  RC(mg_geom(h, n, pc, rank, world, &g));
  const int np = g.np;
  if (g.nmine > 0) RC(launch_gemv_t(h, np, (int)g.nmine, Xp, np, ypad, nullptr, z_mine));
  ProfScope ps__(h, PC_SOLVE);
  const int ks = 8;
  if (g.nmine > 0) {
    dim3 grid(np / 128, ks);
    mg_gemv_n_kernel<<<grid, 128, 0, h->stream>>>(np, g.nmine, Xp, np, z_mine, part);
    GPB_LAUNCH_CHECK(h);
    mg_sum_parts_kernel<<<(np + 255) / 256, 256, 0, h->stream>>>(np, ks, part, a_part);
    GPB_LAUNCH_CHECK(h);
  } else {
    GPB_CUDA(h, cudaMemsetAsync(a_part, 0, sizeof(double) * np, h->stream));
  }
  mg_qf_logdet_kernel<<<1, 256, 0, h->stream>>>(n, np, g.nmine, z_mine, Lsq, sums2);
  GPB_LAUNCH_CHECK(h);
  return 0;
}

// sums3 = (sum M e, sum M e d^2, tr G_r) with G_r = X_r X_r^T and M = avec avec^T - G_r over i, j < n
// (avec: a = K^-1 y on ONE rank and zeros on the others, so that the a a^T term is counted once).
// theta3: device (alpha, rho, sigma).  partial: scratch of 16 * ntile * (ntile + 1) / 2 doubles, ntile = np / 128.
extern "C" int gpb200_mg_trace_partials(gpb200_handle_t h, int n, int pc, int rank, int world, const double *Xp, const double *x,
                                        const double *avec, const double *theta3, double *partial, double *sums3) {
  CHECK_H(h);
  MgGeom g;
  RC(mg_geom(h, n, pc, rank, world, &g));
  const int np = g.np, nt = np / TILE;
  pc = g.pc;
  TaskList tl;
  const long long key = mgkey(TK_MG_XTRACE, nt, pc / TILE, rank, world);
  if (!cached(h, key, &tl)) {
    std::vector<TileTask> t;
    for (int I = 0; I < nt; I++) {
      const int PI = (I * TILE) / pc;                       // panel that holds the columns at row tile I
      const long long k0 = (long long)g.first_q_at_or_after(PI) * pc;  // first packed column that can be non-zero in row tile I
      const long long klen = g.nmine - k0;
      if (klen <= 0) continue;
      for (int J = 0; J <= I; J++)
        t.push_back({I * TILE, (int)k0, J * TILE, (int)k0, I * TILE, J * TILE, (int)klen, I == J ? TF_DIAG : 0});
    }
    sort_desc(t, 0);
    std::vector<int> off = {0, (int)t.size()};
    RC(upload_tasks(h, key, t, off, &tl));
  }
  const int ntasks = tl.count(0);
  if (ntasks == 0) {
    GPB_CUDA(h, cudaMemsetAsync(sums3, 0, 3 * sizeof(double), h->stream));
    return 0;
  }
  GemmParams p{};
  p.A = mref(const_cast<double *>(Xp), np, 0);
  p.B = mref(const_cast<double *>(Xp), np, 0);
  p.C = mref(nullptr, np, 0);
  p.tasks = tl.at(0);
  p.x = x; p.x_stride = 0;
  p.avec = avec; p.a_stride = 0;
  p.theta = theta3;
  p.partial = partial;
  p.n = n;
  p.ntasks = ntasks;
  RC(launch_gemm(h, LAYOUT_NT, EPI_TRACE, p, ntasks, 1));
  ProfScope ps__(h, PC_OTHER);
  mg_trace_sum_kernel<<<1, 256, 0, h->stream>>>((long long)ntasks * gemm_nsplit(h, ntasks, 1), partial, sums3);
  GPB_LAUNCH_CHECK(h);
  return 0;
}

// The a a^T half of the gradient, a = K^-1 y:  sums2 = (a^T E a, a^T (E o D2) a) restricted to this rank's slice of ROWS
// (a contiguous n / world share), E_ij = exp(-d_ij^2 / 2 rho^2): N^2 / P kernel evaluations, summed over ranks by the caller.
// part: scratch of 2 * ceil(n / 128) doubles.
extern "C" int gpb200_mg_quadform_partials(gpb200_handle_t h, int n, int rank, int world, const double *x, const double *a,
                                           const double *theta3, double *part, double *sums2) {
  CHECK_H(h);
  if (n < 1 || world < 1 || rank < 0 || rank >= world) BAD_ARG(h, 2, "mg_quadform_partials: bad n / rank / world");
  const int lo = (int)((long long)n * rank / world), hi = (int)((long long)n * (rank + 1) / world);
  const int nb = (hi - lo + 127) / 128;
  ProfScope ps__(h, PC_OTHER);
  if (nb > 0) {
    mg_quadform_kernel<<<nb, 256, 0, h->stream>>>(n, lo, hi, x, a, theta3, part);
    GPB_LAUNCH_CHECK(h);
  }
  mg_sum2_kernel<<<1, 256, 0, h->stream>>>(nb, part, sums2);
  GPB_LAUNCH_CHECK(h);
  return 0;
}
