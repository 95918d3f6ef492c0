// C-ABI entry points of libgpb200.so (see include/gpb200.h) and the host-side orchestration of the
// tiled algorithms: left-looking batched Cholesky, in-place recursive triangular inverse, fused
// LAUUM + trace, forward-mode tangent, conditioning.  Everything O(N^3) goes through the DMMA tile
// GEMM (gemm.cu); everything here is launch sequencing, task-list construction and workspace
// management.  No cuBLAS / cuSOLVER, no CPU fallback.
#include "../../include/gpb200.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <new>

#include "common.cuh"
#include "gram.cuh"

using namespace gpb;

struct gpb200_handle_s : public Handle {};

namespace {

// ---------------------------------------------------------------------------------------------
// workspace arena
// ---------------------------------------------------------------------------------------------
struct Arena {
  char *base = nullptr;
  size_t cap = 0, off = 0;
  template <typename T>
  T *take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    if (off + bytes > cap) return nullptr;
    T *p = reinterpret_cast<T *>(base + off);
    off += bytes;
    return p;
  }
};

size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

int ws_reserve(Handle *h, size_t bytes, Arena *a) {
  if (bytes > h->ws_bytes) {
    if (h->ws) {
      GPB_CUDA(h, cudaStreamSynchronize(h->stream));
      if (h->gstream) GPB_CUDA(h, cudaStreamSynchronize(h->gstream));
      for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);  // they bake in workspace addresses
      h->graphs.clear();
      GPB_CUDA(h, cudaFree(h->ws));
      h->ws = nullptr;
      h->ws_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&h->ws, bytes);
    if (e != cudaSuccess) {
      snprintf(h->err, sizeof(h->err), "workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
      (void)cudaGetLastError();
      return -1002;
    }
    h->ws_bytes = bytes;
  }
  a->base = reinterpret_cast<char *>(h->ws);
  a->cap = h->ws_bytes;
  a->off = 0;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// task lists (host-built once per tile count, cached on the device)
// ---------------------------------------------------------------------------------------------
enum TaskKind { TK_CHOL = 1, TK_CHOL_TRAIL, TK_TRTRI_S, TK_TRTRI_W, TK_LAUUM, TK_TAN_T1, TK_TAN_A, TK_TAN_LDOT, TK_COND_V,
                TK_COND_COV, TK_MUL_WB, TK_MUL_WTB };

struct TaskList {
  const TileTask *dev = nullptr;
  const std::vector<int> *offsets = nullptr;
  int count(int step) const { return (*offsets)[step + 1] - (*offsets)[step]; }
  const TileTask *at(int step) const { return dev + (*offsets)[step]; }
  int steps() const { return (int)offsets->size() - 1; }
};

int upload_tasks(Handle *h, long long key, const std::vector<TileTask> &tasks, const std::vector<int> &offsets,
                 TaskList *out) {
  auto it = h->task_cache.find(key);
  if (it == h->task_cache.end()) {
    TileTask *dev = nullptr;
    const size_t bytes = std::max<size_t>(tasks.size(), 1) * sizeof(TileTask);
    GPB_CUDA(h, cudaMalloc(&dev, bytes));
    if (!tasks.empty())
      GPB_CUDA(h, cudaMemcpyAsync(dev, tasks.data(), tasks.size() * sizeof(TileTask), cudaMemcpyHostToDevice, h->stream));
    GPB_CUDA(h, cudaStreamSynchronize(h->stream));
    it = h->task_cache.emplace(key, std::make_pair(dev, offsets)).first;
  }
  out->dev = it->second.first;
  out->offsets = &it->second.second;
  return 0;
}

bool cached(Handle *h, long long key, TaskList *out) {
  auto it = h->task_cache.find(key);
  if (it == h->task_cache.end()) return false;
  out->dev = it->second.first;
  out->offsets = &it->second.second;
  return true;
}

long long tkey(int kind, int a, int b = 0) { return ((long long)kind << 48) | ((long long)a << 24) | (long long)b; }

void sort_desc(std::vector<TileTask> &v, size_t from) {
  std::stable_sort(v.begin() + from, v.end(), [](const TileTask &x, const TileTask &y) { return x.k_len > y.k_len; });
}

// Cholesky task lists.  Panels of `pt` tile columns: inside a panel the factorisation is
// left-looking (step j updates block column j with the panel's columns to its left, K = (j-p0)*128);
// after a panel is complete one right-looking launch applies it to everything to its right
// (K = pt*128).  pt = nt is the pure left-looking algorithm used for large batches; a narrow panel
// gives a single large matrix enough tiles per launch to fill 148 SMs.
int tasks_chol(Handle *h, int nt, int pt, TaskList *upd, TaskList *trail) {
  const long long k1 = tkey(TK_CHOL, nt, pt), k2 = tkey(TK_CHOL_TRAIL, nt, pt);
  if (cached(h, k1, upd) && cached(h, k2, trail)) return 0;
  std::vector<TileTask> t, tt;
  std::vector<int> off(1, 0), offt(1, 0);
  for (int j = 0; j < nt; j++) {
    const int p0 = (j / pt) * pt;
    if (j > p0)
      for (int i = j; i < nt; i++)
        t.push_back({i * TILE, p0 * TILE, j * TILE, p0 * TILE, i * TILE, j * TILE, (j - p0) * TILE, i == j});
    off.push_back((int)t.size());
  }
  for (int p0 = 0; p0 < nt; p0 += pt) {
    const int p1 = std::min(nt, p0 + pt);
    for (int jj = p1; jj < nt; jj++)
      for (int i = jj; i < nt; i++)
        tt.push_back({i * TILE, p0 * TILE, jj * TILE, p0 * TILE, i * TILE, jj * TILE, (p1 - p0) * TILE, i == jj});
    offt.push_back((int)tt.size());
  }
  int rc = upload_tasks(h, k1, t, off, upd);
  if (rc) return rc;
  return upload_tasks(h, k2, tt, offt, trail);
}

struct Node { int lo, mid, hi, level; };
int build_nodes(int lo, int hi, std::vector<Node> &nodes) {
  if (hi - lo <= 1) return 0;
  const int mid = lo + (hi - lo + 1) / 2;
  const int l1 = build_nodes(lo, mid, nodes), l2 = build_nodes(mid, hi, nodes);
  const int lvl = std::max(l1, l2) + 1;
  nodes.push_back({lo, mid, hi, lvl});
  return lvl;
}

// in-place recursive inverse of the lower-triangular factor: per node
//   S   = W22 * L21          (scratch buffer)        S[i,j] = sum_{k=mid..i} W[i,k] L[k,j]
//   W21 = -S * W11           (over L21 in place)     W[i,j] = -sum_{k=j..mid-1} S[i,k] W[k,j]
int tasks_trtri(Handle *h, int nt, TaskList *s_out, TaskList *w_out) {
  const long long ks = tkey(TK_TRTRI_S, nt), kw = tkey(TK_TRTRI_W, nt);
  if (cached(h, ks, s_out) && cached(h, kw, w_out)) return 0;
  std::vector<Node> nodes;
  const int top = build_nodes(0, nt, nodes);
  std::vector<TileTask> ts, tw;
  std::vector<int> os(1, 0), ow(1, 0);
  for (int lvl = 1; lvl <= top; lvl++) {
    const size_t fs = ts.size(), fw = tw.size();
    for (const Node &nd : nodes) {
      if (nd.level != lvl) continue;
      for (int i = nd.mid; i < nd.hi; i++)
        for (int j = nd.lo; j < nd.mid; j++) {
          ts.push_back({i * TILE, nd.mid * TILE, nd.mid * TILE, j * TILE, i * TILE, j * TILE, (i - nd.mid + 1) * TILE, 0});
          tw.push_back({i * TILE, j * TILE, j * TILE, j * TILE, i * TILE, j * TILE, (nd.mid - j) * TILE, 0});
        }
    }
    sort_desc(ts, fs);
    sort_desc(tw, fw);
    os.push_back((int)ts.size());
    ow.push_back((int)tw.size());
  }
  int rc = upload_tasks(h, ks, ts, os, s_out);
  if (rc) return rc;
  return upload_tasks(h, kw, tw, ow, w_out);
}

// G = W^T W, lower tiles: G[i,j] = sum_{k>=i} W[k,i]^T W[k,j]
int tasks_lauum(Handle *h, int nt, TaskList *out) {
  const long long key = tkey(TK_LAUUM, nt);
  if (cached(h, key, out)) return 0;
  std::vector<TileTask> t;
  for (int i = 0; i < nt; i++)
    for (int j = 0; j <= i; j++) t.push_back({i * TILE, i * TILE, i * TILE, j * TILE, i * TILE, j * TILE, (nt - i) * TILE, i == j});
  sort_desc(t, 0);
  std::vector<int> off = {0, (int)t.size()};
  return upload_tasks(h, key, t, off, out);
}

// ---------------------------------------------------------------------------------------------
// engines on padded device buffers
// ---------------------------------------------------------------------------------------------
MatRef mref(double *p, long long ld, long long stride) { return MatRef{p, ld, stride}; }

// Batched tiled Cholesky, in place on Lbuf (np x np per item, lower tiles valid on entry).
// Panel width: pure left-looking when the batch alone fills the GPU, 8-tile (1024-column) panels
// with right-looking trailing updates otherwise.
int chol_panel_tiles(int nt, int batch) {
  if (batch >= 64 || nt <= 8) return nt;
  return 8;
}

int chol_batched(Handle *h, double *Lbuf, int np, long long stride, int n, int batch, int *info_dev, double *dvec) {
  const int nt = np / TILE;
  const int pt = h->chol_panel_override > 0 ? std::min(nt, h->chol_panel_override) : chol_panel_tiles(nt, batch);
  TaskList tl, tr;
  int rc = tasks_chol(h, nt, pt, &tl, &tr);
  if (rc) return rc;
  GemmParams p{};
  p.A = mref(Lbuf, np, stride);
  p.B = mref(Lbuf, np, stride);
  p.C = mref(Lbuf, np, stride);
  p.C0 = mref(Lbuf, np, stride);
  p.alpha = -1.0;
  p.beta = 1.0;
  for (int j = 0; j < nt; j++) {
    if (tl.count(j) > 0) {
      p.tasks = tl.at(j);
      rc = launch_gemm(h, LAYOUT_NT, EPI_AXPBY, p, tl.count(j), batch);
      if (rc) return rc;
    }
    rc = launch_potrf_tile(h, Lbuf, np, stride, j, n, batch, info_dev);
    if (rc) return rc;
    rc = launch_trsm_tiles(h, Lbuf, np, stride, j, nt - 1 - j, batch);
    if (rc) return rc;
    if ((j + 1) % pt == 0 && j + 1 < nt) {
      const int panel = j / pt;
      p.tasks = tr.at(panel);
      rc = launch_gemm(h, LAYOUT_NT, EPI_AXPBY, p, tr.count(panel), batch);
      if (rc) return rc;
    }
  }
  (void)dvec;
  return 0;
}

// In-place inverse of the lower-triangular factor held in Lbuf; Sbuf is same-shaped scratch.
int trtri_batched(Handle *h, double *Lbuf, double *Sbuf, int np, long long stride, int batch) {
  const int nt = np / TILE;
  int rc = launch_tile_inverse(h, Lbuf, Lbuf, np, stride, nt, batch);
  if (rc) return rc;
  if (nt == 1) return 0;
  TaskList ts, tw;
  rc = tasks_trtri(h, nt, &ts, &tw);
  if (rc) return rc;
  for (int lvl = 0; lvl < ts.steps(); lvl++) {
    GemmParams p{};
    p.A = mref(Lbuf, np, stride);
    p.B = mref(Lbuf, np, stride);
    p.C = mref(Sbuf, np, stride);
    p.alpha = 1.0;
    p.beta = 0.0;
    p.tasks = ts.at(lvl);
    rc = launch_gemm(h, LAYOUT_NN, EPI_AXPBY, p, ts.count(lvl), batch);
    if (rc) return rc;
    GemmParams q{};
    q.A = mref(Sbuf, np, stride);
    q.B = mref(Lbuf, np, stride);
    q.C = mref(Lbuf, np, stride);
    q.alpha = -1.0;
    q.beta = 0.0;
    q.tasks = tw.at(lvl);
    rc = launch_gemm(h, LAYOUT_NN, EPI_AXPBY, q, tw.count(lvl), batch);
    if (rc) return rc;
  }
  return 0;
}

__global__ void extract_diag_kernel(int np, const double *__restrict__ L, long long stride, double *__restrict__ dvec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) dvec[(long long)blockIdx.y * np + i] = L[(long long)blockIdx.y * stride + i + (long long)i * np];
}

int extract_diag(Handle *h, int np, const double *L, long long stride, double *dvec, int batch) {
  dim3 grid((np + 255) / 256, batch);
  ProfScope ps__(h, PC_OTHER);
  extract_diag_kernel<<<grid, 256, 0, h->stream>>>(np, L, stride, dvec);
  GPB_LAUNCH_CHECK(h);
  return 0;
}

// host<->device staging helpers -----------------------------------------------------------------
int to_device(Handle *h, const double *src, double *dev, size_t count) {
  GPB_CUDA(h, cudaMemcpyAsync(dev, src, count * sizeof(double), h->device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  return 0;
}
int from_device(Handle *h, const void *dev, void *dst, size_t bytes) {
  GPB_CUDA(h, cudaMemcpyAsync(dst, dev, bytes, h->device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  return 0;
}
int to_device_2d(Handle *h, const double *src, long long lds, double *dev, long long ldd, int rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  GPB_CUDA(h, cudaMemcpy2DAsync(dev, ldd * sizeof(double), src, lds * sizeof(double), rows * sizeof(double), cols,
                               h->device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
  return 0;
}
int from_device_2d(Handle *h, const double *dev, long long lds, double *dst, long long ldd, int rows, int cols) {
  if (rows <= 0 || cols <= 0) return 0;
  GPB_CUDA(h, cudaMemcpy2DAsync(dst, ldd * sizeof(double), dev, lds * sizeof(double), rows * sizeof(double), cols,
                               h->device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  return 0;
}
int finish(Handle *h) {
  if (!h->device_ptrs) GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int read_info(Handle *h, const int *info_dev, int *out) {
  GPB_CUDA(h, cudaMemcpyAsync(out, info_dev, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

#define CHECK_H(h)                   \
  do {                               \
    if (!(h)) return -1;             \
    (h)->err[0] = 0;                 \
    if (cudaSetDevice((h)->device) != cudaSuccess) return -1000; \
  } while (0)
#define BAD_ARG(h, k, msg)                                  \
  do {                                                      \
    snprintf((h)->err, sizeof((h)->err), "%s", msg);       \
    return -(k);                                            \
  } while (0)
#define RC(x)               \
  do {                      \
    int rc__ = (x);         \
    if (rc__) return rc__;  \
  } while (0)

}  // namespace

// =================================================================================================
// handle
// =================================================================================================
extern "C" int gpb200_version(void) { return 100; }

extern "C" int gpb200_create(gpb200_handle_t *out, int device) {
  if (!out) return -1;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
    (void)cudaGetLastError();
    return -1000;  // no CUDA device: there is no CPU fallback
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1000;
  if (prop.major < 10) return -1003;  // built for sm_100a only
  if (cudaSetDevice(device) != cudaSuccess) return -1000;
  gpb200_handle_s *h = new (std::nothrow) gpb200_handle_s();
  if (!h) return -1002;
  h->device = device;
  if (panel_smem_setup(h) || gemm_smem_setup(h)) { delete h; return -1000; }
  if (cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->g_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->g_out, cudaEventDisableTiming) != cudaSuccess) { delete h; return -1000; }
  const char *ng = getenv("GPB200_NO_GRAPH");
  if (ng && ng[0] == '1') h->graphs_enabled = 0;
  *out = h;
  return 0;
}

extern "C" int gpb200_destroy(gpb200_handle_t h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second.exec);
  if (h->gstream) { cudaStreamSynchronize(h->gstream); cudaStreamDestroy(h->gstream); }
  if (h->g_in) cudaEventDestroy(h->g_in);
  if (h->g_out) cudaEventDestroy(h->g_out);
  if (h->ws) cudaFree(h->ws);
  for (auto &kv : h->task_cache) cudaFree(kv.second.first);
  delete h;
  return 0;
}

extern "C" int gpb200_set_stream(gpb200_handle_t h, void *s) {
  if (!h) return -1;
  h->stream = reinterpret_cast<cudaStream_t>(s);
  return 0;
}
extern "C" int gpb200_set_pointer_mode(gpb200_handle_t h, int dev) {
  if (!h) return -1;
  h->device_ptrs = dev ? 1 : 0;
  return 0;
}
extern "C" int gpb200_synchronize(gpb200_handle_t h) {
  CHECK_H(h);
  GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}
extern "C" const char *gpb200_last_error(gpb200_handle_t h) { return h ? h->err : "null handle"; }
extern "C" long long gpb200_launch_count(gpb200_handle_t h) { return h ? h->launches : 0; }
extern "C" int gpb200_set_workspace_limit(gpb200_handle_t h, long long bytes) {
  if (!h) return -1;
  h->ws_limit = bytes;
  return 0;
}

// test/tuning knob: force the Cholesky panel width in 128-column tiles (0 = automatic)
extern "C" int gpb200_set_chol_panel_tiles(gpb200_handle_t h, int tiles) {
  if (!h || tiles < 0) return -1;
  h->chol_panel_override = tiles;
  return 0;
}

extern "C" int gpb200_set_profiling(gpb200_handle_t h, int on) {
  if (!h) return -1;
  h->profiling = on ? 1 : 0;
  h->prof.clear();
  h->ev_used = 0;
  if (on && h->ev_pool.size() < 1024) {
    // create the event pool up front: cudaEventCreate inside a timed region costs host time
    if (cudaSetDevice(h->device) != cudaSuccess) return -1000;
    while (h->ev_pool.size() < 1024) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) return -1000;
      h->ev_pool.push_back(e);
    }
  }
  return 0;
}

// Sums the event-bracketed launch durations per kernel class since profiling was switched on (or
// since the last call), then resets.  ms_out / count_out: arrays of 6 (gemm, potrf tile, trsm tile,
// gram, solves, other).  Synchronises the stream.
extern "C" int gpb200_get_profile(gpb200_handle_t h, double *ms_out, long long *count_out) {
  CHECK_H(h);
  GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  for (int c = 0; c < PC_COUNT; c++) { ms_out[c] = 0.0; count_out[c] = 0; }
  for (const auto &r : h->prof) {
    float ms = 0.f;
    GPB_CUDA(h, cudaEventElapsedTime(&ms, r.e0, r.e1));
    ms_out[r.cls] += ms;
    count_out[r.cls]++;
  }
  h->prof.clear();
  h->ev_used = 0;
  return 0;
}

// =================================================================================================
// a9 kernels
// =================================================================================================
extern "C" int gpb200_kernel_eval(gpb200_handle_t h, int kind, long long len, const double *tj, const double *tk,
                                  double amp2, double l, double *out) {
  CHECK_H(h);
  if (kind < 0 || kind > 9) BAD_ARG(h, 2, "kernel_eval: unknown kind");
  if (len < 0) BAD_ARG(h, 3, "kernel_eval: negative length");
  if (len == 0) return 0;
  if (h->device_ptrs) return launch_kernel_eval(h, kind, len, tj, tk, amp2, l, out);
  Arena a;
  RC(ws_reserve(h, 3 * pad256(len * sizeof(double)), &a));
  double *dj = a.take<double>(len), *dk = a.take<double>(len), *dout = a.take<double>(len);
  RC(to_device(h, tj, dj, len));
  RC(to_device(h, tk, dk, len));
  RC(launch_kernel_eval(h, kind, len, dj, dk, amp2, l, dout));
  RC(from_device(h, dout, out, len * sizeof(double)));
  return finish(h);
}

extern "C" int gpb200_gram_outer(gpb200_handle_t h, int kind, int n, int m, const double *x, const double *y,
                                 double amp2, double l, double *K, int ldk) {
  CHECK_H(h);
  if (kind < 0 || kind > 9) BAD_ARG(h, 2, "gram_outer: unknown kind");
  if (n < 0 || m < 0) BAD_ARG(h, 3, "gram_outer: negative size");
  if (ldk < std::max(1, n)) BAD_ARG(h, 10, "gram_outer: ldk < n");
  if (n == 0 || m == 0) return 0;
  if (h->device_ptrs) return launch_gram_outer(h, kind, n, m, x, y, amp2, l, K, ldk);
  Arena a;
  const int ldd = round_up(n, 2);
  RC(ws_reserve(h, pad256(n * 8) + pad256(m * 8) + pad256((size_t)ldd * m * 8), &a));
  double *dx = a.take<double>(n), *dy = a.take<double>(m), *dK = a.take<double>((size_t)ldd * m);
  RC(to_device(h, x, dx, n));
  RC(to_device(h, y, dy, m));
  RC(launch_gram_outer(h, kind, n, m, dx, dy, amp2, l, dK, ldd));
  RC(from_device_2d(h, dK, ldd, K, ldk, n, m));
  return finish(h);
}

extern "C" int gpb200_gram_ard(gpb200_handle_t h, int n, int m, int D, const double *X, int ldx, const double *Y,
                               int ldy, double alpha, const double *rho, double *K, int ldk) {
  CHECK_H(h);
  if (n < 0 || m < 0 || D < 1) BAD_ARG(h, 2, "gram_ard: bad sizes");
  if (ldx < std::max(1, n) || ldy < std::max(1, m) || ldk < std::max(1, n)) BAD_ARG(h, 6, "gram_ard: bad leading dimension");
  if (n == 0 || m == 0) return 0;
  if (h->device_ptrs) return launch_gram_ard(h, n, m, D, X, ldx, Y, ldy, alpha, rho, D, K, ldk);
  Arena a;
  RC(ws_reserve(h, pad256((size_t)n * D * 8) + pad256((size_t)m * D * 8) + pad256(D * 8) + pad256((size_t)n * m * 8), &a));
  double *dX = a.take<double>((size_t)n * D), *dY = a.take<double>((size_t)m * D), *dr = a.take<double>(D);
  double *dK = a.take<double>((size_t)n * m);
  RC(to_device_2d(h, X, ldx, dX, n, n, D));
  RC(to_device_2d(h, Y, ldy, dY, m, m, D));
  RC(to_device(h, rho, dr, D));
  RC(launch_gram_ard(h, n, m, D, dX, n, dY, m, alpha, dr, D, dK, n));
  RC(from_device_2d(h, dK, n, K, ldk, n, m));
  return finish(h);
}

extern "C" int gpb200_gram_se(gpb200_handle_t h, int n, const double *x, double alpha, double rho, double diag_add,
                              double *K, int ldk) {
  CHECK_H(h);
  if (n < 0) BAD_ARG(h, 2, "gram_se: negative n");
  if (ldk < std::max(1, n)) BAD_ARG(h, 8, "gram_se: ldk < n");
  if (n == 0) return 0;
  const int np = round_up(n, TILE);
  Arena a;
  RC(ws_reserve(h, pad256((size_t)np * np * 8) + pad256(n * 8) + 256, &a));
  double *dK = a.take<double>((size_t)np * np), *dx = a.take<double>(n), *dth = a.take<double>(3);
  RC(to_device(h, x, dx, n));
  // sigma^2 + jitter = diag_add: pass sigma = 0 and jitter = diag_add
  const double th[3] = {alpha, rho, 0.0};
  GPB_CUDA(h, cudaMemcpyAsync(dth, th, sizeof(th), cudaMemcpyHostToDevice, h->stream));
  GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  RC(launch_gram_se_batched(h, n, np, dx, 0, dth, diag_add, 0, dK, 0, 1));
  RC(from_device_2d(h, dK, np, K, ldk, n, n));
  return finish(h);
}

extern "C" int gpb200_gram_deriv(gpb200_handle_t h, int n, const double *t, double alpha, double rho, int nblocks,
                                 const double *noise, double jitter, int quirk, double *K, int ldk) {
  CHECK_H(h);
  if (n < 0) BAD_ARG(h, 2, "gram_deriv: negative n");
  if (nblocks < 1 || nblocks > 3) BAD_ARG(h, 6, "gram_deriv: nblocks must be 1..3");
  const int N = n * nblocks;
  if (ldk < std::max(1, N)) BAD_ARG(h, 11, "gram_deriv: ldk too small");
  if (n == 0) return 0;
  double nz[3] = {0, 0, 0};
  if (h->device_ptrs) {
    GPB_CUDA(h, cudaMemcpyAsync(nz, noise, nblocks * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    GPB_CUDA(h, cudaStreamSynchronize(h->stream));
    return launch_gram_deriv(h, n, nblocks, t, alpha, rho, nz, jitter, quirk, K, ldk);
  }
  for (int b = 0; b < nblocks; b++) nz[b] = noise[b];
  Arena a;
  RC(ws_reserve(h, pad256(n * 8) + pad256((size_t)N * N * 8), &a));
  double *dt = a.take<double>(n), *dK = a.take<double>((size_t)N * N);
  RC(to_device(h, t, dt, n));
  RC(launch_gram_deriv(h, n, nblocks, dt, alpha, rho, nz, jitter, quirk, dK, N));
  RC(from_device_2d(h, dK, N, K, ldk, N, N));
  return finish(h);
}

extern "C" int gpb200_approx_L_basis(gpb200_handle_t h, int n, int M, double scale, const double *x, double sigma,
                                     double l, double *out, int ldo) {
  CHECK_H(h);
  if (n < 0 || M < 1) BAD_ARG(h, 2, "approx_L_basis: bad sizes");
  if (ldo < std::max(1, n)) BAD_ARG(h, 9, "approx_L_basis: ldo < n");
  if (n == 0) return 0;
  if (h->device_ptrs) return launch_approx_basis(h, n, M, scale, x, sigma, l, out, ldo);
  Arena a;
  RC(ws_reserve(h, pad256(n * 8) + pad256((size_t)n * M * 8), &a));
  double *dx = a.take<double>(n), *dout = a.take<double>((size_t)n * M);
  RC(to_device(h, x, dx, n));
  RC(launch_approx_basis(h, n, M, scale, dx, sigma, l, dout, n));
  RC(from_device_2d(h, dout, n, out, ldo, n, M));
  return finish(h);
}

// =================================================================================================
// a6-a8 factorisation and solves on caller matrices
// =================================================================================================
extern "C" int gpb200_potrf(gpb200_handle_t h, int n, double *A, int lda) {
  CHECK_H(h);
  if (n < 0) BAD_ARG(h, 2, "potrf: negative n");
  if (lda < std::max(1, n)) BAD_ARG(h, 4, "potrf: lda < n");
  if (n == 0) return 0;
  const int np = round_up(n, TILE);
  Arena a;
  const size_t mat = pad256((size_t)np * np * 8);
  RC(ws_reserve(h, mat + (h->device_ptrs ? 0 : mat) + 512, &a));
  double *Lbuf = a.take<double>((size_t)np * np);
  int *info = a.take<int>(1);
  GPB_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int), h->stream));
  const double *src = A;
  long long lds = lda;
  if (!h->device_ptrs) {
    double *stage = a.take<double>((size_t)np * np);
    RC(to_device_2d(h, A, lda, stage, n, n, n));
    src = stage;
    lds = n;
  }
  RC(launch_pack(h, n, n, src, lds, np, np, Lbuf, 1, 0.0));
  RC(chol_batched(h, Lbuf, np, (long long)np * np, n, 1, info, nullptr));
  int hinfo = 0;
  RC(read_info(h, info, &hinfo));
  if (h->device_ptrs) {
    RC(launch_unpack(h, n, n, Lbuf, np, A, lda, 1, 0.0));
  } else {
    double *stage = const_cast<double *>(src);
    RC(launch_unpack(h, n, n, Lbuf, np, stage, n, 1, 0.0));
    RC(from_device_2d(h, stage, n, A, lda, n, n));
  }
  RC(finish(h));
  return hinfo;
}

namespace {
// shared body of trsm_lower / potrs / trmv / mvn lpdf: bring L (n x n lower) into a padded buffer
int stage_lower(Handle *h, Arena &a, int n, int np, const double *L, int ldl, double **Lbuf_out) {
  double *Lbuf = a.take<double>((size_t)np * np);
  if (!Lbuf) BAD_ARG(h, 1002, "workspace exhausted");
  if (h->device_ptrs) {
    RC(launch_pack(h, n, n, L, ldl, np, np, Lbuf, 2, 0.0));
  } else {
    double *stage = a.take<double>((size_t)n * n);
    if (!stage) BAD_ARG(h, 1002, "workspace exhausted");
    RC(to_device_2d(h, L, ldl, stage, n, n, n));
    RC(launch_pack(h, n, n, stage, n, np, np, Lbuf, 2, 0.0));
  }
  *Lbuf_out = Lbuf;
  return 0;
}

// X = W * B (W lower-triangular inverse in Lbuf, B np x rp padded) and optionally X = W^T * X
int tasks_mul(Handle *h, int kind, int nt, int rt, TaskList *out) {
  const long long key = tkey(kind, nt, rt);
  if (cached(h, key, out)) return 0;
  std::vector<TileTask> t;
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < rt; j++) {
      if (kind == TK_MUL_WB)  // X[i,j] = sum_{k<=i} W[i,k] B[k,j]   (NN)
        t.push_back({i * TILE, 0, 0, j * TILE, i * TILE, j * TILE, (i + 1) * TILE, 0});
      else                    // X[i,j] = sum_{k>=i} W[k,i] B[k,j]   (TN)
        t.push_back({i * TILE, i * TILE, i * TILE, j * TILE, i * TILE, j * TILE, (nt - i) * TILE, 0});
    }
  sort_desc(t, 0);
  std::vector<int> off = {0, (int)t.size()};
  return upload_tasks(h, key, t, off, out);
}

int solve_common(Handle *h, int n, int nrhs, const double *L, int ldl, double *B, int ldb, bool both) {
  const int np = round_up(n, TILE), rp = round_up(nrhs, TILE);
  const size_t mat = pad256((size_t)np * np * 8), rhs = pad256((size_t)np * rp * 8);
  Arena a;
  RC(ws_reserve(h, 3 * mat + 3 * rhs + 1024, &a));
  double *Lbuf = nullptr;
  RC(stage_lower(h, a, n, np, L, ldl, &Lbuf));
  double *Sbuf = a.take<double>((size_t)np * np);
  double *B0 = a.take<double>((size_t)np * rp), *B1 = a.take<double>((size_t)np * rp);
  const double *src = B;
  long long lds = ldb;
  if (!h->device_ptrs) {
    double *stage = a.take<double>((size_t)n * nrhs);
    RC(to_device_2d(h, B, ldb, stage, n, n, nrhs));
    src = stage;
    lds = n;
  }
  RC(launch_pack(h, n, nrhs, src, lds, np, rp, B0, 0, 0.0));
  RC(trtri_batched(h, Lbuf, Sbuf, np, (long long)np * np, 1));
  const int nt = np / TILE, rt = rp / TILE;
  TaskList t1;
  RC(tasks_mul(h, TK_MUL_WB, nt, rt, &t1));
  GemmParams p{};
  p.A = mref(Lbuf, np, 0);
  p.B = mref(B0, np, 0);
  p.C = mref(B1, np, 0);
  p.alpha = 1.0;
  p.tasks = t1.at(0);
  RC(launch_gemm(h, LAYOUT_NN, EPI_AXPBY, p, t1.count(0), 1));
  double *res = B1;
  if (both) {
    TaskList t2;
    RC(tasks_mul(h, TK_MUL_WTB, nt, rt, &t2));
    GemmParams q{};
    q.A = mref(Lbuf, np, 0);
    q.B = mref(B1, np, 0);
    q.C = mref(B0, np, 0);
    q.alpha = 1.0;
    q.tasks = t2.at(0);
    RC(launch_gemm(h, LAYOUT_TN, EPI_AXPBY, q, t2.count(0), 1));
    res = B0;
  }
  if (h->device_ptrs) {
    RC(launch_unpack(h, n, nrhs, res, np, B, ldb, 0, 0.0));
  } else {
    double *stage = const_cast<double *>(src);
    RC(launch_unpack(h, n, nrhs, res, np, stage, n, 0, 0.0));
    RC(from_device_2d(h, stage, n, B, ldb, n, nrhs));
  }
  return finish(h);
}
}  // namespace

extern "C" int gpb200_trsm_lower(gpb200_handle_t h, int n, int nrhs, const double *L, int ldl, double *B, int ldb) {
  CHECK_H(h);
  if (n < 0 || nrhs < 0) BAD_ARG(h, 2, "trsm_lower: negative size");
  if (ldl < std::max(1, n) || ldb < std::max(1, n)) BAD_ARG(h, 5, "trsm_lower: bad leading dimension");
  if (n == 0 || nrhs == 0) return 0;
  return solve_common(h, n, nrhs, L, ldl, B, ldb, false);
}

extern "C" int gpb200_potrs(gpb200_handle_t h, int n, int nrhs, const double *L, int ldl, double *B, int ldb) {
  CHECK_H(h);
  if (n < 0 || nrhs < 0) BAD_ARG(h, 2, "potrs: negative size");
  if (ldl < std::max(1, n) || ldb < std::max(1, n)) BAD_ARG(h, 5, "potrs: bad leading dimension");
  if (n == 0 || nrhs == 0) return 0;
  return solve_common(h, n, nrhs, L, ldl, B, ldb, true);
}

extern "C" int gpb200_trmv_lower(gpb200_handle_t h, int n, const double *L, int ldl, const double *z, double *f) {
  CHECK_H(h);
  if (n < 0) BAD_ARG(h, 2, "trmv_lower: negative n");
  if (ldl < std::max(1, n)) BAD_ARG(h, 4, "trmv_lower: ldl < n");
  if (n == 0) return 0;
  const int np = round_up(n, TILE);
  Arena a;
  RC(ws_reserve(h, 2 * pad256((size_t)np * np * 8) + 4 * pad256(np * 8), &a));
  double *Lbuf = nullptr;
  RC(stage_lower(h, a, n, np, L, ldl, &Lbuf));
  double *dz = a.take<double>(np), *df = a.take<double>(np);
  RC(to_device(h, z, dz, n));
  RC(launch_trmv_lower_n(h, np, Lbuf, 0, dz, 0, n, df, 0, 1));
  RC(from_device(h, df, f, n * sizeof(double)));
  return finish(h);
}

extern "C" int gpb200_trmv_lower_t(gpb200_handle_t h, int n, const double *L, int ldl, const double *z, double *f) {
  CHECK_H(h);
  if (n < 0) BAD_ARG(h, 2, "trmv_lower_t: negative n");
  if (ldl < std::max(1, n)) BAD_ARG(h, 4, "trmv_lower_t: ldl < n");
  if (n == 0) return 0;
  const int np = round_up(n, TILE);
  Arena a;
  RC(ws_reserve(h, 2 * pad256((size_t)np * np * 8) + 4 * pad256(np * 8), &a));
  double *Lbuf = nullptr;
  RC(stage_lower(h, a, n, np, L, ldl, &Lbuf));
  double *dz = a.take<double>(np), *df = a.take<double>(np);
  GPB_CUDA(h, cudaMemsetAsync(dz, 0, sizeof(double) * np, h->stream));
  RC(to_device(h, z, dz, n));
  RC(launch_trmv_lower_t(h, np, Lbuf, 0, dz, 0, df, 0, 1));
  RC(from_device(h, df, f, n * sizeof(double)));
  return finish(h);
}

extern "C" int gpb200_mvn_chol_lpdf(gpb200_handle_t h, int n, const double *y, const double *mu, const double *L,
                                    int ldl, int drop_constants, double *lp) {
  CHECK_H(h);
  if (n < 0) BAD_ARG(h, 2, "mvn_chol_lpdf: negative n");
  if (ldl < std::max(1, n)) BAD_ARG(h, 6, "mvn_chol_lpdf: ldl < n");
  const int np = round_up(std::max(n, 1), TILE);
  Arena a;
  RC(ws_reserve(h, 3 * pad256((size_t)np * np * 8) + 8 * pad256(np * 8), &a));
  double *Lbuf = nullptr;
  RC(stage_lower(h, a, n, np, L, ldl, &Lbuf));
  double *Wd = a.take<double>((size_t)np * np);
  double *dy = a.take<double>(np), *dmu = a.take<double>(np), *dz = a.take<double>(np), *out2 = a.take<double>(2);
  RC(to_device(h, y, dy, n));
  if (mu) RC(to_device(h, mu, dmu, n));
  RC(launch_tile_inverse(h, Lbuf, Wd, np, 0, np / TILE, 1));
  double *dacc = a.take<double>(np);
  RC(launch_trsv_sweep(h, np, Lbuf, Wd, 0, dy, 0, mu ? dmu : nullptr, n, dz, dacc, np, 1));
  RC(launch_sumsq_logdiag(h, n, dz, Lbuf, np, out2));
  double r[2];
  GPB_CUDA(h, cudaMemcpyAsync(r, out2, sizeof(r), cudaMemcpyDeviceToHost, h->stream));
  GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  double v = -r[1] - 0.5 * r[0];
  if (!drop_constants) v -= 0.5 * n * 1.8378770664093454835606594728112;
  if (h->device_ptrs) {
    GPB_CUDA(h, cudaMemcpyAsync(lp, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    GPB_CUDA(h, cudaStreamSynchronize(h->stream));
  } else {
    *lp = v;
  }
  return 0;
}

// =================================================================================================
// CS-A fused LML + gradient
// =================================================================================================
extern "C" int gpb200_lml_grad_batched(gpb200_handle_t h, int n, int B, const double *x, long long x_stride,
                                       const double *y, long long y_stride, const double *theta, double jitter,
                                       int want_grad, double *lml, double *grad, int *info) {
  CHECK_H(h);
  if (n < 1) BAD_ARG(h, 2, "lml_grad_batched: n must be >= 1");
  if (B < 0) BAD_ARG(h, 3, "lml_grad_batched: negative batch");
  if (x_stride != 0 && x_stride < n) BAD_ARG(h, 5, "lml_grad_batched: x_stride < n");
  if (y_stride != 0 && y_stride < n) BAD_ARG(h, 7, "lml_grad_batched: y_stride < n");
  if (B == 0) return 0;
  const int np = round_up(n, TILE), nt = np / TILE;
  const long long mat = (long long)np * np;
  TaskList tl;
  RC(tasks_lauum(h, nt, &tl));
  const int ntasks = tl.count(0);

  // chunk the batch so that the resident set fits the workspace limit.  cudaMemGetInfo costs
  // milliseconds with tens of GB allocated, so it is only consulted when the workspace must grow.
  const size_t per_item = pad256(mat * 8) * 2 + 3 * pad256(np * 8) + pad256((size_t)ntasks * 32) + 64;
  const size_t fixed = pad256((size_t)B * (x_stride ? n : 0) * 8 + n * 8) + pad256((size_t)B * (y_stride ? n : 0) * 8 + n * 8) +
                       pad256((size_t)B * 24) + pad256((size_t)B * 8) + pad256((size_t)B * 24) + pad256((size_t)B * 4) + 4096;
  int Bc = B;
  if (h->ws_limit > 0 || fixed + per_item * (size_t)B + 8192 > h->ws_bytes) {
    size_t limit = (size_t)h->ws_limit;
    if (h->ws_limit <= 0) {
      size_t freeb = 0, totalb = 0;
      GPB_CUDA(h, cudaMemGetInfo(&freeb, &totalb));
      limit = (size_t)((freeb + h->ws_bytes) * 0.85);
    }
    if (limit < fixed + per_item + 8192) BAD_ARG(h, 1002, "lml_grad_batched: workspace limit too small for one item");
    Bc = (int)std::min<size_t>((size_t)B, (limit - fixed - 8192) / per_item);
  }
  Bc = std::min(Bc, 65535);  // gridDim.y of the batched launches
  Arena a;
  RC(ws_reserve(h, fixed + per_item * (size_t)Bc + 8192, &a));

  const long long xs = x_stride ? n : 0, ys = y_stride ? n : 0;
  double *dx = a.take<double>(x_stride ? (size_t)B * n : n);
  double *dy = a.take<double>(y_stride ? (size_t)B * n : n);
  double *dth = a.take<double>((size_t)B * 3);
  double *dlml = a.take<double>(B), *dgrad = a.take<double>((size_t)B * 3);
  int *dinfo = a.take<int>(B);
  double *Lbuf = a.take<double>((size_t)Bc * mat), *Sbuf = a.take<double>((size_t)Bc * mat);
  double *zbuf = a.take<double>((size_t)Bc * np), *abuf = a.take<double>((size_t)Bc * np), *dvec = a.take<double>((size_t)Bc * np);
  double *partial = a.take<double>((size_t)Bc * ntasks * 4);
  if (!partial) BAD_ARG(h, 1002, "lml_grad_batched: workspace arithmetic error");

  // The kernel sequence (everything between staging the inputs and reading the outputs).
  auto run_sequence = [&]() -> int {
    GPB_CUDA(h, cudaMemsetAsync(dinfo, 0, (size_t)B * sizeof(int), h->stream));
    for (int b0 = 0; b0 < B; b0 += Bc) {
      const int bc = std::min(Bc, B - b0);
      const double *cx = dx + (long long)b0 * xs, *cy = dy + (long long)b0 * ys, *cth = dth + (long long)b0 * 3;
      RC(launch_gram_se_batched(h, n, np, cx, xs, cth, jitter, 1, Lbuf, mat, bc));
      RC(chol_batched(h, Lbuf, np, mat, n, bc, dinfo + b0, nullptr));
      RC(extract_diag(h, np, Lbuf, mat, dvec, bc));
      if (want_grad) {
        RC(trtri_batched(h, Lbuf, Sbuf, np, mat, bc));
        RC(launch_trmv_lower_n(h, np, Lbuf, mat, cy, ys, n, zbuf, np, bc));
        RC(launch_trmv_lower_t(h, np, Lbuf, mat, zbuf, np, abuf, np, bc));
        GemmParams p{};
        p.A = mref(Lbuf, np, mat);
        p.B = mref(Lbuf, np, mat);
        p.C = mref(nullptr, np, mat);
        p.tasks = tl.at(0);
        p.x = cx; p.x_stride = xs;
        p.avec = abuf; p.a_stride = np;
        p.theta = cth;
        p.partial = partial;
        p.n = n;
        p.ntasks = ntasks;
        RC(launch_gemm(h, LAYOUT_TN, EPI_TRACE, p, ntasks, bc));
      } else {
        RC(launch_tile_inverse(h, Lbuf, Sbuf, np, mat, nt, bc));
        if (bc >= 32) RC(launch_trsv_blocked(h, np, Lbuf, Sbuf, mat, cy, ys, nullptr, n, zbuf, np, bc));
        else RC(launch_trsv_sweep(h, np, Lbuf, Sbuf, mat, cy, ys, nullptr, n, zbuf, abuf, np, bc));
      }
      RC(launch_finalize(h, n, np, want_grad, dvec, zbuf, abuf, partial, ntasks, cth, dlml + b0, dgrad + (long long)b0 * 3, bc));
    }
    return 0;
  };

  // Small problems are launch-latency bound (tens of launches of a few microseconds each): replay
  // them as one CUDA graph on the handle's own stream, ordered against the caller's stream by events.
  const bool use_graph = h->graphs_enabled && !h->profiling && Bc == B && nt <= 16 && (long long)B * nt * nt <= 4096;
  cudaStream_t user_stream = h->stream;
  if (use_graph) {
    GPB_CUDA(h, cudaEventRecord(h->g_in, user_stream));
    GPB_CUDA(h, cudaStreamWaitEvent(h->gstream, h->g_in, 0));
    h->stream = h->gstream;
  }
  int rc = 0;
  do {
    if (x_stride) rc = to_device_2d(h, x, x_stride, dx, n, n, B); else rc = to_device(h, x, dx, n);
    if (rc) break;
    if (y_stride) rc = to_device_2d(h, y, y_stride, dy, n, n, B); else rc = to_device(h, y, dy, n);
    if (rc) break;
    if ((rc = to_device(h, theta, dth, (size_t)B * 3))) break;
    if (!use_graph) {
      rc = run_sequence();
    } else {
      long long jbits;
      memcpy(&jbits, &jitter, sizeof(jbits));
      const std::vector<long long> key = {n, B, want_grad, xs, ys, jbits, (long long)(uintptr_t)h->ws, h->chol_panel_override};
      auto it = h->graphs.find(key);
      if (it == h->graphs.end()) {
        // task lists allocate and synchronise on first use: make sure they exist before the capture
        TaskList t1, t2, t3;
        const int pt = h->chol_panel_override > 0 ? std::min(nt, h->chol_panel_override) : chol_panel_tiles(nt, B);
        if ((rc = tasks_chol(h, nt, pt, &t1, &t2))) break;
        if (nt > 1 && (rc = tasks_trtri(h, nt, &t1, &t3))) break;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(h->gstream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { rc = -1000; break; }
        const long long before = h->launches;
        rc = run_sequence();
        const cudaError_t ce = cudaStreamEndCapture(h->gstream, &graph);
        if (rc || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); if (!rc) rc = -1000; break; }
        cudaGraphExec_t exec = nullptr;
        if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) { cudaGraphDestroy(graph); rc = -1000; break; }
        cudaGraphDestroy(graph);
        it = h->graphs.emplace(key, Handle::GraphEntry{exec, h->launches - before}).first;
        h->launches = before;  // counted per replay below
      }
      if (cudaGraphLaunch(it->second.exec, h->gstream) != cudaSuccess) { rc = -1001; break; }
      h->launches += it->second.nodes;
      h->graph_replays++;
    }
    if (rc) break;
    if ((rc = from_device(h, dlml, lml, (size_t)B * sizeof(double)))) break;
    if (want_grad && grad && (rc = from_device(h, dgrad, grad, (size_t)B * 3 * sizeof(double)))) break;
    if (info && (rc = from_device(h, dinfo, info, (size_t)B * sizeof(int)))) break;
    rc = finish(h);
  } while (0);
  if (use_graph) {
    h->stream = user_stream;
    cudaEventRecord(h->g_out, h->gstream);
    cudaStreamWaitEvent(user_stream, h->g_out, 0);
  }
  if (rc && !h->err[0]) snprintf(h->err, sizeof(h->err), "lml_grad_batched: CUDA graph path failed (%d)", rc);
  return rc;
}

// number of CUDA-graph replays so far (small evaluations are replayed as one graph)
extern "C" long long gpb200_graph_replays(gpb200_handle_t h) { return h ? h->graph_replays : 0; }

extern "C" int gpb200_lml_grad(gpb200_handle_t h, int n, const double *x, const double *y, const double *theta,
                               double jitter, double *lml, double *grad) {
  CHECK_H(h);
  if (h->device_ptrs) {
    // info must live on the device in device-pointer mode; use a private slot
    int *dinfo = nullptr;
    GPB_CUDA(h, cudaMalloc(&dinfo, sizeof(int)));
    int rc = gpb200_lml_grad_batched(h, n, 1, x, 0, y, 0, theta, jitter, grad != nullptr, lml, grad, dinfo);
    int hinfo = 0;
    if (rc == 0) rc = read_info(h, dinfo, &hinfo);
    cudaFree(dinfo);
    return rc ? rc : hinfo;
  }
  int info = 0;
  int rc = gpb200_lml_grad_batched(h, n, 1, x, 0, y, 0, theta, jitter, grad != nullptr, lml, grad, &info);
  return rc ? rc : info;
}

// =================================================================================================
// a1-a3 rbf_cov_chol and its consumers
// =================================================================================================
namespace {
int tasks_tangent(Handle *h, int nt, TaskList *t1, TaskList *ta, TaskList *tl) {
  const long long k1 = tkey(TK_TAN_T1, nt), k2 = tkey(TK_TAN_A, nt), k3 = tkey(TK_TAN_LDOT, nt);
  if (cached(h, k1, t1) && cached(h, k2, ta) && cached(h, k3, tl)) return 0;
  std::vector<TileTask> v1, v2, v3;
  for (int i = 0; i < nt; i++)
    for (int j = 0; j <= i; j++) {
      // T1[i,j] = sum_{k<=i} W[i,k] Kdot[k,j]                       (NN)
      v1.push_back({i * TILE, 0, 0, j * TILE, i * TILE, j * TILE, (i + 1) * TILE, 0});
      // A[i,j]  = sum_{k<=j} T1[i,k] W[j,k]                          (NT)
      v2.push_back({i * TILE, 0, j * TILE, 0, i * TILE, j * TILE, (j + 1) * TILE, 0});
      // Ldot[i,j] = sum_{k=j..i} L[i,k] Phi[k,j]                     (NN)
      v3.push_back({i * TILE, j * TILE, j * TILE, j * TILE, i * TILE, j * TILE, (i - j + 1) * TILE, 0});
    }
  sort_desc(v1, 0); sort_desc(v2, 0); sort_desc(v3, 0);
  std::vector<int> o1 = {0, (int)v1.size()}, o2 = {0, (int)v2.size()}, o3 = {0, (int)v3.size()};
  RC(upload_tasks(h, k1, v1, o1, t1));
  RC(upload_tasks(h, k2, v2, o2, ta));
  return upload_tasks(h, k3, v3, o3, tl);
}
}  // namespace

namespace {
// L = chol(S), dL = L Phi(L^-1 Sdot L^-T) for the Gram/tangent pair selected by `mode`
int chol_tangent_common(Handle *h, int n, const double *x1, double alpha, const double *ls_host, int P, double dadd,
                        int mode, double *L, double *dLdl, int *info_out_host) {
  const int np = round_up(n, TILE), nt = np / TILE;
  const size_t mat = (size_t)np * np;
  const long long st = (long long)mat;
  Arena a;
  RC(ws_reserve(h, 4 * P * pad256(mat * 8) + pad256(n * 8) + pad256((size_t)n * n * 8) + pad256(P * 8) + pad256(P * 4) + 2048, &a));
  double *Lbuf = a.take<double>(mat * P), *Sbuf = a.take<double>(mat * P), *Dbuf = a.take<double>(mat * P),
         *Lkeep = a.take<double>(mat * P);
  double *dx = a.take<double>(n), *dls = a.take<double>(P);
  int *info = a.take<int>(P);
  double *stage = h->device_ptrs ? nullptr : a.take<double>((size_t)n * n);
  if (!info || (!h->device_ptrs && !stage)) BAD_ARG(h, 1002, "chol_tangent: workspace exhausted");
  GPB_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int) * P, h->stream));
  RC(to_device(h, x1, dx, n));
  GPB_CUDA(h, cudaMemcpyAsync(dls, ls_host, sizeof(double) * P, cudaMemcpyHostToDevice, h->stream));
  GPB_CUDA(h, cudaStreamSynchronize(h->stream));  // ls_host may be a stack temporary
  RC(launch_gram_tangent(h, n, np, dx, alpha, dls, dadd, mode, Lbuf, Dbuf, st, P));
  RC(chol_batched(h, Lbuf, np, st, n, P, info, nullptr));
  GPB_CUDA(h, cudaMemcpyAsync(info_out_host, info, sizeof(int) * P, cudaMemcpyDeviceToHost, h->stream));
  GPB_CUDA(h, cudaMemcpyAsync(Lkeep, Lbuf, mat * P * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  RC(trtri_batched(h, Lbuf, Sbuf, np, st, P));
  TaskList t1, ta, tl;
  RC(tasks_tangent(h, nt, &t1, &ta, &tl));
  {  // T1 = W Kdot  -> Sbuf (lower tiles)
    GemmParams p{};
    p.A = mref(Lbuf, np, st); p.B = mref(Dbuf, np, st); p.C = mref(Sbuf, np, st); p.alpha = 1.0; p.tasks = t1.at(0);
    RC(launch_gemm(h, LAYOUT_NN, EPI_AXPBY, p, t1.count(0), P));
  }
  {  // A = T1 W^T -> Dbuf (lower tiles)
    GemmParams p{};
    p.A = mref(Sbuf, np, st); p.B = mref(Lbuf, np, st); p.C = mref(Dbuf, np, st); p.alpha = 1.0; p.tasks = ta.at(0);
    RC(launch_gemm(h, LAYOUT_NT, EPI_AXPBY, p, ta.count(0), P));
  }
  RC(launch_phi_lower(h, np, Dbuf, st, P));
  {  // Ldot = L Phi(A) -> Sbuf (lower tiles)
    GemmParams p{};
    p.A = mref(Lkeep, np, st); p.B = mref(Dbuf, np, st); p.C = mref(Sbuf, np, st); p.alpha = 1.0; p.tasks = tl.at(0);
    RC(launch_gemm(h, LAYOUT_NN, EPI_AXPBY, p, tl.count(0), P));
  }
  const size_t nn = (size_t)n * n;
  for (int q = 0; q < P; q++) {
    if (h->device_ptrs) {
      RC(launch_unpack(h, n, n, Lkeep + q * mat, np, L + q * nn, n, 1, 0.0));
      RC(launch_unpack(h, n, n, Sbuf + q * mat, np, dLdl + q * nn, n, 1, 0.0));
    } else {
      RC(launch_unpack(h, n, n, Lkeep + q * mat, np, stage, n, 1, 0.0));
      RC(from_device(h, stage, L + q * nn, nn * sizeof(double)));
      GPB_CUDA(h, cudaStreamSynchronize(h->stream));
      RC(launch_unpack(h, n, n, Sbuf + q * mat, np, stage, n, 1, 0.0));
      RC(from_device(h, stage, dLdl + q * nn, nn * sizeof(double)));
      GPB_CUDA(h, cudaStreamSynchronize(h->stream));
    }
  }
  GPB_CUDA(h, cudaStreamSynchronize(h->stream));  // info_out_host is valid from here
  return 0;
}
}  // namespace

extern "C" int gpb200_rbf_cov_chol(gpb200_handle_t h, int n, const double *x1, double l, double *L, double *dLdl) {
  CHECK_H(h);
  if (n < 0) BAD_ARG(h, 2, "rbf_cov_chol: negative n");
  if (n == 0) return 0;
  int info = 0;
  RC(chol_tangent_common(h, n, x1, 1.0, &l, 1, 1e-10, 0, L, dLdl, &info));
  return info;
}

// P length-scales in one call: the tables Ls[P], dLdls[P] that approx_L / approx_Lz interpolate
// (models/interpolated_gp.stan:15-21 builds them with P separate Choleskys; test_interpolate.R:9 uses
// P = 10).  ls and info are HOST arrays of length P; L, dLdl hold P consecutive n x n matrices.
extern "C" int gpb200_rbf_cov_chol_batched(gpb200_handle_t h, int n, const double *x1, int P, const double *ls,
                                           double *L, double *dLdl, int *info) {
  CHECK_H(h);
  if (n < 0) BAD_ARG(h, 2, "rbf_cov_chol_batched: negative n");
  if (P < 0) BAD_ARG(h, 4, "rbf_cov_chol_batched: negative P");
  if (n == 0 || P == 0) return 0;
  return chol_tangent_common(h, n, x1, 1.0, ls, P, 1e-10, 0, L, dLdl, info);
}

extern "C" int gpb200_se_chol_tangent(gpb200_handle_t h, int n, const double *x, double alpha, double rho,
                                      double diag_add, int wrt, double *L, double *dL) {
  CHECK_H(h);
  if (n < 0) BAD_ARG(h, 2, "se_chol_tangent: negative n");
  if (wrt != 0 && wrt != 1) BAD_ARG(h, 7, "se_chol_tangent: wrt must be 0 (alpha) or 1 (rho)");
  if (n == 0) return 0;
  int info = 0;
  RC(chol_tangent_common(h, n, x, alpha, &rho, 1, diag_add, wrt == 1 ? 1 : 2, L, dL, &info));
  return info;
}

namespace {
int bracket(double l, int P, const double *lp) {  // covariance.cpp:56-61
  int lidx = 0;
  for (; lidx < P - 1; lidx++)
    if (lp[lidx + 1] >= l) break;
  if (lidx > P - 2) lidx = P - 2;
  return lidx;
}
}  // namespace

extern "C" int gpb200_approx_Lz(gpb200_handle_t h, int n, double l, int P, const double *lp, const double *const *Ls,
                                const double *const *dLdls, const double *z, double *vz, double *dvdl_z) {
  CHECK_H(h);
  if (n < 1) BAD_ARG(h, 2, "approx_L: n must be >= 1");
  if (P < 2) BAD_ARG(h, 4, "approx_L: need at least two grid points");
  // lp and the pointer tables are always host arrays; the tables they point to follow the pointer mode
  const int lidx = bracket(l, P, lp);
  const size_t nn = (size_t)n * n;
  const int np = round_up(n, TILE);
  Arena a;
  RC(ws_reserve(h, 6 * pad256(nn * 8) + 2 * pad256((size_t)np * np * 8) + 4 * pad256(np * 8), &a));
  const double *t[4] = {Ls[lidx], Ls[lidx + 1], dLdls[lidx], dLdls[lidx + 1]};
  const double *d[4];
  for (int q = 0; q < 4; q++) {
    if (h->device_ptrs) d[q] = t[q];
    else {
      double *s = a.take<double>(nn);
      RC(to_device(h, t[q], s, nn));
      d[q] = s;
    }
  }
  double *v = a.take<double>(nn), *dv = a.take<double>(nn);
  RC(launch_hermite(h, (long long)nn, n, d[0], d[1], d[2], d[3], lp[lidx], lp[lidx + 1], l, v, z ? dv : nullptr));
  if (!z) {  // approx_L proper: return the interpolated factor in vz
    RC(from_device(h, v, vz, nn * sizeof(double)));
    return finish(h);
  }
  double *Vp = a.take<double>((size_t)np * np), *Dp = a.take<double>((size_t)np * np);
  double *dz = a.take<double>(np), *o1 = a.take<double>(np), *o2 = a.take<double>(np);
  RC(launch_pack(h, n, n, v, n, np, np, Vp, 2, 0.0));
  RC(launch_pack(h, n, n, dv, n, np, np, Dp, 2, 0.0));
  RC(to_device(h, z, dz, n));
  RC(launch_trmv_lower_n(h, np, Vp, 0, dz, 0, n, o1, 0, 1));
  // the padded diagonal of Dp is 1 (identity padding) but z is zero there, so it contributes nothing
  RC(launch_trmv_lower_n(h, np, Dp, 0, dz, 0, n, o2, 0, 1));
  RC(from_device(h, o1, vz, n * sizeof(double)));
  if (dvdl_z) RC(from_device(h, o2, dvdl_z, n * sizeof(double)));
  return finish(h);
}

extern "C" int gpb200_approx_L(gpb200_handle_t h, int n, double l, int P, const double *lp, const double *const *Ls,
                               const double *const *dLdls, double *out) {
  return gpb200_approx_Lz(h, n, l, P, lp, Ls, dLdls, nullptr, out, nullptr);
}

// =================================================================================================
// a10 conditioning
// =================================================================================================
namespace {
int tasks_cond(Handle *h, int nt, int mt, TaskList *tv, TaskList *tc) {
  const long long k1 = tkey(TK_COND_V, nt, mt), k2 = tkey(TK_COND_COV, nt, mt);
  if (cached(h, k1, tv) && cached(h, k2, tc)) return 0;
  std::vector<TileTask> v1, v2;
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < mt; j++)  // V[i,j] = sum_{k<=i} W[i,k] Ks[j,k]   (NT), V is np x mp
      v1.push_back({i * TILE, 0, j * TILE, 0, i * TILE, j * TILE, (i + 1) * TILE, 0});
  for (int i = 0; i < mt; i++)
    for (int j = 0; j <= i; j++)  // C[i,j] = Kss[i,j] - sum_k V[k,i] V[k,j]   (TN), lower tiles
      v2.push_back({0, i * TILE, 0, j * TILE, i * TILE, j * TILE, nt * TILE, 0});
  sort_desc(v1, 0);
  std::vector<int> o1 = {0, (int)v1.size()}, o2 = {0, (int)v2.size()};
  RC(upload_tasks(h, k1, v1, o1, tv));
  return upload_tasks(h, k2, v2, o2, tc);
}

// all pointers here are DEVICE pointers; K (n x n, ldk), Ks (m x n), Kss (m x m); rhs length n;
// mean_add (length m) may be null.
int condition_device(Handle *h, Arena &a, int n, int m, const double *K, long long ldk, const double *Ks,
                     long long ldks, const double *Kss, long long ldkss, const double *rhs, const double *mean_add,
                     double noise_var, double jitter, double *mu, double *cov, long long ldcov, int *hinfo) {
  const int np = round_up(n, TILE), mp = round_up(m, TILE);
  const size_t mat = (size_t)np * np;
  double *Lbuf = a.take<double>(mat), *Sbuf = a.take<double>(mat);
  double *Ksp = a.take<double>((size_t)mp * np), *V = a.take<double>((size_t)np * mp), *Cp = a.take<double>((size_t)mp * mp);
  double *z = a.take<double>(np);
  int *info = a.take<int>(1);
  if (!info) BAD_ARG(h, 1002, "gp_condition: workspace exhausted");
  GPB_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int), h->stream));
  RC(launch_pack(h, n, n, K, ldk, np, np, Lbuf, 1, noise_var));
  RC(launch_pack(h, m, n, Ks, ldks, mp, np, Ksp, 0, 0.0));
  RC(launch_pack(h, m, m, Kss, ldkss, mp, mp, Cp, 0, 0.0));
  RC(chol_batched(h, Lbuf, np, (long long)mat, n, 1, info, nullptr));
  RC(read_info(h, info, hinfo));
  RC(trtri_batched(h, Lbuf, Sbuf, np, (long long)mat, 1));
  RC(launch_trmv_lower_n(h, np, Lbuf, 0, rhs, 0, n, z, 0, 1));
  TaskList tv, tc;
  RC(tasks_cond(h, np / TILE, mp / TILE, &tv, &tc));
  {
    GemmParams p{};
    p.A = mref(Lbuf, np, 0); p.B = mref(Ksp, mp, 0); p.C = mref(V, np, 0); p.alpha = 1.0; p.tasks = tv.at(0);
    RC(launch_gemm(h, LAYOUT_NT, EPI_AXPBY, p, tv.count(0), 1));
  }
  {
    GemmParams p{};
    p.A = mref(V, np, 0); p.B = mref(V, np, 0); p.C = mref(Cp, mp, 0); p.C0 = mref(Cp, mp, 0);
    p.alpha = -1.0; p.beta = 1.0; p.tasks = tc.at(0);
    RC(launch_gemm(h, LAYOUT_TN, EPI_AXPBY, p, tc.count(0), 1));
  }
  RC(launch_gemv_t(h, np, m, V, np, z, mean_add, mu));
  RC(launch_unpack(h, m, m, Cp, mp, cov, ldcov, 2, jitter));
  return 0;
}

size_t cond_ws(int n, int m) {
  const size_t np = round_up(n, TILE), mp = round_up(m, TILE);
  return 2 * pad256(np * np * 8) + 2 * pad256(np * mp * 8) + pad256(mp * mp * 8) + pad256(np * 8) + 1024;
}
}  // namespace

extern "C" int gpb200_gp_condition(gpb200_handle_t h, int n, int m, const double *K, int ldk, const double *Ks,
                                   int ldks, const double *Kss, int ldkss, const double *y, double noise_var,
                                   double jitter, double *mu, double *cov, int ldcov) {
  CHECK_H(h);
  if (n < 1 || m < 1) BAD_ARG(h, 2, "gp_condition: sizes must be >= 1");
  if (ldk < n || ldks < m || ldkss < m || ldcov < m) BAD_ARG(h, 5, "gp_condition: bad leading dimension");
  Arena a;
  const size_t stage = h->device_ptrs ? 0 : pad256((size_t)n * n * 8) + pad256((size_t)m * n * 8) + 2 * pad256((size_t)m * m * 8) + pad256(n * 8) + pad256(m * 8);
  RC(ws_reserve(h, cond_ws(n, m) + stage + 1024, &a));
  int hinfo = 0;
  if (h->device_ptrs) {
    RC(condition_device(h, a, n, m, K, ldk, Ks, ldks, Kss, ldkss, y, nullptr, noise_var, jitter, mu, cov, ldcov, &hinfo));
    return hinfo;
  }
  double *dK = a.take<double>((size_t)n * n), *dKs = a.take<double>((size_t)m * n), *dKss = a.take<double>((size_t)m * m);
  double *dcov = a.take<double>((size_t)m * m), *dy = a.take<double>(n), *dmu = a.take<double>(m);
  RC(to_device_2d(h, K, ldk, dK, n, n, n));
  RC(to_device_2d(h, Ks, ldks, dKs, m, m, n));
  RC(to_device_2d(h, Kss, ldkss, dKss, m, m, m));
  RC(to_device(h, y, dy, n));
  RC(condition_device(h, a, n, m, dK, n, dKs, m, dKss, m, dy, nullptr, noise_var, jitter, dmu, dcov, m, &hinfo));
  RC(from_device(h, dmu, mu, m * sizeof(double)));
  RC(from_device_2d(h, dcov, m, cov, ldcov, m, m));
  RC(finish(h));
  return hinfo;
}

namespace {
__global__ void sub_kernel(int n, const double *a, const double *b, double *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] - (b ? b[i] : 0.0);
}
}  // namespace

extern "C" int gpb200_cond_mvn(gpb200_handle_t h, int ng, int nd, const double *mean, const double *sigma, int lds,
                               const double *x_given, double *cond_mean, double *cond_var, int ldv) {
  CHECK_H(h);
  if (ng < 1 || nd < 1) BAD_ARG(h, 2, "cond_mvn: sizes must be >= 1");
  const int N = ng + nd;
  if (lds < N || ldv < nd) BAD_ARG(h, 6, "cond_mvn: bad leading dimension");
  Arena a;
  const size_t stage = h->device_ptrs ? pad256(ng * 8) : pad256((size_t)N * N * 8) + 3 * pad256(N * 8) + pad256((size_t)nd * nd * 8) + pad256(nd * 8);
  RC(ws_reserve(h, cond_ws(ng, nd) + stage + 1024, &a));
  const double *dS = sigma, *dmean = mean, *dxg = x_given;
  long long ld = lds;
  double *dcm = cond_mean, *dcv = cond_var;
  long long ldo = ldv;
  if (!h->device_ptrs) {
    double *s = a.take<double>((size_t)N * N), *mm = a.take<double>(N), *xg = a.take<double>(ng);
    RC(to_device_2d(h, sigma, lds, s, N, N, N));
    if (mean) RC(to_device(h, mean, mm, N));
    RC(to_device(h, x_given, xg, ng));
    dS = s; dmean = mean ? mm : nullptr; dxg = xg; ld = N;
    dcm = a.take<double>(nd); dcv = a.take<double>((size_t)nd * nd); ldo = nd;
  }
  double *rhs = a.take<double>(ng);
  ProfScope ps__(h, PC_OTHER);
  sub_kernel<<<(ng + 255) / 256, 256, 0, h->stream>>>(ng, dxg, dmean, rhs);
  GPB_LAUNCH_CHECK(h);
  int hinfo = 0;
  // given block first: D = S[0:ng,0:ng], C = S[ng:,0:ng] (nd x ng), B = S[ng:,ng:]
  RC(condition_device(h, a, ng, nd, dS, ld, dS + ng, ld, dS + ng + (long long)ng * ld, ld, rhs,
                      dmean ? dmean + ng : nullptr, 0.0, 0.0, dcm, dcv, ldo, &hinfo));
  if (!h->device_ptrs) {
    RC(from_device(h, dcm, cond_mean, nd * sizeof(double)));
    RC(from_device_2d(h, dcv, nd, cond_var, ldv, nd, nd));
  }
  RC(finish(h));
  return hinfo;
}

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
