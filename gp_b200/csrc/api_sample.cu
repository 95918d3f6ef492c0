// C-ABI entry points for posterior sampling (f-4): device normal numbers and mvrnorm.
#include "host.cuh"

using namespace gpb;

extern "C" int gpb200_normal_fill(gpb200_handle_t h, unsigned long long seed, unsigned long long offset, long long len,
                                  double *out) {
  CHECK_H(h);
  if (len < 0) BAD_ARG(h, 4, "normal_fill: negative length");
  if (offset & 1ULL) BAD_ARG(h, 3, "normal_fill: offset must be even (normals are generated in Box-Muller pairs)");
  if (len == 0) return 0;
  if (h->device_ptrs) {
    RC(launch_normal_fill(h, seed, offset, len, len, len, out));
    return 0;
  }
  Arena a;
  RC(ws_reserve(h, pad256((size_t)len * 8) + 256, &a));
  double *d = a.take<double>((size_t)len);
  RC(launch_normal_fill(h, seed, offset, len, len, len, d));
  RC(from_device(h, d, out, (size_t)len * sizeof(double)));
  return finish(h);
}

namespace {
int tasks_lz(Handle *h, int nt, int dt, TaskList *out) {
  const long long key = tkey(TK_MUL_LZ, nt, dt);
  if (cached(h, key, out)) return 0;
  std::vector<TileTask> v;
  for (int i = 0; i < nt; i++)
    for (int j = 0; j < dt; j++)  // X[i,j] = sum_{k<=i} L[i,k] Z[k,j]   (NN)
      v.push_back({i * TILE, 0, 0, j * TILE, i * TILE, j * TILE, (i + 1) * TILE, TF_A_TRI_LAST});
  sort_desc(v, 0);
  std::vector<int> off = {0, (int)v.size()};
  return upload_tasks(h, key, v, off, out);
}
}  // namespace

extern "C" int gpb200_mvrnorm(gpb200_handle_t h, int ndraws, int m, const double *mu, const double *Sigma, int lds,
                              double jitter, unsigned long long seed, double *out, int ldo) {
  CHECK_H(h);
  if (ndraws < 0) BAD_ARG(h, 2, "mvrnorm: negative number of draws");
  if (m < 1) BAD_ARG(h, 3, "mvrnorm: dimension must be >= 1");
  if (m > MAX_DENSE_N) BAD_ARG(h, 3, "mvrnorm: dimensions above 65407 are not supported");
  if (lds < m) BAD_ARG(h, 6, "mvrnorm: lds < m");
  if (ldo < std::max(1, ndraws)) BAD_ARG(h, 10, "mvrnorm: ldo < ndraws");
  if (ndraws == 0) return 0;
  const int mp = round_up(m, TILE), dp = round_up(ndraws, TILE);
  const size_t mat = (size_t)mp * mp, zsz = (size_t)mp * dp;
  Arena a;
  const size_t stage = h->device_ptrs ? 0 : pad256((size_t)m * m * 8) + pad256((size_t)m * 8) + pad256((size_t)ndraws * m * 8);
  RC(ws_reserve(h, pad256(mat * 8) + 2 * pad256(zsz * 8) + stage + 1024, &a));
  double *Lbuf = a.take<double>(mat), *Z = a.take<double>(zsz), *X = a.take<double>(zsz);
  int *info = a.take<int>(1);
  const double *dS = Sigma, *dmu = mu;
  long long ld = lds;
  double *dout = out;
  long long ldout = ldo;
  if (!h->device_ptrs) {
    double *s = a.take<double>((size_t)m * m), *mm = a.take<double>(m);
    dout = a.take<double>((size_t)ndraws * m);
    if (!dout) BAD_ARG(h, 1002, "mvrnorm: workspace exhausted");
    RC(to_device_2d(h, Sigma, lds, s, m, m, m));
    if (mu) RC(to_device(h, mu, mm, m));
    dS = s; ld = m; dmu = mu ? mm : nullptr; ldout = ndraws;
  }
  GPB_CUDA(h, cudaMemsetAsync(info, 0, sizeof(int), h->stream));
  RC(launch_pack(h, m, m, dS, ld, mp, mp, Lbuf, 1, jitter));
  RC(chol_batched(h, Lbuf, mp, (long long)mat, m, 1, info));
  int hinfo = 0;
  RC(read_info(h, info, &hinfo));
  if (hinfo) return hinfo;
  // Z: m x ndraws standard normals, element (i, d) = normal number i + d*m of the stream; zero padding
  GPB_CUDA(h, cudaMemsetAsync(Z, 0, zsz * sizeof(double), h->stream));
  RC(launch_normal_fill(h, seed, 0, (long long)m * ndraws, m, mp, Z));
  TaskList tl;
  RC(tasks_lz(h, mp / TILE, dp / TILE, &tl));
  GemmParams p{};
  p.A = mref(Lbuf, mp, 0); p.B = mref(Z, mp, 0); p.C = mref(X, mp, 0); p.alpha = 1.0; p.tasks = tl.at(0);
  RC(launch_gemm(h, LAYOUT_NN, EPI_AXPBY, p, tl.count(0), 1));
  RC(launch_add_mean_transpose(h, m, ndraws, X, mp, dmu, dout, ldout));
  if (!h->device_ptrs) RC(from_device_2d(h, dout, ndraws, out, ldo, ndraws, m));
  RC(finish(h));
  return 0;
}
