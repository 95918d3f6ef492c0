"""Small end-to-end pass over every C-ABI entry point for compute-sanitizer (memcheck):
  compute-sanitizer --tool memcheck python tests/support/sanitize_smoke.py
Sizes are tiny on purpose (sanitizer slows kernels ~100x)."""
import sys

import numpy as np

sys.path.insert(0, ".")
from gp_b200 import capi  # noqa: E402
from oracle import gp_oracle as o  # noqa: E402

h = capi.Handle(0)
n = 200
x, y = o.synth_xy(n, 3)
th = o.synth_theta(3, 7)
lml, grad, info = h.lml_grad_batched(x, y, th)
ref = o.lml_grad(x, y, *th[0])
assert abs(lml[0] - ref[0]) < 1e-9 * abs(ref[0])
h.lml_grad_batched(x, y, th, want_grad=False)
h.set_chol_panel_tiles(1)
K = o.gram_se(x, 1.0, 1.0, 0.09)
L = h.potrf(K)
h.set_chol_panel_tiles(0)
h.trsm_lower(L, y); h.potrs(L, y); h.trmv_lower(L, y); h.trmv_lower_t(L, y); h.mvn_chol_lpdf(y, None, L)
h.rbf_cov_chol(np.arange(150) * 1.0, 0.7)
h.se_chol_tangent(x, 1.2, 0.9, 1e-6, 0)
t = np.linspace(0, 5, 70)
Kd = h.gram_deriv(t, 1.0, 1.0, [0.1, 0.0], 1e-6, nblocks=2)
h.cond_mvn(np.zeros(140), Kd, 70, np.sin(t))
h.gram_outer("TT", t, t, 1.0); h.kernel_eval("RT", t, t[::-1], 1.0); h.gram_ard(np.stack([t, t], 1), np.stack([t, t], 1), 1.0, [1.0, 2.0])
h.gram_se(t, 1.0, 1.0, 0.1)
h.gp_condition(h.gram_outer("QQ", t, t, 1.0), h.gram_outer("RQ", t, t, 1.0), h.gram_outer("RR", t, t, 1.0), np.sin(t), 0.01, 1e-8)
tabs = [o.rbf_cov_chol(np.arange(40) * 0.9, l) for l in (0.5, 0.8)]
h.approx_Lz(0.6, [0.5, 0.8], [a[0] for a in tabs], [a[1] for a in tabs], np.ones(40))
# round 2: the one-CTA kernel (n <= 128), the single-matrix latency schedule (look-ahead streams, quarter tiles, left-looking
# panel kernels, split mat-vec), a batch large enough for the diagonal-split launch, the reverse-mode adjoint
xs, ys = o.synth_xy(100, 5)
l2, g2, i2 = h.lml_grad_batched(xs, ys, o.synth_theta(4, 9))
r2 = o.lml_grad(xs, ys, *o.synth_theta(4, 9)[3])
assert abs(l2[3] - r2[0]) < 1e-9 * abs(r2[0])
h.lml_grad_batched(xs[:37], ys[:37], o.synth_theta(2, 9), want_grad=False)
x5, y5 = o.synth_xy(520, 6)
l5, g5, _ = h.lml_grad_batched(x5, y5, o.synth_theta(1, 11))
r5 = o.lml_grad(x5, y5, *o.synth_theta(1, 11)[0])
assert abs(l5[0] - r5[0]) < 1e-9 * abs(r5[0]) and np.max(np.abs(g5[0] - r5[1])) < 1e-9 * np.max(np.abs(r5[1]))
x3, y3 = o.synth_xy(300, 8)
h.lml_grad_batched(x3, y3, o.synth_theta(24, 12))
z = np.sin(np.arange(200) * 0.37)
f = h.latent_forward(x, 1.1, 0.9, 1e-6, z)
h.latent_backward(x, 1.1, 0.9, 1e-6, z, (y - f) / 0.09)
print("sanitize smoke ok, launches", h.launch_count())
h.close()
