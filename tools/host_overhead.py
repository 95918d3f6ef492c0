import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from gp_b200 import capi
dev = torch.device("cuda", 0)
h = capi.Handle(0)
stream = torch.cuda.current_stream(dev)
h.set_stream(stream.cuda_stream); h.set_pointer_mode(True)
n, B = 100, 1
rng = np.random.default_rng(0)
x = np.sort(rng.uniform(0, 5, n)); y = np.sin(x)
dx = torch.from_numpy(x).to(dev); dy = torch.from_numpy(y).to(dev)
dth = torch.tensor([[1.0, 1.0, 0.3]], dtype=torch.float64, device=dev)
lml = torch.empty(B, dtype=torch.float64, device=dev); grad = torch.empty(B, 3, dtype=torch.float64, device=dev)
info = torch.zeros(B, dtype=torch.int32, device=dev)
def loop(tag, reps=20):
    for _ in range(3): h.lml_grad_batched_device(n, B, dx, 0, dy, 0, dth, 0.0, True, lml, grad, info)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): h.lml_grad_batched_device(n, B, dx, 0, dy, 0, dth, 0.0, True, lml, grad, info)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(tag, "host enqueue ms/call %.3f" % ((t1 - t0) / reps * 1e3), "incl sync %.3f" % ((t2 - t0) / reps * 1e3), flush=True)
loop("fresh handle")
h.set_profiling(True); loop("profiling on"); h.get_profile(); h.set_profiling(False)
loop("profiling off again")
# big workspace then small again
nb, Bb = 2048, 64
xb = np.sort(rng.uniform(0, 100, nb)); dxb = torch.from_numpy(xb).to(dev); dyb = torch.from_numpy(np.sin(xb)).to(dev)
thb = torch.tensor([[1.0, 1.0, 0.3]] * Bb, dtype=torch.float64, device=dev)
l2 = torch.empty(Bb, dtype=torch.float64, device=dev); g2 = torch.empty(Bb, 3, dtype=torch.float64, device=dev); i2 = torch.zeros(Bb, dtype=torch.int32, device=dev)
h.lml_grad_batched_device(nb, Bb, dxb, 0, dyb, 0, thb, 0.0, True, l2, g2, i2); torch.cuda.synchronize()
loop("after a 4.3 GB workspace call")
