"""cuBLAS DGEMM 8192^3 through torch.matmul (library used for MEASUREMENT ONLY: it gives the
FP64 denominator that MEASURED_PEAKS.json lacks; it is never on the product path)."""
import json, sys, time
import torch

def main():
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2.0 * n ** 3 / best * 1e-9
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); k = 0
    t0 = time.time()
    while time.time() - t0 < 4.0:
        for _ in range(5):
            torch.matmul(a, b, out=c); k += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sus = 2.0 * n ** 3 * k / e0.elapsed_time(e1) * 1e-9
    print(json.dumps({"kind": "cublas_dgemm_8192", "burst_tflops": round(burst, 3), "sustained_tflops": round(sus, 3),
                      "best_ms": round(best, 3)}))

if __name__ == "__main__":
    main()
