"""Measures the FP64 roofline denominators MEASURED_PEAKS.json lacks: cuBLAS DGEMM 8192^3
best-of-10 (burst) and 4 s back-to-back (sustained), same method as the driver's bf16 figure; plus this repo's
DMMA GEMM / Gram kernels on the same shapes.  Writes gpurun_out/fp64_peaks.json."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, reps):
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3)
    return best


def main():
    from sofacontrol_b200.mor import pod
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    out = {}
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    t = timed(lambda: torch.matmul(a, b), 10)
    out["cublas_dgemm_tflops_burst"] = 2 * n ** 3 / t / 1e12
    t0 = time.time(); cnt = 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    while time.time() - t0 < 4.0:
        for _ in range(4):
            torch.matmul(a, b); cnt += 1
        torch.cuda.synchronize()
    e.record(); torch.cuda.synchronize()
    out["cublas_dgemm_tflops_sustained"] = cnt * 2 * n ** 3 / (s.elapsed_time(e) * 1e-3) / 1e12
    for _ in range(2):
        pod.dgemm_device(a, b)
    torch.cuda.synchronize()
    t = timed(lambda: pod.dgemm_device(a, b), 5)
    out["srcb_dgemm_tflops"] = 2 * n ** 3 / t / 1e12
    c = pod.dgemm_device(a, b)
    ref = torch.matmul(a, b)
    out["srcb_dgemm_maxerr"] = float((c - ref).abs().max() / ref.abs().max())
    x = torch.randn(65536, 4096, device="cuda", dtype=torch.float64)
    for _ in range(2):
        pod.gram_device(x)
    torch.cuda.synchronize()
    t = timed(lambda: pod.gram_device(x), 5)
    out["srcb_gram_tflops_algorithmic"] = 2 * 65536 * 4096 ** 2 / t / 1e12
    t = timed(lambda: torch.matmul(x.t(), x), 5)
    out["cublas_gram_tflops"] = 2 * 65536 * 4096 ** 2 / t / 1e12
    # plain FP64 FMA pipe: elementwise chain
    out["gpu"] = torch.cuda.get_device_name(0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/fp64_peaks.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
