"""Measures the FP64 GEMM peak of the GPU (cuBLAS DGEMM through torch.matmul, 8192^3), burst
(best of 10) and sustained (back-to-back for ~3 s). MEASURED_PEAKS.json has no FP64 entry
(SURVEY fact 9), so this is the denominator for the DMMA roofline. Writes JSON to stdout."""
import json
import time

import torch


def main(n=8192):
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    flops = 2.0 * n ** 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    reps = 0
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(5):
            torch.matmul(a, b, out=c)
        reps += 5
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = flops * reps / (e0.elapsed_time(e1) * 1e-3)
    print(json.dumps({"fp64_tflops": flops / best / 1e12, "fp64_tflops_sustained": sustained / 1e12, "n": n,
                      "gpu": torch.cuda.get_device_name(0),
                      "how": "torch.matmul fp64 8192^3 (cuBLAS DGEMM): best of 10 (burst) and back to back for 3 s (sustained)"}))


if __name__ == "__main__":
    main()
