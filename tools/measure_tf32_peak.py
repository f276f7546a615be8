#!/usr/bin/env python
"""Measures the TF32 dense GEMM peak of this GPU the way MEASURED_PEAKS.json measures the bf16 one (driver recipe):
torch.matmul fp32 8192^3 with TF32 allowed (cuBLAS), best of 10 (burst) and back to back for 4 s (sustained), CUDA
events. Also re-measures bf16 the same way in the same process for the ratio. Writes one JSON line (profiles/)."""
import json
import sys
import time

import torch


def peak(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    fl = 2.0 * n ** 3
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) * 1e-12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0, k = time.time(), 0
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            a @ b
        k += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return best, k * fl / (e0.elapsed_time(e1) * 1e-3) * 1e-12


if __name__ == "__main__":
    tb, ts = peak(torch.float32, True)
    bb, bs = peak(torch.bfloat16, False)
    out = {"tf32_tflops": tb, "tf32_tflops_sustained": ts, "bf16_tflops": bb, "bf16_tflops_sustained": bs,
           "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "how": "torch.matmul 8192^3 (2*N^3): fp32 with allow_tf32 (cuBLAS TF32) and bf16; best of 10 (burst) and back "
                  "to back for 4 s (sustained), CUDA events"}
    print(json.dumps(out))
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
