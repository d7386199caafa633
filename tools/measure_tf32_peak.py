#!/usr/bin/env python
"""Measures the dense TF32 tensor-core peak of this GPU the way the driver measured the bf16 figure in MEASURED_PEAKS.json:
torch.matmul on 8192^3 fp32 operands with TF32 allowed (cuBLAS), best of 10 (burst) and back to back for 4 s (sustained), CUDA events.
BASELINE.md section 2 asks for this denominator before a tensor-pipe fraction of the 3xTF32 GEMMs is quoted.  Also re-measures bf16
with the same code so the two figures are comparable.  Prints one JSON object (committed as profiles/r02_tf32_peak.json)."""
import json
import sys
import time

import torch


def measure(dtype, allow_tf32, n=8192):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    flops = 2.0 * n ** 3
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record()
        torch.cuda.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    reps = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            torch.matmul(a, b)
        reps += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = reps * flops / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return best, sustained


def main():
    out = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "how": "torch.matmul 8192^3: fp32 operands with torch.backends.cuda.matmul.allow_tf32=True (TF32 tensor cores through cuBLAS), "
                  "best of 10 (burst) and back to back for 4 s (sustained), CUDA events; bf16 re-measured by the same code"}
    tf_b, tf_s = measure(torch.float32, True)
    bf_b, bf_s = measure(torch.bfloat16, True)
    fp_b, _ = measure(torch.float32, False, n=4096)
    out.update(tf32_tflops=tf_b, tf32_tflops_sustained=tf_s, bf16_tflops=bf_b, bf16_tflops_sustained=bf_s, fp32_simt_tflops=fp_b)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    sys.exit(main())
