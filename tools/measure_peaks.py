"""cuBLAS TF32 and FP64 GEMM peaks on this B200, measured the way the driver measured MEASURED_PEAKS.json's bf16
figures: torch.matmul N^3 (2 N^3 flop), best of 10 (burst) and back to back for 4 s (sustained), CUDA events.
    gpurun -- python tools/measure_peaks.py > gpurun_out/peaks.json     (then copy to profiles/measured_peaks_tf32_f64.json)
The driver's file has no TF32 / FP64 entry; bench.py's roofline denominators for the TF32 panel and the FP64 fit come
from the stored copy so that roofline.frac is comparable from run to run."""
import json
import time

import torch


def measure(dtype, tf32, n, seconds=4.0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, dtype=dtype, device="cuda")
    b = torch.randn(n, n, dtype=dtype, device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(8, int(seconds * 1e3 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / reps
    flop = 2.0 * n ** 3
    return flop / best / 1e9, flop / sus / 1e9, reps


def main():
    out = {"gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "how": "torch.matmul N^3 (2*N^3 flop), best of 10 (burst) and back to back for 4 s (sustained), CUDA events: "
                  "tf32 = float32 inputs with allow_tf32 at N=8192, f64 at N=4096 (MEASURED_PEAKS.json method)"}
    b, s, r = measure(torch.float32, True, 8192)
    out.update(tf32_tflops=b, tf32_tflops_sustained=s, tf32_sustained_launches=r)
    b, s, r = measure(torch.float64, False, 4096)
    out.update(f64_tflops=b, f64_tflops_sustained=s, f64_sustained_launches=r)
    b, s, r = measure(torch.bfloat16, False, 8192)
    out.update(bf16_tflops=b, bf16_tflops_sustained=s)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
