"""TF32 tensor peak of this B200 measured the way MEASURED_PEAKS.json measures bf16 (BASELINE.md section 2 leaves it open):
torch.matmul on 8192^3 fp32 operands with TF32 allowed (cuBLAS kind::tf32), best of 10 (burst) and back to back for 4 s
(sustained), plus the same loop in bf16 on the same box for the ratio.  Prints one JSON line."""
import json
import time

import torch


def run(dtype, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=dtype)
    b = torch.randn(n, n, device="cuda", dtype=dtype)
    flops = 2.0 * n ** 3
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    t0 = time.time(); k = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(20):
            a @ b
        k += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return flops / best / 1e9, flops * k / e0.elapsed_time(e1) / 1e9


if __name__ == "__main__":
    tf_b, tf_s = run(torch.float32, True)
    bf_b, bf_s = run(torch.bfloat16, False)
    print(json.dumps({"tf32_tflops": tf_b, "tf32_tflops_sustained": tf_s, "bf16_tflops": bf_b, "bf16_tflops_sustained": bf_s,
                      "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
                      "how": "torch.matmul 8192^3, best of 10 (burst) and back to back for 4 s (sustained), CUDA events"}))
