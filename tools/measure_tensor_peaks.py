"""Measured dense tensor throughput of this B200 next to MEASURED_PEAKS.json's bf16 figure (GPU box):
cuBLAS fp16 and TF32 GEMMs 8192^3, best of 10 (burst) with CUDA events -- the denominators for the tensor-bound rows of DESIGN.md
section 4 (the tcgen05 engine runs kind::f16 3-pass where the channel count allows and kind::tf32 3-pass elsewhere).
   python tools/measure_tensor_peaks.py > profiles/r2_measured_tensor_peaks.json"""
import json
import torch

dev = torch.device("cuda:0")
N = 8192
out = {"gpu": torch.cuda.get_device_name(0), "n": N, "how": "torch.matmul N^3, 2*N^3 flop, best of 10 after 3 warm-ups, CUDA events"}
for name, dt, tf32 in (("fp16", torch.float16, False), ("bf16", torch.bfloat16, False), ("tf32", torch.float32, True), ("fp32_simt", torch.float32, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(N, N, device=dev, dtype=dt)
    b = torch.randn(N, N, device=dev, dtype=dt)
    for _ in range(3):
        a @ b
    best = 1e9
    for _ in range(10 if name != "fp32_simt" else 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name + "_tflops"] = round(2 * N ** 3 / (best * 1e-3) / 1e12, 1)
    del a, b
print(json.dumps(out))
