"""Library yardsticks on this box (NOT part of the product, never called by it): cuBLAS SGEMM with TF32 off (CUDA cores), with TF32 on
(1xTF32 tensor cores) and bf16, 8192^3, through torch.matmul — what NVIDIA's own kernels reach next to this repository's."""
import json
import torch

n = 8192
out = {}
for name, dtype, tf32 in (("cublas_fp32_cuda_cores", torch.float32, False), ("cublas_1xtf32", torch.float32, True), ("cublas_bf16", torch.bfloat16, True)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = (torch.rand((n, n), device="cuda") * 2 - 1).to(dtype)
    b = (torch.rand((n, n), device="cuda") * 2 - 1).to(dtype)
    for _ in range(3):
        c = a @ b
    best, tot, iters = 1e9, 0.0, 20
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        c = a @ b
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best, tot = min(best, ms), tot + ms
    fl = 2.0 * n ** 3
    out[name] = {"tflops_best": round(fl / best / 1e9, 1), "tflops_mean": round(fl / (tot / iters) / 1e9, 1)}
print(json.dumps(out))
