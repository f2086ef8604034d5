"""cuBLAS GEMM rates on this GPU: the denominators for the f64 / tf32 / fp32 GEMM-class steps."""
import json, sys, torch
def rate(dtype, n, tf32=False, reps=5):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, dtype=dtype, device="cuda"); b = torch.randn(n, n, dtype=dtype, device="cuda")
    for _ in range(2): a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * n ** 3 / best / 1e9
out = {"f64_tflops": rate(torch.float64, 8192), "f64_4096_tflops": rate(torch.float64, 4096),
       "tf32_tflops": rate(torch.float32, 8192, tf32=True), "fp32_tflops": rate(torch.float32, 8192, tf32=False),
       "gpu": torch.cuda.get_device_name(0)}
print(json.dumps(out))
