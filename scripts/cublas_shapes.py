"""cuBLAS fp16 throughput on the GEMM shapes of the heavy conv layers (reference point for the conv kernel)."""
import torch
shapes = [(173056, 256, 1152), (43264, 512, 2304), (10816, 1024, 4608), (20480, 256, 2304), (20480, 1024, 4608), (8192, 8192, 8192)]
for M, N, K in shapes:
    a = torch.randn(M, K, device="cuda", dtype=torch.float16)
    b = torch.randn(N, K, device="cuda", dtype=torch.float16)
    for _ in range(3):
        c = a @ b.t()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        c = a @ b.t()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"cublas M={M} N={N} K={K}: {ms:.4f} ms  {2.0 * M * N * K / ms * 1e-9:.1f} TFLOP/s", flush=True)
