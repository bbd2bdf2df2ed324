"""First-contact diagnostics for the tcgen05 GEMM on a real B200: small shapes, structured operands, verbose diffs."""
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import surgvid_b200  # noqa: F401
from surgvid_b200 import ops

dev = "cuda:0"
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))


def run(M, N, K, structured=False, **kw):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    if structured:
        a = torch.zeros(M, K)
        a[torch.arange(M), torch.arange(M) % K] = 1.0
        w = (torch.arange(N)[:, None] * 0.5 + torch.arange(K)[None, :] * 0.001953125).float()
    else:
        a = torch.randn(M, K, generator=g)
        w = torch.randn(N, K, generator=g) / math.sqrt(K)
    a, w = a.to(dev).bfloat16(), w.to(dev).bfloat16()
    ref = a.float() @ w.float().t()
    out = ops.gemm_bf16(a, w, out_dtype=torch.float32, **kw)
    torch.cuda.synchronize()
    err = (out - ref).abs()
    print(f"GEMM M={M} N={N} K={K} structured={structured}: max err {err.max().item():.4e} (max|ref| {ref.abs().max().item():.3f})", flush=True)
    if err.max().item() > 1e-2 * max(1.0, ref.abs().max().item()):
        bad = (err > 1e-2 * max(1.0, ref.abs().max().item()))
        print("  bad fraction", bad.float().mean().item(), "bad rows", bad.any(1).sum().item(), "bad cols", bad.any(0).sum().item())
        print("  out[0:4,0:8]\n", out[0:4, 0:8].cpu(), "\n  ref[0:4,0:8]\n", ref[0:4, 0:8].cpu())
        print("  out[32:34,0:8]\n", out[32:34, 0:8].cpu(), "\n  ref[32:34,0:8]\n", ref[32:34, 0:8].cpu())
        return False
    return True


ok = True
for shp in [(128, 64, 64), (128, 16, 16), (128, 64, 128), (128, 256, 64), (256, 128, 256), (1000, 320, 1280), (200, 64, 152), (129, 24, 72)]:
    ok &= run(*shp, structured=True)
    ok &= run(*shp)
# timing sanity on a big problem
M, N, K = 148 * 128 * 4, 1280, 320
a = torch.randn(M, K, device=dev).bfloat16()
w = torch.randn(N, K, device=dev).bfloat16()
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm_bf16(a, w, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.gemm_bf16(a, w, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"GEMM {M}x{N}x{K}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s (includes per-call tensor-map encode)")
ref = (a[:512].float() @ w.float().t())
print("  spot err", (out[:512].float() - ref).abs().max().item())
print("FIRST CONTACT", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
