"""Stand-alone timing of the non-GEMM kernels at the shapes of one 200-frame micro-batch (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import surgvid_b200  # noqa
from surgvid_b200 import ops
dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(os.environ.get("REPS", "5"))

def timeit(name, fn, nbytes):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{name}: {us:8.1f} us  {nbytes/us/1e3:7.1f} GB/s", flush=True)

B = int(os.environ.get("B", "200"))
if which in ("all", "dwconv"):
    for (H, W, C) in [(14, 14, 1280), (56, 56, 256), (28, 28, 512), (7, 7, 2048)]:
        x = torch.randn(B, H, W, C, device=dev).bfloat16(); w = torch.randn(9, C, device=dev); b = torch.randn(C, device=dev)
        timeit(f"dwconv {H}x{W}x{C}", lambda: ops.dwconv3x3_gelu(x, w, b), 2 * x.numel() * 2)
if which in ("all", "ln"):
    for (rows, C) in [(39200, 320), (627200, 64), (156800, 128), (9800, 512)]:
        x = torch.randn(rows, C, device=dev); g = torch.ones(C, device=dev); bb = torch.zeros(C, device=dev)
        timeit(f"layernorm {rows}x{C}", lambda: ops.layernorm(x, g, bb, 1e-6), x.numel() * 6)
if which in ("all", "attn"):
    for (heads, Nq, Nkv, hd) in [(5, 196, 49, 64), (1, 3136, 49, 64), (2, 784, 49, 64), (8, 49, 49, 64), (8, 196, 196, 40)]:
        C = heads * hd
        q = torch.randn(B * Nq, C, device=dev).bfloat16(); kv = torch.randn(B * Nkv, 2 * C, device=dev).bfloat16()
        timeit(f"attention h{heads} Nq{Nq} Nkv{Nkv} hd{hd}", lambda: ops.attention(q, kv[:, :C], kv[:, C:], B, heads, hd, hd ** -0.5),
               (2 * q.numel() + kv.numel()) * 2)
if which in ("all", "im2col"):
    x = torch.randn(B, 3, 224, 224, device=dev)
    timeit("im2col nchw 3x224x224 k7s4", lambda: ops.im2col(x, 7, 4, 3), x.numel() * 4 + B * 3136 * 152 * 2)
    t = torch.randn(B, 14, 14, 320, device=dev).bfloat16()
    timeit("im2col nhwc 14x14x320 k2s2", lambda: ops.im2col(t, 2, 2, 0), 2 * t.numel() * 2)
    t = torch.randn(B, 56, 56, 64, device=dev).bfloat16()
    timeit("im2col nhwc 56x56x64 k3s2p1", lambda: ops.im2col(t, 3, 2, 1), t.numel() * 2 + B * 784 * 576 * 2)
    timeit("gauss5x5 600x224x224", lambda: ops.gauss5x5(x), 2 * x.numel() * 4)
if which in ("all", "stem"):
    x = torch.randn(B, 3, 224, 224, device=dev)
    w = torch.randn(64, 3, 7, 7, device=dev) * 0.1; bb = torch.randn(64, device=dev) * 0.1; g = torch.ones(64, device=dev); be = torch.zeros(64, device=dev)
    timeit("stem conv 3->64 + LN (fp32+bf16 out)", lambda: ops.stem_conv(x, w, bb, g, be), x.numel() * 4 + B * 3136 * 64 * 6)
    w16 = torch.randn(16, 3, 7, 7, device=dev) * 0.1; b16 = torch.randn(16, device=dev) * 0.1; g16 = torch.ones(16, device=dev); be16 = torch.zeros(16, device=dev)
    timeit("stem conv 3->16 + LN", lambda: ops.stem_conv(x, w16, b16, g16, be16), x.numel() * 4 + B * 3136 * 16 * 6)
