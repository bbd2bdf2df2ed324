"""Fused DWConv+GELU->fc2 kernel (mixffn.cu) against the stand-alone DWConv kernel + tcgen05 GEMM at the shapes of one 800-frame micro-batch."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import surgvid_b200  # noqa
from surgvid_b200 import ops
dev = "cuda:0"
B = int(os.environ.get("FRAMES", "800"))
reps = int(os.environ.get("REPS", "5"))

def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for (H, W, hid, N, tc) in [(14, 14, 1280, 320, 80), (56, 56, 256, 64, 16), (28, 28, 512, 128, 32)]:
    M = B * H * W
    h1 = torch.randn(B, H, W, hid, device=dev).bfloat16()
    w9 = torch.randn(9, hid, device=dev) * 0.3; bdw = torch.randn(hid, device=dev) * 0.1
    wcat = (torch.randn(N, hid + tc, device=dev) / math.sqrt(hid)).bfloat16(); bias = torch.randn(N, device=dev)
    h2 = torch.empty(M, hid + tc, device=dev, dtype=torch.bfloat16)
    x = torch.randn(M, N, device=dev)
    tail = h2[:, hid:]
    def unfused():
        rc = surgvid_b200._native.lib().sv_op_dwconv3x3_gelu(ops._ptr(h1), ops._ptr(w9), ops._ptr(bdw), B, H, W, hid, ops._ptr(h2), ops._stream_ptr(h1.device))
        ops.gemm_bf16(h2, wcat, bias, residual=x, out_dtype=torch.float32, out=x)
    t_dw = timeit(lambda: ops.dwconv3x3_gelu(h1, w9, bdw))
    t_g = timeit(lambda: ops.gemm_bf16(h2, wcat, bias, residual=x, out_dtype=torch.float32, out=x))
    w10 = torch.cat([w9, bdw.view(1, -1)], 0).contiguous()
    lib = surgvid_b200._native.lib()
    def fused():
        rc = lib.sv_op_mixffn_fc2(ops._ptr(h1), ops._ptr(w10), ops._ptr(wcat), wcat.stride(0), ops._ptr(bias), ops._ptr(tail), tail.stride(0), tc, ops._ptr(x),
                                  x.stride(0), B, H, W, hid, N, ops._stream_ptr(h1.device))
        assert rc == 0
    t_f = timeit(fused)
    print(f"{H}x{W} hid{hid} N{N}: dwconv {t_dw:7.1f} us + fc2 gemm {t_g:7.1f} us = {t_dw+t_g:7.1f} us   fused {t_f:7.1f} us   ({(t_dw+t_g)/t_f:.2f}x)", flush=True)
