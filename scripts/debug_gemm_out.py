import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, surgvid_b200
from surgvid_b200 import ops
dev = "cuda:0"
def run(M, N, K, ldc_mult, col_off, lda_pad=0):
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn(M, K + lda_pad, device=dev, generator=g).bfloat16()[:, :K]
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    big = torch.zeros((M, ldc_mult * N), dtype=torch.bfloat16, device=dev)
    out = big[:, col_off * N:(col_off + 1) * N]
    ops.gemm_bf16(a, w, out=out)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    err = (out.float() - ref).abs()
    bad = err > 3e-2
    rows = bad.any(1).nonzero().flatten().tolist()
    cols = bad.any(0).nonzero().flatten().tolist()
    outside = big.clone(); outside[:, col_off * N:(col_off + 1) * N] = 0
    print(f"M={M} N={N} K={K} ldc={ldc_mult*N} off={col_off*N} lda_pad={lda_pad}: max err {err.max().item():.3f} bad rows {len(rows)} {rows[:6]}..{rows[-3:]} bad cols {cols[:4]}..{cols[-4:] if cols else []} wrote outside slice: {bool((outside != 0).any())}")
run(777, 64, 256, 1, 0)
run(777, 64, 256, 4, 2)
run(777, 64, 256, 4, 0)
run(777, 64, 256, 4, 2, 24)
run(777, 64, 256, 1, 0, 24)
run(768, 64, 256, 4, 2)
run(777, 128, 128, 2, 1)
run(1000, 256, 64, 1, 0)
run(1000, 512, 128, 1, 0)
run(3136*2, 48, 16, 1, 0)
run(3136*2, 128, 32, 1, 0)
