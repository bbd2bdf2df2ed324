"""Times the tcgen05 GEMM on the shapes the model launches (CUDA events, tensor-map encode excluded via warm plan reuse is
not available through the op wrapper, so each call includes ~2 host-side encodes; device time is what the events see)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import surgvid_b200  # noqa
from surgvid_b200 import ops

dev = "cuda:0"
shapes = [  # M, N, K, out_fp32, resid
    (627200, 256, 64, 0, 0), (627200, 64, 64, 0, 0), (627200, 64, 256, 1, 1), (39200, 1280, 320, 0, 0), (39200, 320, 1280, 1, 1),
    (39200, 320, 320, 0, 0), (39200, 320, 320, 1, 1), (156800, 512, 128, 0, 0), (9800, 2048, 8192, 1, 0), (75776, 1280, 320, 0, 0),
    (156800, 1280, 320, 0, 0), (156800, 320, 1360, 1, 1), (156800, 320, 320, 1, 1), (156800, 320, 320, 0, 0), (450800, 320, 1360, 1, 1),
    (450800, 1280, 320, 0, 0), (39200, 2048, 8192, 1, 0), (112700, 512, 2176, 1, 1), (112700, 2048, 512, 0, 0),
]
if len(sys.argv) > 1:
    shapes = [shapes[int(i)] for i in sys.argv[1].split(",")]
reps = int(os.environ.get("REPS", "5"))
for (M, N, K, o32, res) in shapes:
    a = torch.randn(M, K, device=dev).bfloat16()
    w = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if o32 else torch.bfloat16)
    resid = torch.randn(M, N, device=dev) if res else None
    for _ in range(2):
        ops.gemm_bf16(a, w, bias=bias, residual=resid, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.gemm_bf16(a, w, bias=bias, residual=resid, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    by = M * K * 2 + N * K * 2 + M * N * (4 if o32 else 2) + (M * N * 4 if res else 0)
    print(f"M={M} N={N} K={K} o32={o32} res={res}: {us:8.1f} us  {2*M*N*K/us/1e6:7.1f} TF/s  {by/us/1e3:7.1f} GB/s", flush=True)
