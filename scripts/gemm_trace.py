"""Cycle-level timeline of CTA 0 of the tcgen05 GEMM (needs lib/libsurgvid_trace.so = gemm_tcgen05.cu built with -DSV_GEMM_TRACE).
Per tile: when the MMA warp got the accumulator, when each k-block's operands had landed, when the producer got each ring slot back,
when the epilogue saw the accumulator full and when it had finished the tile.  usage: gemm_trace.py M N K out_fp32 resid [tiles]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import surgvid_b200
from surgvid_b200 import _native, ops
_native.LIB_PATH = os.path.join(os.path.dirname(_native.LIB_PATH), "libsurgvid_trace.so")
M, N, K, o32, res = [int(v) for v in sys.argv[1:6]]
show = int(sys.argv[6]) if len(sys.argv) > 6 else 10
dev = "cuda:0"
a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16(); bias = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=torch.float32 if o32 else torch.bfloat16)
resid = torch.randn(M, N, device=dev) if res else None
for _ in range(3): ops.gemm_bf16(a, w, bias=bias, residual=resid, out=out)
torch.cuda.synchronize()
buf = np.zeros((4, 4096), dtype=np.uint64)
lib = _native.lib()
lib.sv_debug_gemm_trace.argtypes = [ctypes.c_void_p]
assert lib.sv_debug_gemm_trace(buf.ctypes.data) == 0
num_kb = (K + 63) // 64
t0 = int(buf[2][0])
rel = lambda v: int(v) - t0
print(f"M={M} N={N} K={K} fp32={o32} resid={res}: k-blocks per tile {num_kb}; cycles relative to the first tile's accumulator grant")
ntiles = int((buf[2] > 0).sum())
print("tiles of CTA 0:", ntiles)
prev_done = 0
for t in range(min(show, ntiles)):
    acc = rel(buf[2][t]); land = [rel(buf[1][t * num_kb + k]) for k in range(num_kb)]; slot = [rel(buf[0][t * num_kb + k]) for k in range(num_kb)]
    full, done = rel(buf[3][16 * t]), rel(buf[3][16 * t + 15])
    chunks = [[rel(buf[3][16 * t + 1 + 3 * c + e]) - full for e in range(3)] for c in range(4) if buf[3][16 * t + 1 + 3 * c] > 0]
    print(f"tile {t:2d}: acc granted {acc:7d} | slot free {slot} | operands landed {land} | epilogue: full {full:7d} done {done:7d} (epilogue {done - full}, since prev done {done - prev_done}) chunks [tmem landed, buffer free, store issued] rel. to full: {chunks}")
    prev_done = done
last = min(ntiles, 250) - 1
print(f"steady state: {(rel(buf[3][16 * last + 15]) - rel(buf[3][16 * 4 + 15])) / max(1, last - 4):.0f} cycles per tile (tiles 4..{last})")
