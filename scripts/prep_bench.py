"""Timing of the on-GPU input transforms (uint8 250x250 frames + segmaps -> fp32 224x224, raw flow -> 224x224) for one batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import surgvid_b200  # noqa
from surgvid_b200.preprocess import FramePreprocessor
dev = "cuda:0"
n = int(os.environ.get("N", "1150"))
prep = FramePreprocessor((250, 250), flow_hw=(250, 250), resize=250, crop=224)
fu = torch.randint(0, 256, (n, 250, 250, 3), dtype=torch.uint8, device=dev)
fl = torch.randn(n, 250, 250, 2, device=dev)
x = torch.empty(n, 3, 224, 224, device=dev); f = torch.empty(n, 2, 224, 224, device=dev)
def t(fn, name, nbytes):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms for {n} frames  ({nbytes / ms / 1e6:.0f} GB/s of in+out bytes)")
t(lambda: prep.images(fu, out=x), "images (uint8 HWC -> fp32 CHW, resize+crop+normalize)", fu.numel() + x.numel() * 4)
t(lambda: prep.flow(fl, out=f), "flow (fp32 HWC -> fp32 CHW, resize+crop+rescale)", fl.numel() * 4 + f.numel() * 4)
