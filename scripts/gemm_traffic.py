"""Derive per-launch DRAM traffic and per-class duration shares from an ncu launch list
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv`).
usage: python scripts/gemm_traffic.py <launches.csv> [out.json] [source note]"""
import collections, csv, json, os, sys

def gemm_kernel_sha256(root):
    """sha256 of the GEMM's device + plan code (gemm_tcgen05.cu up to the extern "C" op wrappers, gemm.cuh, gemm_epi.cuh, ptx.cuh):
    recorded next to the ncu-measured traffic so that a bench run can tell whether that capture still describes the kernel it runs."""
    import hashlib, os
    d = os.path.join(root, "deep-learning-for-surgical-video-analysis_b200", "csrc")
    h = hashlib.sha256()
    src = open(os.path.join(d, "gemm_tcgen05.cu")).read()
    h.update(src.split('extern "C"')[0].encode())
    for f in ("gemm.cuh", "gemm_epi.cuh", "ptx.cuh"):
        h.update(open(os.path.join(d, f)).read().encode())
    return h.hexdigest()


path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) >= 15 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    d = per.setdefault(int(r[0]), {"kernel": r[4]})
    d[r[12]] = float(r[14]) * (1e-3 if r[13] == "ns" and False else 1.0)
    d.setdefault("units", {})[r[12]] = r[13]
cls = collections.OrderedDict()
def klass(name):
    for k in ("gemm_bf16_tcgen05", "layernorm", "im2col", "dwconv3x3_gelu", "attention", "gauss5x5", "bilinear", "token_mean", "stem_conv", "mstcn", "classify", "prep_"):
        if k in name: return k
    return "other"
tot = 0.0
for i, d in per.items():
    k = klass(d["kernel"])
    c = cls.setdefault(k, {"launches": 0, "time": 0.0, "rd": 0.0, "wr": 0.0})
    c["launches"] += 1
    t = d.get("gpu__time_duration.sum", 0.0)
    if d["units"].get("gpu__time_duration.sum") == "us": t *= 1e3
    if d["units"].get("gpu__time_duration.sum") == "ms": t *= 1e6
    c["time"] += t; tot += t
    c["rd"] += d.get("dram__bytes_read.sum", 0.0); c["wr"] += d.get("dram__bytes_write.sum", 0.0)
print(f"{len(per)} launches, {tot/1e6:.2f} ms total under ncu (cold-cache, serialised)")
for k, c in cls.items():
    print(f"  {k:20s} launches {c['launches']:4d}  time share {c['time']/tot:.3f}  dram rd {c['rd']/1e9:7.2f} GB  wr {c['wr']/1e9:7.2f} GB")
g = cls.get("gemm_bf16_tcgen05")
if g and len(sys.argv) > 2:
    frames = int(sys.argv[4]) if len(sys.argv) > 4 else 800   # frames of the profiled micro-batch
    out = {"kernel": "gemm_bf16_tcgen05_kernel", "traffic_bytes_per_launch": int((g["rd"] + g["wr"]) / g["launches"]), "launches": g["launches"],
           "frames_profiled": frames, "traffic_bytes_per_frame": (g["rd"] + g["wr"]) / frames,
           "dram_bytes_read": g["rd"], "dram_bytes_write": g["wr"], "source": sys.argv[3] if len(sys.argv) > 3 else path,
           "class_time_shares_under_ncu": {k: round(c["time"] / tot, 4) for k, c in cls.items()},
           "gemm_kernel_sha256": gemm_kernel_sha256(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print("wrote", sys.argv[2], out["traffic_bytes_per_launch"])
