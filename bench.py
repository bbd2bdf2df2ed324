#!/usr/bin/env python
"""bench.py — LFB frames/s of the B200-native path (MiT-EVP mit_b3_evp + SegFormer embedding head with flow,
return_features, then MS-TCN MultiStageModel_S over the extracted features).

  python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N>1, one rank per GPU, no collective on the data path)
  python bench.py --impl reference --gpus N --steps K ...   # reference arm: the reference's algorithm on the host cores (oracle port)

A "step" = one synthetic Cholec80-length video (2 300 frames, 224x224; BASELINE.json configs[1]) per GPU: LFB extraction in
the reference's batches of 200 frames (generate_evp_LFB.py:36) followed by MS-TCN over the 2 300 features
(trans_SV_output.py:268-280).  Weak scaling: every rank processes its own video per step; videos are independent.
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

FRAMES_PER_VIDEO = 2300
BATCH = 800  # throughput-optimal on B200; the reference driver uses 200 (generate_evp_LFB.py:36), see DESIGN.md for both
FLOPS_PER_FRAME_REF = 16.664e9  # reference graph @224^2 with flow


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES_PER_VIDEO)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--micro-batch", type=int, default=int(os.environ.get("SURGVID_MICRO_BATCH", "800")))
    ap.add_argument("--fold-head", type=int, default=int(os.environ.get("SURGVID_FOLD_HEAD", "0")))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms; the process is started one warm-up step early (it needs
    ~0.2 s to deliver its first line) and only samples that arrive between mark_begin() and mark_end() are used."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0, self.t1 = None, None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (ts, ln) in self.lines if (self.t0 is None or ts >= self.t0) and (self.t1 is None or ts <= self.t1 + 0.05)]
        if not inside:  # a very short timed region: fall back to every sample taken under the same load
            inside = [ln for (_, ln) in self.lines]
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_port_frames_per_sec(seconds, batch=8, frames_per_video=FRAMES_PER_VIDEO, steps=None, warmup=1):
    """The reference's algorithm for this path on the host cores: oracle port (plain PyTorch fp32 restatement, pinned to the
    reference's outputs by tests/test_oracle_cpu.py), all host threads.  Sample: `batch` frames through the encoder+head with
    flow, plus MS-TCN over `frames_per_video` features; frames/s = 1 / (t_enc/batch + t_tcn/frames_per_video)."""
    import surgvid_b200  # noqa: F401
    from oracle import evp_oracle, mstcn_oracle
    from surgvid_b200 import synthetic as S

    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    sd = S.synth_state_dict(S.evp_key_shapes("mit_b3_evp"), seed=0, mode="ref_init")
    msd = S.synth_mstcn_state_dict(mode="phase")
    cfg = S.EVP_CONFIGS["mit_b3_evp"]
    x, seg, flow = S.synth_frames(batch, seed=7)
    lfb = S.synth_lfb_features(frames_per_video, seed=3).unsqueeze(0).transpose(2, 1)

    def one():
        t0 = time.perf_counter()
        evp_oracle.evp_forward(sd, cfg, x, seg, flow)
        t1 = time.perf_counter()
        mstcn_oracle.mstcn_forward(msd, lfb)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    for _ in range(warmup):
        one()
    enc, tcn, n = 0.0, 0.0, 0
    t_start = time.perf_counter()
    while True:
        a, b = one()
        enc += a; tcn += b; n += 1
        if steps is not None and n >= steps:
            break
        if steps is None and time.perf_counter() - t_start >= seconds:
            break
    per_frame = enc / (n * batch) + tcn / (n * frames_per_video)
    return {"value": 1.0 / per_frame, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{n} x (oracle encoder+head fwd on {batch} frames 224x224 fp32 with flow + MS-TCN over {frames_per_video} features), "
                      f"{enc + tcn:.1f} s of CPU work", "ms_per_step": 1e3 * (enc + tcn) / n}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_port_frames_per_sec(0.0, steps=max(1, args.steps), warmup=max(1, args.warmup))
    line = {"impl": "reference", "metric": "lfb_frames_per_sec", "value": r["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"LFB extraction (mit_b3_evp + head, flow) + MS-TCN, one {args.frames}-frame synthetic Cholec80-length video per GPU per step, 224x224",
                       "arm": "reference algorithm on host cores (oracle port; /root/reference is absent on the GPU box)"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch.distributed as dist

    import surgvid_b200  # noqa: F401
    from surgvid_b200 import _native, lfb
    from surgvid_b200 import synthetic as S
    from surgvid_b200.models.mix_transformer_evp import mit_b3_evp
    from surgvid_b200.mstcn import MultiStageModel_S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    T, B = args.frames, args.batch

    model = mit_b3_evp()
    model.load_state_dict(S.synth_state_dict(S.evp_key_shapes("mit_b3_evp"), seed=0, mode="ref_init"), strict=True)
    model.micro_batch, model.fold_head = args.micro_batch, bool(args.fold_head)
    model = model.to(dev).eval()
    tcn = MultiStageModel_S(2, 8, 32, 2048, 14, True)
    tcn.load_state_dict(S.synth_mstcn_state_dict(mode="phase"), strict=True)
    tcn = tcn.to(dev).eval()

    # synthetic video resident in HBM (3.7 GB fp32 per rank, far larger than the 126 MB L2), seeded per rank
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn((T, 3, 224, 224), device=dev, generator=g)
    mask = (torch.rand((T, 1, 224, 224), device=dev, generator=g) > 0.7).float().expand(T, 3, 224, 224)
    seg = ((mask - torch.tensor(S.NORM_MEAN, device=dev).view(1, 3, 1, 1)) / torch.tensor(S.NORM_STD, device=dev).view(1, 3, 1, 1)).contiguous()
    del mask
    flow = 2.0 * torch.randn((T, 2, 224, 224), device=dev, generator=g)
    feats = torch.empty((T, 2048), dtype=torch.float32, device=dev)
    launches = {"n": 0}

    @torch.no_grad()
    def step():
        n = 0
        for b0 in range(0, T, B):
            b1 = min(T, b0 + B)
            feats[b0:b1] = model(x[b0:b1], seg[b0:b1], flow[b0:b1], return_features=True)
            n += model.last_launch_count(dev)
        logits = tcn.forward_videos(feats, [T])
        n += tcn.last_launch_count(dev)
        launches["n"] = n
        return logits

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    nwarm = max(3, args.warmup)
    for i in range(nwarm):
        if i == nwarm - 1 and rank == 0:
            sampler.start()
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        logits = step()
    e1.record()
    barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * T * args.steps / (ms_total / 1e3)
    launches_per_step = launches["n"]

    # ---- end-to-end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    e2e = None
    numa_cpus = lfb.bind_to_gpu_numa_node(local) if world > 1 else None   # pinned staging memory on the GPU's own NUMA node
    if not args.no_e2e:
        xh, sh, fh = x.cpu().pin_memory(), seg.cpu().pin_memory(), flow.cpu().pin_memory()
        ext = lfb.LFBExtractor(model, batch_size=B, device=dev)
        out_h = torch.empty((T, 2048), dtype=torch.float32).pin_memory()

        e2e_steps = max(1, min(args.steps, 3))
        outs_h = [out_h] + [torch.empty((T, 2048), dtype=torch.float32).pin_memory() for _ in range(e2e_steps - 1)]
        logits_h = torch.empty((2, 14, T * e2e_steps), dtype=torch.float32).pin_memory()

        @torch.no_grad()
        def e2e_run(k):
            # k steps = k videos through the driver-level call: one pipelined pass (H2D of every batch from pinned host memory,
            # forward, D2H of the features), then MS-TCN over the k feature sequences and D2H of the phase logits
            f_hs = ext.extract_videos([(xh, sh, fh)] * k, outs=outs_h[:k])
            feats_d = torch.cat([f.to(dev, non_blocking=True) for f in f_hs], 0)
            lg = tcn.forward_videos(feats_d, [T] * k)
            logits_h[:, :, :T * k].copy_(lg, non_blocking=True)
            torch.cuda.synchronize()

        e2e_run(1)
        barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * T * e2e_steps / float(dt.item()), "unit": "frames/s",
               "h2d_bytes_per_step": int(ext.h2d_bytes // e2e_steps + T * 2048 * 4), "d2h_bytes_per_step": int(ext.d2h_bytes // e2e_steps + 2 * 14 * T * 4),
               "steps": e2e_steps, "host_cores_bound": (len(numa_cpus) if numa_cpus else None),
               "api": "LFBExtractor.extract_videos(model=mit_b3_evp drop-in; the timed steps are videos of ONE pipelined call) + MultiStageModel_S.forward_videos"}
        del xh, sh, fh
        # same call chain from what the reference's dataset class holds after JPEG decode: uint8 250x250 frames and
        # segmentation maps + the raw fp32 RAFT field; Resize/CenterCrop/ToTensor/Normalize and the flow resize run on the GPU
        g = torch.Generator().manual_seed(11 + rank)
        fr_u8 = torch.randint(0, 256, (T, 250, 250, 3), dtype=torch.uint8, generator=g).pin_memory()
        sg_u8 = ((torch.rand((T, 250, 250, 1), generator=g) > 0.7).to(torch.uint8) * 255).expand(T, 250, 250, 3).contiguous().pin_memory()
        fl_raw = (2.0 * torch.randn((T, 250, 250, 2), generator=g)).pin_memory()

        @torch.no_grad()
        def e2e_raw_run(k):
            f_hs = ext.extract_raw_videos([(fr_u8, sg_u8, fl_raw)] * k, outs=outs_h[:k])
            feats_d = torch.cat([f.to(dev, non_blocking=True) for f in f_hs], 0)
            lg = tcn.forward_videos(feats_d, [T] * k)
            logits_h[:, :, :T * k].copy_(lg, non_blocking=True)
            torch.cuda.synchronize()

        e2e_raw_run(1)
        barrier()
        t0 = time.perf_counter()
        e2e_raw_run(e2e_steps)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e["from_uint8_frames"] = {"value": world * T * e2e_steps / float(dt.item()), "unit": "frames/s",
                                    "h2d_bytes_per_step": int(ext.h2d_bytes // e2e_steps + T * 2048 * 4),
                                    "d2h_bytes_per_step": int(ext.d2h_bytes // e2e_steps + 2 * 14 * T * 4), "steps": e2e_steps,
                                    "api": "LFBExtractor.extract_raw_videos(uint8 250x250 frames + segmaps, fp32 250x250 flow; one pipelined call) + MultiStageModel_S.forward_videos"}
        del fr_u8, sg_u8, fl_raw

    # ---- per-kernel-class device timing (CUDA events around every launch, on the launching stream; untimed extra pass)
    roofline, classes = None, None
    if rank == 0:
        peaks = measured_peaks()
        lib = _native.lib()
        h = model._native[local]["handle"]
        lib.sv_evp_set_profile(h, 1)
        with torch.no_grad():
            for b0 in range(0, T, B):
                model(x[b0:min(T, b0 + B)], seg[b0:min(T, b0 + B)], flow[b0:min(T, b0 + B)], return_features=True)
        torch.cuda.synchronize()
        ms_k = (ctypes.c_double * 16)()
        n_k = (ctypes.c_int64 * 16)()
        fl = ctypes.c_double(0)
        by_k = (ctypes.c_double * 16)()
        lib.sv_evp_get_profile(h, ms_k, n_k, ctypes.byref(fl), by_k)
        if os.environ.get("SURGVID_PROFILE_CSV"):
            lib.sv_evp_dump_profile(h, os.environ["SURGVID_PROFILE_CSV"].encode())
        lib.sv_evp_set_profile(h, 0)
        names = ["gemm_tcgen05", "layernorm", "im2col", "dwconv3x3_gelu", "attention", "gauss5x5", "bilinear", "token_mean", "stem_conv"]
        tot = sum(ms_k[i] for i in range(len(names)))
        # per class: device ms, launches, share of the step, algorithmic HBM bytes (operands + results of each launch once)
        # and the HBM bandwidth / fraction of the measured copy peak they imply
        classes = {}
        for i, nm in enumerate(names):
            gbs = by_k[i] / (ms_k[i] / 1e3) / 1e9 if ms_k[i] else None
            classes[nm] = {"ms_per_step": ms_k[i], "launches_per_step": int(n_k[i]), "share": (ms_k[i] / tot if tot else 0.0),
                           "alg_mbytes_per_frame": by_k[i] / T / 1e6, "hbm_gbs": gbs, "hbm_frac": (gbs / peaks["hbm_gbs"] if gbs else None)}
        gemm_ms_per_launch = ms_k[0] / max(1, n_k[0])
        tflops = fl.value / (ms_k[0] / 1e3) / 1e12 if ms_k[0] else 0.0
        gbs = by_k[0] / (ms_k[0] / 1e3) / 1e9 if ms_k[0] else 0.0
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r01", "gemm_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, traffic_src = tj.get("traffic_bytes_per_launch"), tj.get("source")
        # The dominant kernel is the tcgen05 GEMM.  On this workload its launches have an aggregate arithmetic intensity of
        # ~130 FLOP/B (K = 64..320 for most of them) against a ridge of ~213 FLOP/B, so the BINDING roofline is HBM; the tensor-pipe
        # figures are reported alongside.
        roofline = {"bound": "hbm", "kernel": "gemm_bf16_tcgen05_kernel", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": gbs / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peaks["source"] + " (copy bandwidth)",
                    "avg_launch_ms": gemm_ms_per_launch, "launches_per_step": int(n_k[0]),
                    "algorithmic_bytes_per_launch": by_k[0] / max(1, n_k[0]), "algorithmic_bytes_per_frame": by_k[0] / T,
                    "share_of_step_device_time": classes["gemm_tcgen05"]["share"],
                    "tensor": {"achieved": tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                               "frac": tflops / peaks["bf16_tflops_sustained"], "algorithmic_flops_per_step": fl.value,
                               "executed_gemm_flops_per_frame": fl.value / T,
                               "peak_note": "sustained cuBLAS bf16 figure of MEASURED_PEAKS.json (kernel timed inside a long step)"}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_frames_per_sec(args.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        peaks = measured_peaks()
        line = {"metric": "lfb_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": f"LFB extraction (mit_b3_evp encoder + SegFormer head with flow, return_features) + MS-TCN MultiStageModel_S(2,8,32,2048,14), "
                                       f"one {T}-frame synthetic Cholec80-length video per GPU per step, 224x224 (BASELINE.json configs[1])",
                           "frames_per_step_per_gpu": T, "batch": B, "micro_batch": args.micro_batch, "fold_head": int(args.fold_head),
                           "weights": "random init (reference distributions), seed 0", "parallelism": f"{world} x independent video shards, no collective",
                           "l2": "inputs (3.7 GB fp32 per GPU) are far larger than the 126 MB L2; no flush needed"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
                "roofline": roofline, "cpu_baseline": cpu, "kernel_classes": classes,
                "tensor_roofline_whole_path": {"ref_graph_flops_per_frame": FLOPS_PER_FRAME_REF,
                                               "achieved_tflops_ref_graph": value / world * FLOPS_PER_FRAME_REF / 1e12,
                                               "frac_of_sustained_peak": value / world * FLOPS_PER_FRAME_REF / 1e12 / peaks["bf16_tflops_sustained"]}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
