"""Key metrics per kernel instance out of `ncu -i <rep> --page raw --csv` (stdin or file): duration, DRAM bytes / throughput %, L2 throughput %,
tensor-pipe and issue utilisation, occupancy, registers, top stall reasons.  usage: python scripts/ncu_summary.py raw.csv [out.csv]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
want = [
    ("dur_us", r"^gpu__time_duration\.sum$"),
    ("dram_rd_MB", r"^dram__bytes_read\.sum$"), ("dram_wr_MB", r"^dram__bytes_write\.sum$"),
    ("dram_pct", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$"),
    ("l2_pct", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$"),
    ("l2_to_sm_MB", r"^l1tex__m_xbar2l1tex_read_bytes\.sum$"),
    ("l2_hit_pct", r"lts__t_sector_hit_rate\.pct$"),
    ("sm_pct", r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed$"),
    ("tensor_pct", r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_elapsed$"),
    ("tc_pct", r"^sm__pipe_tc_cycles_active\.avg\.pct_of_peak_sustained_elapsed$"),
    ("lsu_wavefront_pct", r"^l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed$"),
    
    ("fma_pct", r"^sm__pipe_fma_cycles_active\.avg\.pct_of_peak_sustained_elapsed$"),
    ("alu_pct", r"^sm__pipe_alu_cycles_active\.avg\.pct_of_peak_sustained_elapsed$"),
    ("xu_pct", r"^sm__inst_executed_pipe_xu\.avg\.pct_of_peak_sustained_active$"),
    ("issue_pct", r"smsp__issue_active\.avg\.pct"), ("ipc", r"sm__inst_executed\.avg\.per_cycle_elapsed$"),
    ("warps_active_pct", r"sm__warps_active\.avg\.pct_of_peak_sustained_active$"),
    ("regs", r"launch__registers_per_thread$"), ("smem_dyn_KB", r"launch__shared_mem_per_block_dynamic$"),
    ("occ_limit_regs", r"launch__occupancy_limit_registers$"), ("occ_limit_smem", r"launch__occupancy_limit_shared_mem$"),
    ("sm_mhz", r"sm__cycles_elapsed\.avg\.per_second$"),
]
stall = [(i, n) for i, n in enumerate(names) if "smsp__average_warp" in n and "issue_stalled" in n and n.endswith("_per_warp_active.pct") or
         ("smsp__average_warps_issue_stalled" in n and n.endswith("per_issue_active.ratio"))]
cols = []
for key, pat in want:
    idx = [i for i, n in enumerate(names) if re.search(pat, n)]
    cols.append((key, idx[0] if idx else None))
out = []
for r in data:
    if len(r) != len(names):
        continue
    d = {"id": r[0], "kernel": re.sub(r"^void |\(.*$", "", r[4])[-60:], "grid": r[8], "block": r[7]}
    for key, i in cols:
        if i is None or r[i] in ("", "n/a"):
            d[key] = ""
            continue
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            d[key] = ""
            continue
        u = units[i]
        if key.endswith("_MB"):
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        if key == "dur_us":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        if key == "sm_mhz":
            v *= {"Ghz": 1e3, "Mhz": 1.0, "hz": 1e-6}.get(u, 1.0)
        d[key] = round(v, 3)
    st = []
    for i, n in stall:
        try:
            st.append((float(r[i]), re.sub(r".*issue_stalled_|_per_warp_active\.pct|_per_issue_active\.ratio", "", n)))
        except ValueError:
            pass
    d["top_stalls"] = " ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:4])
    out.append(d)
keys = ["id", "kernel", "grid", "block"] + [k for k, _ in want] + ["top_stalls"]
w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
w.writerow(keys)
for d in out:
    w.writerow([d.get(k, "") for k in keys])
