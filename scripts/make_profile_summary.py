"""Summarise ncu outputs of scripts/gpu_profile.sh into profiles/ (tracked).  Usage:
   python scripts/make_profile_summary.py <tag> <out_name> [kernel-regex-note]"""
import csv, os, subprocess, sys, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles"); os.makedirs(P, exist_ok=True)
lines = [f"# {out} — ncu summary (tag {tag})", ""]
# launch list
lp = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(P, f"{out}_launches.csv"))
    hdr = None; rows = []
    for r in csv.reader(open(lp)):
        if r and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr): rows.append(r)
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    seq = [(r[ki], r[gi], float(r[vi].replace(",", "")) / 1000.0) for r in rows]
    idx = [i for i, s in enumerate(seq) if "normalize_rows" in s[0]]
    if len(idx) > 5:
        a, b = idx[4], idx[5]
        step = [s for s in seq[a:b] if "rvo::" in s[0]]
        tot = sum(s[2] for s in step)
        lines += ["## Launch list of one timed step (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
                  "| kernel | grid | µs | share |", "|---|---|---:|---:|"]
        for s in step:
            lines.append(f"| `{s[0][:70]}` | {s[1]} | {s[2]:.1f} | {100 * s[2] / tot:.1f}% |")
        lines += [f"| **sum of library kernels** | | **{tot:.1f}** | |", ""]
# full capture
rp = os.path.join(G, f"{tag}_scan.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
            "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic"]
    lines += ["## `ncu --set full --clock-control none` (one value per captured launch)", "", "| metric | unit | values |", "|---|---|---|"]
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                lines.append(f"| {w} | {units[i]} | " + " ; ".join(r[i][:60] for r in rows[2:]) + " |")
    lines.append("")
pl = os.path.join(G, f"{tag}_plain.log")
if os.path.exists(pl):
    last = open(pl).read().strip().splitlines()[-1]
    lines += ["## bench.py line of the plain (un-profiled) run of the same command", "", "```json", last, "```", ""]
open(os.path.join(P, f"{out}.md"), "w").write("\n".join(lines))
print("wrote", os.path.join(P, f"{out}.md"))
