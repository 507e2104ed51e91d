"""Summarise the round-2 ncu captures in gpurun_out/ (scripts/r2_profiles.sh) into profiles/r02_*.md + profiles/traffic.json."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__inst_executed.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def launch_list(tag, marker, which=1, lib_only=True):
    rows = list(csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))))
    hdr, seq = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            seq.append((r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")], float(r[hdr.index("Metric Value")].replace(",", "")) / 1000.0))
    idx = [i for i, s in enumerate(seq) if marker in s[0]]
    # one step = from one occurrence of the step's first kernel to the next
    a, b = idx[which], idx[which + 1]
    step = [s for s in seq[a:b] if (not lib_only or "rvo::" in s[0])]
    tot = sum(s[2] for s in step)
    out = ["| kernel | grid | µs | share |", "|---|---|---:|---:|"]
    for s in step:
        out.append(f"| `{s[0][:86]}` | {s[1]} | {s[2]:.1f} | {100 * s[2] / tot:.1f}% |")
    out.append(f"| **sum of library kernels** | | **{tot:.1f}** | |")
    return out, tot


def full_table(tag):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, f"{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["| metric | unit | " + " ; ".join(f"launch {i + 1}" for i in range(len(rows) - 2)) + " |", "|---|---|---|"]
    recs = []
    for r in rows[2:]:
        recs.append({h: (v, u) for h, v, u in zip(hdr, r, units)})
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {units[i]} | " + " ; ".join(r[i][:70] for r in rows[2:]) + " |")
    return out, recs


def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]


def last_json(tag):
    p = os.path.join(G, f"{tag}.log")
    if not os.path.exists(p):
        return None
    ls = [x for x in open(p) if x.startswith("{")]
    return ls[-1].strip() if ls else None


def write(name, lines):
    open(os.path.join(P, name), "w").write("\n".join(lines) + "\n")
    print("wrote", name)


traffic = json.load(open(os.path.join(P, "traffic.json")))

# ---- configs[1] step ------------------------------------------------------------------------------------------------
ll, tot = launch_list("k1_cfg1", "normalize_rows", which=4)
chain, _ = full_table("k1_chain")
scan, srec = full_table("h1_cur")
sb = [r for r in srec if "scan_tc2_kernel<1>" in r["Kernel Name"][0]][0]
traffic["cfg1"] = int(to_bytes(*sb["dram__bytes_read.sum"]) + to_bytes(*sb["dram__bytes_write.sum"]))
traffic["_source"] = "profiles/r02_cfg1_step.md: dram__bytes_read.sum + dram__bytes_write.sum of scan_tc2_kernel<1> (ncu --set full, per launch)"
b = json.loads(last_json("k1_cfg1_plain"))
write("r02_cfg1_step.md", [
    "# r02_cfg1_step — configs[1] (1M x 1024 bf16, Q = 256, top-100) after the round-2 changes", "",
    "Command: `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-north-star` (plain first, then under ncu).", "",
    "## Launch list of one timed step (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: compare shares)", "",
    *ll, "",
    "Round 1 (profiles/r01_cfg1_step_final.md): memset + normalise 6.6 + seed scan 15.1 + seed threshold 30.2 (9 with the maxima kernel) + scan 367.1 + "
    "select 48.3 = 467 us.  Now: the memset is folded into the (register-resident) normalise launch, the seed scan writes one maximum per 32 sample "
    "rows instead of every score, the last-level select reads only the HOT sub-lists (the ~600 best survivors per query instead of ~6100), "
    "re-scores one row per warp and ranks by counting.  The same chain as it runs IN THE STREAM, with the gaps between launches: "
    "profiles/r02_chain_timeline.md.", "",
    "## `ncu --set full` of the chain kernels (select_kernel<1> twice, normalise, seed_tau)", "", *chain, "",
    "## `ncu --set full` of the two scan launches of one step (DENSE seed sample and full-shard FILTER, in capture order)", "", *scan, "",
    f"DRAM traffic of the full-shard scan: {traffic['cfg1'] / 1e9:.4f} GB for 2.048 GB algorithmic (x{traffic['cfg1'] / 2.048e9:.3f}).", "",
    "## bench.py line of the plain (un-profiled) run of the same command", "", "```json", json.dumps({k: b[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "roofline", "verify", "results_ok")}), "```", "",
    "## Reading",
    "* The scan is unchanged in substance since round 1 (same DRAM bytes, tensor pipe ~93 % active).  Isolated under ncu it takes ~388 us on "
    "every box; in the stream it takes 355 us right after an idle period and 381-401 us from the second back-to-back step on, depending on the "
    "box: the board's power cap (profiles/r02_chain_timeline.md), the same effect that separates cuBLAS's burst (1638 TF/s) and sustained "
    "(1410 TF/s) bf16 figures in MEASURED_PEAKS.json.  524 GFLOP / 0.382 ms = 1372 TF/s (0.97 of sustained).",
    "* select_kernel<1>: 48 -> 23 us.  55 MB of its DRAM reads are the fp32 re-score gathers (256 queries x ~101 rows x 2 KB): at ~8 us for that "
    "phase it runs at HBM speed, so the kernel is within ~2x of its floor.",
    "* seed level: DENSE seed scan 16.3 -> ~13 us and seed threshold 9.9 -> ~5 us under ncu (one maximum per 32 sample rows: 0.5 MB instead of "
    "17 MB written and read back); normalise 8.5 -> ~4 us (row kept in registers).",
    "* Same-box, alternating runs: round-1 build 0.4747 -> 0.4585 ms/step (hot lists), then 0.4621 -> 0.4577 serial on another box (seed maxima + "
    "normalise).",
])

# ---- configs[0] ---------------------------------------------------------------------------------------------------------
ll0, tot0 = launch_list("k1_cfg0", "scan_small", which=5)
small, _ = full_table("k1_small")
b0 = json.loads(last_json("k1_cfg0_plain"))
write("r02_cfg0_step.md", [
    "# r02_cfg0_step — configs[0]: 10k x 1024, ONE query, top-10 (the reference's own operating point, core_system.py:657-664)", "",
    "Command: `python bench.py --workload cfg0 --steps 2 --warmup 1 --no-cpu-baseline --no-north-star`.", "",
    "## Launch list of one step", "", *ll0, "",
    "Round 1: memset + normalise + scan + 2 x chunk_topk + final = 6 launches, 69 us per query resident / 156 us e2e.  Now 2 launches: the scan "
    "normalises the query itself and leaves every score (40 KB); one CTA finds the exact top-k of the dense row with its keys in registers.", "",
    "## `ncu --set full`", "", *small, "",
    "## bench.py line (plain run)", "", "```json", json.dumps({k: b0[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "roofline", "results_ok")}), "```",
])

# ---- Q = 1 on a configs[3] shard ------------------------------------------------------------------------------------------
big, brec = full_table("k1_q1big")
fb = [r for r in brec if "0, 0>" in r["Kernel Name"][0]][0]
traffic["cfg3shardq1"] = int(to_bytes(*fb["dram__bytes_read.sum"]) + to_bytes(*fb["dram__bytes_write.sum"]))
bq = json.loads(last_json("k1_q1big_plain"))
rows = list(csv.reader(open(os.path.join(G, "k1_q1big_launches.csv"))))
hdr = [r for r in rows if r and r[0] == "ID"][0]
seq = [(r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")], float(r[hdr.index("Metric Value")].replace(",", "")) / 1000.0)
       for r in rows if len(r) == len(hdr) and r[0] != "ID" and "rvo::" in r[hdr.index("Kernel Name")]]
step = seq[-4:]
write("r02_q1_big_shard.md", [
    "# r02_q1_big_shard — ONE query against a configs[3] shard (12.5M x 1280 bf16, 32 GB): the reference's Q = 1 at the metric's size", "",
    "Command: `python bench.py --workload cfg3shardq1 --steps 2 --warmup 1 --no-cpu-baseline --no-north-star`.", "",
    "## The four launches of one search (ncu launch list, us)", "", "| kernel | grid | µs |", "|---|---|---:|",
    *[f"| `{s[0][:90]}` | {s[1]} | {s[2]:.1f} |" for s in step], "",
    "Round 1 wrote 50 MB of dense scores and ran a 0.7 ms chunked top-k chain on top of the 5.4 ms scan (+13 %).  Now a strided row sample "
    "(1/96 of the shard) is scanned first, the k-th best of its per-thread maxima becomes a KEY threshold, and the full scan appends only "
    "the rows whose (score, row) key reaches it (~10k of 12.5M): scores never reach HBM.", "",
    "## `ncu --set full` of the two scans (sample pass, then the filtered full scan)", "", *big, "",
    f"DRAM traffic of the full scan: {traffic['cfg3shardq1'] / 1e9:.3f} GB for 32.000 GB algorithmic.", "",
    "## bench.py line (plain run)", "", "```json", json.dumps({k: bq[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches", "roofline", "results_ok")}), "```",
])

# ---- K1 at PE-Core-G14 width ------------------------------------------------------------------------------------------------
pool, _ = full_table("k1_pool1280")
write("r02_maskpool_1280.md", [
    "# r02_maskpool_1280 — K1 at PE-Core-G14 width (D = 1280) with 64 / 50 masks per image: region groups on the tensor path", "",
    "`python scripts/bench_maskpool.py` (plain), then `ncu --set full -k regex:mask_pool_tc`.", "",
    "Round 1 sent D = 1280 with more than 48 regions (10 slabs x 64 regions = 640 > 512 TMEM columns) to the CUDA-core kernels (4 launches, "
    "13 % of the HBM roofline, no fused ingest).  Now an image's regions are split into two groups (48 + 16), one work item each, in the same "
    "single cooperative launch; the second group re-streams the image's feature tiles (mostly from L2).", "",
    "## plain run", "", "```", *[l.strip() for l in open(os.path.join(G, "k1_pool1280_plain.log")) if l.startswith("{")], "```", "",
    "## `ncu --set full` (two launches of the 256 x 64 x 576 x 1280 shape)", "", *pool, "",
    "Reading: 387 MB of DRAM reads per launch for 377 MB of features + 9 MB of masks — the second group's pass over an image's features is served by "
    "L2 (hit rate 30 %), HBM traffic stays 1x; the cost of the split is SM time (two passes of MMAs, tensor pipe 17 % active), 0.64 of the HBM "
    "roofline vs 0.76-0.79 for D = 1024.",
])

# ---- self-join ------------------------------------------------------------------------------------------------------------------
sj, sjrec = full_table("k1_selfjoin")
write("r02_selfjoin.md", [
    "# r02_selfjoin — configs[4] building block: one 4096-row query block of the near-duplicate self-join under ncu", "",
    "`ncu --set full -k regex:\"scan_tc2|pairs_kernel\" -s 40 -c 2 python scripts/bench_selfjoin.py 400000` (a 400k-row DB keeps the capture short; "
    "the kernels and tile shapes are those of the 2M-row run).", "", *sj, "",
    "Reading: the FILTER scan of a query block keeps the tensor pipe 99.5 % active; its DRAM reads are the DB rows at or after the block "
    "(each block re-streams the tail of the DB: N/4096 blocks x ~N/2 rows x 2 KB ~ N^2/4 bytes ~ 1 TB of HBM reads for N = 2M, i.e. ~0.35 TB/s over "
    "the 2.9 s run — 5 % of the HBM peak, harmless while the kernel is tensor-bound at 0.85-0.89 of the bf16 peak).",
])
traffic["selfjoin"] = None
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in traffic.items() if not k.startswith("_")}))
