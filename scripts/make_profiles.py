#!/usr/bin/env python
"""Turns the raw ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (launch list + per-kernel metrics)."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
go = os.path.join(ROOT, "gpurun_out"); out = os.path.join(ROOT, "profiles"); os.makedirs(out, exist_ok=True)

# ---- launch list -----------------------------------------------------------------------------------------------------
lines = [l for l in open(os.path.join(go, f"{tag}_launches.csv")) if not l.startswith("==")]
agg = collections.OrderedDict(); seq = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; seq.append((name, v))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list of `python bench.py --steps 1 --warmup 1 --no-cpu-baseline`\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised launches: compare SHARES, not absolutes).\n"
            f"{len(seq)} launches captured, {tot / 1e3:.2f} ms total.\n\n| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.2f} | {v[1] / tot * 100:.1f} % | {v[1] / v[0]:.1f} |\n")
    f.write("\nOne wave (8 sample passes of 1920x1080 = 16.6 M paths), launch order:\n\n| # | kernel | us |\n|---|---|---|\n")
    start = next(i for i, (n, _) in enumerate(seq) if n == "k_camera")
    start = next(i for i, (n, _) in enumerate(seq) if n == "k_camera" and i > start)   # second wave: warm
    for i, (n, v) in enumerate(seq[start:start + 22]):
        f.write(f"| {i} | `{n}` | {v:.1f} |\n")
        if n == "k_accumulate":
            break

# ---- full capture ----------------------------------------------------------------------------------------------------
rep = os.path.join(go, f"{tag}_full.ncu-rep")
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / inst"), ("smsp__inst_executed.sum", "warp insts"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__t_bytes.sum", "L1 bytes")]
with open(os.path.join(out, f"{tag}_kernels.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of the hot kernels (second wave of the bench step)\n\n")
    f.write("| kernel | " + " | ".join(n for _, n in want) + " |\n|---|" + "---|" * len(want) + "\n")
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for m, _ in want:
            if m in idx:
                cells.append(f"{d[idx[m]]} {units[idx[m]]}".strip())
            else:
                cells.append("-")
        f.write(f"| `{name}` | " + " | ".join(cells) + " |\n")
    f.write("\nReading: the traversal kernels (`k_trace`, `k_shadow`) are instruction-issue bound (issue active 70-80 %) with DRAM at a few % of peak:\n"
            "the 15 MB scene is L1/L2 resident (SURVEY H6) and the exact (un-fused) slab/triangle arithmetic costs ~76 instructions per internal-node step.\n"
            "Their SIMD efficiency falls from ~28 active lanes per instruction on primary rays to ~12 on bounce and shadow rays.\n")
print(open(os.path.join(out, f"{tag}_launches.md")).read())
print(open(os.path.join(out, f"{tag}_kernels.md")).read())
