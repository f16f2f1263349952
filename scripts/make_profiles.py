#!/usr/bin/env python
"""Turns the raw ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/make_profiles.py TAG [--step-kernels N]

inputs  gpurun_out/TAG_launches.csv   ncu --metrics gpu__time_duration.sum --clock-control none --csv of `bench.py --quick --steps 2 --warmup 3`
        gpurun_out/TAG_full.ncu-rep   ncu --set full --clock-control none --import-source on of the kernels of ONE step (one wave of 32 passes)
outputs profiles/TAG_launches.md      kernel shares of a step
        profiles/TAG_kernels.md       per-launch metrics (time, registers, issue, active lanes, cache hit rates, DRAM, L2 bytes)
        profiles/TAG_capture.json     the figures bench.py quotes in `roofline` (DRAM bytes / thread instructions of the closest-hit launches per step, lane-issue
                                      fractions, L2 GB/s), with the commit they were captured at
        profiles/TAG_sass_inner_loop.txt   SASS of the closest-hit inner-node loop with executed-instruction counts (ncu source page)
"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
go = os.path.join(ROOT, "gpurun_out"); out = os.path.join(ROOT, "profiles"); os.makedirs(out, exist_ok=True)
commit = subprocess.check_output(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], text=True).strip()
SMS, SCHED, LANES = 148, 4, 32


def short(name):
    return name.split("(")[0].replace("void ", "")


# ---- launch list -----------------------------------------------------------------------------------------------------
lp = os.path.join(go, f"{tag}_launches.csv")
if os.path.exists(lp):
    lines = [l for l in open(lp) if not l.startswith("==")]
    agg = collections.OrderedDict(); seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = short(row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; seq.append((name, v))
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of `python bench.py --quick --steps 2 --warmup 3` (commit {commit})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv` (cold-cache, serialised launches: compare SHARES, not absolutes).\n"
                f"{len(seq)} launches captured, {tot / 1e3:.2f} ms total.\n\n| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.2f} | {v[1] / tot * 100:.1f} % | {v[1] / v[0]:.1f} |\n")
        f.write("\nOne step = one wave (32 sample passes of 1920x1080 = 66.4 M paths), launch order of the last captured wave:\n\n| # | kernel | us |\n|---|---|---|\n")
        starts = [i for i, (n, _) in enumerate(seq) if n == "k_trace_primary"]
        if starts:
            for i, (n, v) in enumerate(seq[starts[-1]:]):
                f.write(f"| {i} | `{n}` | {v:.1f} |\n")
                if n == "k_accumulate":
                    break
    print(open(os.path.join(out, f"{tag}_launches.md")).read())

# ---- full capture ----------------------------------------------------------------------------------------------------
rep = os.path.join(go, f"{tag}_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(d, m, default=0.0):
        if m not in idx or d[idx[m]] in ("", "n/a"):
            return default
        v = float(d[idx[m]].replace(",", "")); u = units[idx[m]]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0,
                 "Ghz": 1e9, "Mhz": 1e6, "cycle/nsecond": 1e9, "cycle/usecond": 1e6}
        return v * scale.get(u, 1.0)

    want = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
            ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / inst"), ("smsp__inst_executed.sum", "warp insts"),
            ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
            ("lts__t_sectors.sum", "L2 sectors (32 B)")]
    cls = collections.defaultdict(lambda: collections.Counter())
    with open(os.path.join(out, f"{tag}_kernels.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of the kernels of one step (one wave of 32 passes, hyperion_rect_lights 1080p), commit {commit}\n\n")
        f.write("| kernel | " + " | ".join(n for _, n in want) + " | lane-issue frac |\n|---|" + "---|" * (len(want) + 1) + "\n")
        for d in data:
            name = short(d[idx["Kernel Name"]])
            t = val(d, "gpu__time_duration.sum"); clk = val(d, "gpc__cycles_elapsed.max.per_second", 1.965e9) or 1.965e9
            wi = val(d, "smsp__inst_executed.sum"); ti = wi * val(d, "smsp__thread_inst_executed_per_inst_executed.ratio")
            frac = ti / (SMS * SCHED * LANES * clk * t) if t > 0 else 0.0
            cells = [f"{d[idx[m]]} {units[idx[m]]}".strip() if m in idx else "-" for m, _ in want]
            f.write(f"| `{name}` | " + " | ".join(cells) + f" | {frac:.3f} |\n")
            c = cls["trace" if name.startswith("k_trace") else name]
            c["time"] += t; c["thread_inst"] += ti; c["warp_inst"] += wi; c["dram"] += val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum")
            c["l2_bytes"] += 32.0 * val(d, "lts__t_sectors.sum"); c["clk_t"] += clk * t
        f.write("\nlane-issue frac = thread instructions executed / (148 SMs x 4 schedulers x 32 lanes x SM clock x duration): the fraction of the SIMD issue capacity the launch uses.\n"
                "The traversal kernels are bound by instruction issue x SIMD width (issue-active 70-83 %, a third of the lanes active on incoherent rays); the 15 MB scene is L1/L2 resident,\n"
                "DRAM stays below 12 % of peak for them.\n")
    tr, sh = cls["trace"], cls["k_shadow"]
    cap = {"workload": "hyperion_rect_lights", "commit": commit, "source": f"gpurun_out/{tag}_full.ncu-rep (ncu --set full --clock-control none, one wave = one 32-spp step)",
           "k_trace_dram_bytes_per_step": tr["dram"], "k_trace_thread_inst_per_step": tr["thread_inst"], "k_trace_warp_inst_per_step": tr["warp_inst"],
           "k_trace_ms_in_capture": tr["time"] * 1e3, "k_trace_lane_issue_frac": tr["thread_inst"] / (SMS * SCHED * LANES * tr["clk_t"]) if tr["clk_t"] else None,
           "k_shadow_lane_issue_frac": sh["thread_inst"] / (SMS * SCHED * LANES * sh["clk_t"]) if sh["clk_t"] else None,
           "k_shadow_thread_inst_per_step": sh["thread_inst"], "k_shadow_ms_in_capture": sh["time"] * 1e3,
           "k_trace_l2_gbs": tr["l2_bytes"] / tr["time"] / 1e9 if tr["time"] else None, "k_shadow_l2_gbs": sh["l2_bytes"] / sh["time"] / 1e9 if sh["time"] else None,
           "note": "k_trace = k_trace_primary + the k_trace launches of the step (the launches bench.py times as trace_ms_per_step)"}
    with open(os.path.join(out, f"{tag}_capture.json"), "w") as f:
        json.dump(cap, f, indent=1)
    print(open(os.path.join(out, f"{tag}_kernels.md")).read()); print(json.dumps(cap, indent=1))

    # ---- SASS of the inner-node loop of the bounce-1 k_trace launch, with executed counts ----
    names = [short(d[idx["Kernel Name"]]) for d in data]
    kid = next((i for i, n in enumerate(names) if n == "k_trace"), None)
    if kid is not None:
        src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid + 1}"], text=True, stderr=subprocess.DEVNULL)
        srows = list(csv.reader(src.splitlines()))
        h = next(r for r in srows if "Instructions Executed" in r)
        iS, iE, iT, iP = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
        sd = [r for r in srows if len(r) > iT and r[iE].isdigit()]
        sd = sd[:len(sd) // 2] if len(sd) > 2 and sd[0][iS] == sd[len(sd) // 2][iS] else sd
        ex = [int(r[iE]) for r in sd]
        peak = max(ex); hot = [i for i, e in enumerate(ex) if e > 0.8 * peak]
        a, b = max(0, hot[0] - 4), min(len(sd), hot[-1] + 8)
        with open(os.path.join(out, f"{tag}_sass_inner_loop.txt"), "w") as f:
            f.write(f"# {tag}: SASS of the inner-node loop of k_trace (bounce-1 launch of the captured wave, commit {commit}); columns: warp-level executions, thread-level executions,\n"
                    f"# stall samples.  Source: ncu --page source of gpurun_out/{tag}_full.ncu-rep.  The loop body is the range executed ~{peak / 1e6:.1f} M times.\n"
                    f"# One iteration = one internal node: 4 x LDG (64-byte node), 12 FADD + 12 FMUL + 16 FMNMX/FMNMX3 (two IEEE slab tests, un-fusable), hit / cull predicates,\n"
                    f"# near-first ordering, shared-memory push / pop.\n")
            for i in range(a, b):
                f.write(f"{ex[i]:>10d} {int(sd[i][iT]):>12d} {int(sd[i][iP]):>6d}  {sd[i][iS].strip()}\n")
            tot_w = sum(ex); loop_w = sum(ex[i] for i in hot)
            f.write(f"# loop body: {len(hot)} SASS instructions, {loop_w / tot_w * 100:.1f} % of the launch's warp instructions, "
                    f"{sum(int(sd[i][iT]) for i in hot) / max(loop_w, 1):.1f} active lanes per instruction\n")
        print(open(os.path.join(out, f"{tag}_sass_inner_loop.txt")).read()[-1500:])
