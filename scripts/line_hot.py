"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])   # inst, samples, long_sb, text
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iL = hdr.index("stall_long_sb"); continue
    if hdr and len(r) > iL and r[iE].isdigit():
        key = (cur_file, int(r[0]) if r[0].isdigit() else -1)
        a = agg[key]; a[0] += int(r[iE]); a[1] += int(r[iS] or 0); a[2] += int(r[iL] or 0)
        if r[1].strip(): a[3] = r[1].strip()[:110]
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print("total inst", tot, "samples", tots)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{k[0]:18s}:{k[1]:5d} inst {a[0]/tot*100:5.2f}% samp {a[1]/tots*100:5.2f}% longsb {a[2]/max(a[1],1)*100:4.0f}% | {a[3]}")
