"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line (run on the dev box).
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python scripts/ncu_lines.py src.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
rows = list(csv.reader(open(path)))
agg = collections.OrderedDict(); cur_file = None; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[2] != "-": continue                      # SASS rows carry an address; source rows have '-'
    g = lambda n: float(r[hdr.index(n)] or 0)
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, dict(src=r[1].strip(), samples=0, inst=0, long_sb=0, short_sb=0, wait=0, barrier=0, math=0, notsel=0, sleep=0))
    a["samples"] += g("# Samples"); a["inst"] += g("Instructions Executed"); a["long_sb"] += g("stall_long_sb"); a["short_sb"] += g("stall_short_sb")
    a["wait"] += g("stall_wait"); a["barrier"] += g("stall_barrier"); a["math"] += g("stall_math"); a["notsel"] += g("stall_not_selected"); a["sleep"] += g("stall_sleep")
ts = sum(a["samples"] for a in agg.values()); ti = sum(a["inst"] for a in agg.values())
print(f"# total samples {ts:.0f}, total warp instructions {ti/1e6:.0f}M")
print("# file:line  samples%  instr%  instr(M)  long_sb short_sb wait barrier math notsel sleep | source")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    print(f"{f:16s}:{ln:4d} {100*a['samples']/ts:6.1f}% {100*a['inst']/ti:6.1f}% {a['inst']/1e6:8.1f} {a['long_sb']:7.0f} {a['short_sb']:6.0f} {a['wait']:6.0f} {a['barrier']:6.0f} {a['math']:6.0f} {a['notsel']:6.0f} {a['sleep']:6.0f} | {a['src'][:110]}")
