"""Per-instruction view of an `ncu --page source --csv` dump (gzip ok): executed-instruction mix, the instructions that
collect the most stall samples, and local-memory traffic.   python tools/ncu_source_hot.py k_image_x_source.csv.gz [top]"""
import collections
import csv
import gzip
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hdr = rows[1]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") or h.startswith("Stall")]
body = rows[2:]
tot = sum(int(r[ie]) for r in body)
ts = sum(int(r[isamp]) for r in body)
mx = max(int(r[ie]) for r in body)
print("total warp-inst %d, samples %d" % (tot, ts))
# hot region = instructions executed at least half as often as the most executed one (the sub-pixel loop)
hot = [r for r in body if int(r[ie]) >= 0.25 * mx]
print("hot-loop instructions: %d static, %.1f%% of executed, %.1f%% of samples" % (
    len(hot), 100.0 * sum(int(r[ie]) for r in hot) / tot, 100.0 * sum(int(r[isamp]) for r in hot) / ts))
ops = collections.Counter()
samp = collections.Counter()
for r in body:
    t = r[ia].split()
    o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[o] += int(r[ie])
    samp[o] += int(r[isamp])
fp64 = sum(ops[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
print("fp64-pipe warp-inst %.1f%% of all" % (100.0 * fp64 / tot))
for o, c in ops.most_common(top):
    print("  %-8s %6.2f%% inst  %6.2f%% samples" % (o, 100.0 * c / tot, 100.0 * samp[o] / ts))
print("top instructions by samples:")
for r in sorted(body, key=lambda r: -int(r[isamp]))[:top]:
    print("  %6.2f%%  x%-9s %s" % (100.0 * int(r[isamp]) / ts, r[ie], r[ia].strip()[:90]))
