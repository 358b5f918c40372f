"""Per-kernel totals and shares of an ncu launch list (the `--metrics gpu__time_duration.sum --csv` pass of
tools/ncu_profile.sh):  python tools/launch_summary.py profiles/<tag>_launches.csv [bench.json] > profiles/<tag>_launches_summary.md"""
import csv
import json
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik].replace("lsim::", "").replace("void ", "")).strip()
    v = float(r[iv].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r[iu]]
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += v
step = {k: v for k, v in tot.items() if k.startswith("k_") and k != "k_fp64_peak"}
s = sum(v[1] for v in step.values())
tag = path.split("/")[-1].split("_")[0]
print("# ncu launch list `%s`: `python bench.py --steps 1 --warmup 3 --batch 9472 --no-cpu --no-secondary`\n" % path)
print("Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv` via `tools/ncu_profile.sh %s` (cold-cache, "
      "serialised: compare SHARES).  Shares exclude the `k_fp64_peak` roofline microbenchmark and torch's own element-wise "
      "kernels (bound checks of the sampler boundary, < 0.1 %%).\n" % tag)
print("| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    if k.startswith("k_"):
        print("| `%s` | %d | %.3f | %s |" % (k, n, ms, "-" if k == "k_fp64_peak" else "%.1f %%" % (100 * ms / s)))
if len(sys.argv) > 2:
    d = json.load(open(sys.argv[2]))
    print("\nLive CUDA-event shares of the same build (`%s`, %d particles, %.0f images/s on that box): %s." % (
        sys.argv[2], d["config"].get("particles", 65536), d["value"],
        " / ".join("%s %.3f" % kv for kv in d["roofline"]["kernel_share_of_step"].items())))
