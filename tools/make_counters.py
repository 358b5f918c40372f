#!/usr/bin/env python
"""profiles/<kernel>_<tag>_ncu_summary.txt (tools/ncu_summary.py output of one `ncu --set full` capture of k_image) ->
profiles/k_image_counters.json, the per-sub-pixel instruction counts and per-set DRAM bytes bench.py's `roofline`
block uses.

    python tools/make_counters.py profiles/image_r2a_ncu_summary.txt [sub-pixels per set = 291600] [sets in the launch]
"""

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = sys.argv[1]
    sub_per_set = int(sys.argv[2]) if len(sys.argv) > 2 else 180 * 180 * 9
    m = {}
    for line in open(path):
        parts = line.split()
        if len(parts) >= 2 and ("__" in parts[0]):
            try:
                m[parts[0]] = float(parts[1])
            except ValueError:
                pass
    cycles = m["sm__cycles_elapsed.max"]
    # sets in the captured launch: tools/ncu_profile.sh profiles one full chunk (--batch 9472 = 64 x 148 SMs); the pixel
    # tiling per set varies with the chunk, so the count is an argument, checked against the image bytes written
    sets = int(sys.argv[3]) if len(sys.argv) > 3 else 9472
    sub = sets * sub_per_set
    per = lambda k: m["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % k] * cycles / sub  # noqa: E731
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}

    def dram(key):
        for line in open(path):
            if line.startswith(key):
                parts = line.split()
                return float(parts[1]) * unit[parts[2]]
        return None

    w = dram("dram__bytes_write.sum")
    assert w is None or 0.8 < w / (sets * 180 * 180 * 8.0) < 1.2, "set count does not match the bytes written"
    out = {
        "capture": os.path.relpath(path, ROOT),
        "kernel": "k_image_fold<1, 3, 1> (headline instantiation)", "sets_in_launch": sets, "subpixels_per_set": sub_per_set,
        "gpu_time_ms": m.get("gpu__time_duration.sum"),
        "dfma_per_subpixel": per("dfma"), "dmul_per_subpixel": per("dmul"), "dadd_per_subpixel": per("dadd"),
        "pipe_fp64_cycles_active_pct": m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": m.get("launch__registers_per_thread"),
        "local_loads": m.get("sass__inst_executed_local_loads"), "local_stores": m.get("sass__inst_executed_local_stores"),
        "dram_read_bytes_per_set": dram("dram__bytes_read.sum") / sets,
        "dram_write_bytes_per_set": dram("dram__bytes_write.sum") / sets,
    }
    with open(os.path.join(ROOT, "profiles", "k_image_counters.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
