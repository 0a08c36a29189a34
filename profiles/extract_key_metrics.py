"""Key counters of the `ncu --set full` captures of one round (profiles/run_profile.sh) -> <tag>_ncu_key_metrics.json.

    python profiles/extract_key_metrics.py r01h [dir]
"""
import csv
import glob
import json
import os
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]

tag = sys.argv[1]
where = sys.argv[2] if len(sys.argv) > 2 else os.path.dirname(os.path.abspath(__file__))
out = {}
for path in sorted(glob.glob(os.path.join(where, "%s_ncu_raw_*.csv" % tag))):
    name = os.path.basename(path)[len(tag) + len("_ncu_raw_"):-4]
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = ("%s %s" % (vals[i], units[i])).strip()
    out[name] = d
dst = os.path.join(where, "%s_ncu_key_metrics.json" % tag)
with open(dst, "w") as f:
    json.dump(out, f, indent=1)
print(dst, sorted(out))
