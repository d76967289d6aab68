"""Rewrites profiles/traffic.json from one `ncu --set full` capture of the headline workload's search kernel.
usage: python scripts/update_traffic.py <summary.txt> <lines.txt> <shots in the captured launch> <name the files are committed under, without suffix>
(the summary / lines files are what scripts/gpu_ncu.sh writes with scripts/ncu_summary.py and scripts/ncu_lines.py)"""
import json
import os
import re
import sys

summary, lines, shots, committed = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
vals = {}
for line in open(summary):
    m = re.match(r"(\S+)\s+([\d.]+)\s*(\S*)", line.strip())
    if m:
        vals[m[1]] = (float(m[2]), m[3])


def nbytes(key):
    v, u = vals[key]
    return v * UNIT[u]


instr = float(re.search(r"total warp instructions ([\d.e+]+)", open(lines).read())[1])
rec = {
    "dram_bytes_per_shot": (nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")) / shots,
    "warp_instructions_per_shot": instr / shots,
    "l2_hit_pct": round(vals["lts__t_sector_hit_rate.pct"][0], 2),
    "l2_throughput_pct": round(vals["lts__throughput.avg.pct_of_peak_sustained_elapsed"][0], 2),
    "achieved_occupancy_pct": round(vals["sm__warps_active.avg.pct_of_peak_sustained_active"][0], 1),
    "dram_pct_of_peak": round(vals["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0], 2),
    "issue_active_pct": round(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"][0], 1),
    "source": committed + "_ncu.txt",
    "capture": committed + "_ncu.txt",
    "shots_in_capture": shots,
}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {
    "_comment": "Per-shot figures of the dominant kernel (tsr::search_fast_kernel, bit-sliced scan) from one `ncu --set full` "
    f"capture of a {shots}-shot launch of the headline workload; bench.py scales them to its launches for roofline.traffic and "
    f"the `issue` record. Source: {committed}_ncu.txt, {committed}_lines.txt (written by scripts/update_traffic.py).",
    "bb144": rec,
}
with open(os.path.join(root, "profiles", "traffic.json"), "w") as f:
    json.dump(out, f, indent=1)
    f.write("\n")
print(json.dumps(rec))
