"""Per-source-line instruction and stall-sample shares from an .ncu-rep captured with --import-source on.
usage: python scripts/ncu_lines.py prof.ncu-rep [top_n]"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
data = []
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        try:
            data.append((int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), int(r[0]), r[1],
                         int(r[hdr.index("L1 Wavefronts Shared")] or 0), int(r[hdr.index("L1 Wavefronts Shared Ideal")] or 0),
                         int(r[hdr.index("stall_long_sb")] or 0), int(r[hdr.index("stall_short_sb")] or 0), int(r[hdr.index("stall_barrier")] or 0)))
        except ValueError:
            pass
ti = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print(f"total warp instructions {ti:.3e}, samples {ts}")
print(" inst%  samp%  line  smem_wavefronts(ideal)  long_sb short_sb barrier  source")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100*d[0]/ti:5.1f}  {100*d[1]/ts:5.1f}  {d[2]:4d}  {d[4]:.2e}({d[5]:.2e})  {d[6]:7d} {d[7]:7d} {d[8]:7d}  {d[3].strip()[:100]}")
