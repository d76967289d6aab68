#!/bin/bash
# The side workloads' headline numbers through bench.py (short).
for spec in "surface_d5 262144" "transcx_d5 131072" "surface_d11 2048" "color_d7 600"; do
  set -- $spec
  echo "== $1 shots=$2"
  timeout 300 bash scripts/bench_short.sh --workload $1 --shots $2 --steps 2 --warmup 1 --no-side --strong-shots 0
done
