#!/bin/bash
# Headline workload at several batch sizes (same box, back to back).
for n in 65536 131072 262144 524288; do
  echo "== shots=$n"
  timeout 300 bash scripts/bench_short.sh --workload bb144 --shots $n --steps 3 --warmup 1 --no-side --strong-shots 0
done
