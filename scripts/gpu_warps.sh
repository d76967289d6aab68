#!/bin/bash
# Warps-per-search sweep through bench.py (value / e2e / frac per setting).  Output gpurun_out/<tag>_warps.txt
tag=${1:-warps}
out=gpurun_out/${tag}_warps.txt
: > $out
run() {  # workload shots warps
  echo "== $1 shots=$2 warps=$3" >> $out
  timeout 300 bash scripts/bench_short.sh --workload $1 --shots $2 --warps $3 --steps 2 --warmup 1 --no-side --strong-shots 0 >> $out 2>&1
}
for w in 1 2 4; do run surface_d5 262144 $w; done
for w in 1 2 4; do run transcx_d5 131072 $w; done
for w in 2 4 8; do run surface_d11 2048 $w; done
for w in 4 8; do run color_d7 600 $w; done
for w in 4 6 8; do run bb144 65536 $w; done
cat $out
