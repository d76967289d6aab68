#!/bin/bash
# Occupancy / L1 experiments on the headline workload.  Output gpurun_out/<tag>_occ.txt
tag=${1:-occ}
out=gpurun_out/${tag}_occ.txt
: > $out
run() {  # label env... -- bench args
  label=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "== $label" >> $out
  env "${envs[@]}" timeout 300 bash scripts/bench_short.sh "$@" --steps 3 --warmup 1 --no-side --strong-shots 0 >> $out 2>&1
}
B="--workload bb144 --shots 131072"
run "bb144 default (8 warps, 4 CTAs/SM)" X=1 -- $B
run "bb144 8 warps, 3 CTAs/SM, carveout 75" TSR_CTAS_PER_SM=3 TSR_CARVEOUT=75 -- $B
run "bb144 8 warps, 3 CTAs/SM" TSR_CTAS_PER_SM=3 -- $B
run "bb144 10 warps" X=1 -- $B --warps 10
run "bb144 10 warps carveout 75" TSR_CARVEOUT=75 -- $B --warps 10
run "bb144 12 warps" X=1 -- $B --warps 12
run "bb144 16 warps" X=1 -- $B --warps 16
run "bb144 16 warps carveout 60" TSR_CARVEOUT=60 -- $B --warps 16
run "d11 8 warps" X=1 -- --workload surface_d11 --shots 2048
run "d11 16 warps" X=1 -- --workload surface_d11 --shots 2048 --warps 16
run "d5 default" X=1 -- --workload surface_d5 --shots 262144
run "transcx default" X=1 -- --workload transcx_d5 --shots 131072
cat $out
