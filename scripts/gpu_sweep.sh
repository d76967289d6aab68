#!/bin/bash
# Tail-widening sweep: expansion budget x wide warps on the heavy-tailed cases.  Output: gpurun_out/<tag>_sweep.jsonl
tag=${1:-sweep}
: > gpurun_out/${tag}_sweep.jsonl
for cfg in "0 0" "64 32" "256 32" "1024 32" "256 16"; do
  set -- $cfg
  echo "{\"budget\": $1, \"wide\": $2}" >> gpurun_out/${tag}_sweep.jsonl
  TSR_EXPANSION_BUDGET=$1 TSR_WIDE_WARPS=$2 timeout 300 python scripts/profile_phases.py ${SWEEP_CASES:-c3_surface_d11_2048 c2_color_d7_first600 bb144_p1e-3_longbeam_256} >> gpurun_out/${tag}_sweep.jsonl 2>> gpurun_out/${tag}_sweep.err
done
cut -c1-420 gpurun_out/${tag}_sweep.jsonl
