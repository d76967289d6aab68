#!/bin/bash
# ncu captures of the search kernel on named cases: full metric set with source (one launch) + launch list.
# usage: scripts/gpu_ncu.sh <tag> <case>[:shots] ...      outputs gpurun_out/<tag>_<case>.{ncu-rep,txt,lines.txt}
tag=$1; shift
for spec in "$@"; do
  case_=${spec%%:*}; shots=${spec#*:}; [ "$shots" = "$spec" ] && shots=""
  python scripts/one_decode.py $case_ $shots > gpurun_out/${tag}_${case_}.plain.log 2>&1 || { echo "plain run failed: $case_"; tail -3 gpurun_out/${tag}_${case_}.plain.log; continue; }
  cat gpurun_out/${tag}_${case_}.plain.log
  # skip the warm-up decode's search launches (tier 0 + resume where searches get paused: SKIP=1 where none are), capture the real tier-0 launch
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:search_fast_kernel --launch-skip ${SKIP:-2} --launch-count 1 \
      -f -o gpurun_out/${tag}_${case_} python scripts/one_decode.py $case_ $shots > gpurun_out/${tag}_${case_}.ncu.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${tag}_${case_}.ncu-rep > gpurun_out/${tag}_${case_}.txt 2>&1
  python scripts/ncu_lines.py gpurun_out/${tag}_${case_}.ncu-rep 70 > gpurun_out/${tag}_${case_}.lines.txt 2>&1
  head -30 gpurun_out/${tag}_${case_}.txt | cut -c1-140
done
