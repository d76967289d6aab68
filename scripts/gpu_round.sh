#!/bin/bash
# One GPU-box session: parity tests, phase profiles, the default bench.  Outputs under gpurun_out/<tag>_*.
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
timeout ${PYTEST_TIMEOUT:-900} python -m pytest tests -m gpu -q --durations=15 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
if [ -n "$PHASES" ]; then
  timeout ${PHASES_TIMEOUT:-600} python scripts/profile_phases.py $PHASES > gpurun_out/${tag}_phases.jsonl 2> gpurun_out/${tag}_phases.err
  cat gpurun_out/${tag}_phases.jsonl | cut -c1-600
fi
if [ -z "$NO_BENCH" ]; then
  timeout ${BENCH_TIMEOUT:-600} python bench.py --steps ${BENCH_STEPS:-3} --warmup ${BENCH_WARMUP:-1} ${BENCH_ARGS} > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
  echo "bench rc=$?"
  tail -c 3000 gpurun_out/${tag}_bench.json; tail -20 gpurun_out/${tag}_bench.err
fi
