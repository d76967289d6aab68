#!/bin/bash
# usage: scripts/bench_short.sh <bench args...>   -> one summary line
python bench.py --no-cpu "$@" 2>gpurun_out/bench_short.err | python -c "
import sys,json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line); r=d['roofline']
    print('value %.0f e2e %.0f frac %.3f path %s kernel_ms/step %.1f  %s' % (d['value'], d['e2e']['value'], r['frac'], r['search_path'], d['search_kernel_ms_per_step'], d['config']['workload'][:40]))
" || tail -5 gpurun_out/bench_short.err
