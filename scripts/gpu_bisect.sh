#!/bin/bash
# Bench the headline workload with the library built at several commits (under _bisect/<commit>/).
for d in _bisect/*/; do
  c=$(basename $d)
  r=$(cd $d && timeout 200 python bench.py --no-cpu --no-side --strong-shots 0 --shots 131072 --steps 3 --warmup 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f frac %.3f kernel_ms/step %.1f' % (d['value'], d['roofline']['frac'], d['search_kernel_ms_per_step']))
")
  echo "$c: $r"
done
