#!/bin/bash
# Final session of a round: GPU tests, the default bench (as the driver runs it), launch list, one full ncu capture.
tag=$1
# 0. one full ncu capture of the headline kernel (16384 shots, after its plain run exits 0), then profiles/traffic.json from it
SKIP=1 bash scripts/gpu_ncu.sh ${tag} c4_bb144_p1e-4_65536:16384 > gpurun_out/${tag}_ncu_capture.log 2>&1
if [ -s gpurun_out/${tag}_c4_bb144_p1e-4_65536.txt ]; then
  python scripts/update_traffic.py gpurun_out/${tag}_c4_bb144_p1e-4_65536.txt gpurun_out/${tag}_c4_bb144_p1e-4_65536.lines.txt 16384 profiles/${FINAL_NAME:-r02_z}_bb144_search_fast_bits && cp profiles/traffic.json gpurun_out/${tag}_traffic.json
fi
timeout 600 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps ${STEPS:-5} --warmup ${WARMUP:-3} > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/${tag}_bench.json; tail -12 gpurun_out/${tag}_bench.err | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --ref-seconds 30 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; tail -c 600 gpurun_out/${tag}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-side --no-cpu --strong-shots 0 > gpurun_out/${tag}_ncu_bench.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/${tag}_launches.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r]
if hdr:
    h=rows[hdr[0]]; k=h.index('Kernel Name'); v=h.index('Metric Value'); u=h.index('Metric Unit')
    tot=collections.Counter(); cnt=collections.Counter()
    for r in rows[hdr[0]+1:]:
        try:
            val=float(r[v].replace(',',''));
            if r[u]=='ns': val/=1e6
            elif r[u]=='us': val/=1e3
            elif r[u]=='s': val*=1e3
            tot[r[k][:60]]+=val; cnt[r[k][:60]]+=1
        except Exception: pass
    s=sum(tot.values())
    for n,t in tot.most_common(8): print(f'{100*t/s:6.2f}%  {t:10.2f} ms  x{cnt[n]:4d}  {n}')
PY
