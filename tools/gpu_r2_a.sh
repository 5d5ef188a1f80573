#!/bin/bash
# round 2, first contact of the slab query form: parity suite, contract bench, query-only shapes
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err
for shape in "10000 1000" "12500 10000" "10000 10000"; do
  set -- $shape
  timeout 600 python bench.py --workload q100k --genomes $1 --queries $2 --steps 5 --warmup 3 > gpurun_out/r2a_q_$1_$2.json 2> gpurun_out/r2a_q_$1_$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline_query') or d.get('roofline')
        print(f, 'value',d['value'],'ms',d['ms_per_step'],'q frac',r.get('frac'),'ms/launch',r.get('ms_per_launch'), d.get('kernel_ms_per_step'))
    except Exception as e:
        print(f,'ERR',e)
PY
