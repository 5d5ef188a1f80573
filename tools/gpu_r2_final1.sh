#!/bin/bash
# round 2, one GPU: full parity suite, contract line, configs[4] tiles, ncu of the granule form at a 50k shard
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 ) > gpurun_out/r2z_pytest.log 2>&1
tail -4 gpurun_out/r2z_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
tail -2 gpurun_out/r2z_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err
( time timeout 600 python bench.py --workload c5 --steps 2 ) > gpurun_out/r2z_c5_1gpu.json 2> gpurun_out/r2z_c5_1gpu.err
tail -2 gpurun_out/r2z_c5_1gpu.err
timeout 300 ncu --set full --clock-control none -k regex:"query_slab|slab_resolve" -c 2 -o gpurun_out/r2z_slab50k \
  python bench.py --workload q100k --genomes 50000 --queries 2000 --steps 1 --warmup 0 > gpurun_out/r2z_ncu.log 2>&1
python - <<'PY'
import json
for f in ['r2z_bench','r2z_bench_reference','r2z_c5_1gpu']:
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json').read().splitlines() if l.startswith('{')][0])
        print(f, 'value',d['value'],d['unit'],'ms',d['ms_per_step'], 'e2e',(d.get('e2e') or {}).get('value'), 'parity',d.get('parity_ok'), 'rq',(d.get('roofline_query') or {}).get('frac'), 'r',(d.get('roofline') or {}).get('frac'))
        if d.get('query_100k'): print('  q100k', d['query_100k']['value'], d['query_100k']['roofline']['frac'])
    except Exception as e:
        print(f,'ERR',e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
