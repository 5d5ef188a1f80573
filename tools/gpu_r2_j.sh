#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 --durations=8 ) > gpurun_out/r2j_pytest.log 2>&1
tail -16 gpurun_out/r2j_pytest.log
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err
tail -3 gpurun_out/r2j_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[0])
    print('value',d['value'],'ms',d['ms_per_step'],'parity',d['parity_ok'])
    print('e2e',d['e2e'])
    print('q100k',d['query_100k']['value'], d['query_100k']['roofline']['frac'], d['query_100k']['parity_ok'])
    print('rq',d['roofline_query']['frac'], d['kernel_ms_per_step'], d['roofline']['frac'])
except Exception as e:
    print('ERR',e); print(open('gpurun_out/r2j_bench.err').read()[-3000:])
PY
