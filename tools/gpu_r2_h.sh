#!/bin/bash
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 100 -k "packed" ) > gpurun_out/r2h_pytest.log 2>&1
tail -40 gpurun_out/r2h_pytest.log
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[0])
    print('value',d['value'],'ms',d['ms_per_step'],'parity',d['parity_ok'])
    print('e2e',d['e2e'])
    print('q100k',json.dumps(d['query_100k'])[:1500])
    print('rq',d['roofline_query']['frac'], d['kernel_ms_per_step'])
    print('cpu',d.get('cpu_baseline'))
except Exception as e:
    print('ERR',e); print(open('gpurun_out/r2h_bench.err').read()[-3000:])
PY
./tools/microbench | grep -i int32
