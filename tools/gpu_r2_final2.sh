#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 ) > gpurun_out/r2y_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/r2y_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py ) > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2y_bench.json').read().splitlines() if l.startswith('{')][0])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'parity',d['parity_ok'],'rq',d['roofline_query']['frac'],d['roofline_query']['kernel'][:30],'traffic',d['roofline_query']['traffic'])
    print('q100k',d['query_100k']['value'],d['query_100k']['roofline']['frac'],d['query_100k']['roofline']['traffic'],d['query_100k']['parity_ok'])
    print('cpu',d['cpu_baseline']['value'],'launches',d['gpu_launches'],'clocks',d['clocks'])
except Exception as e:
    print('ERR',e); print(open('gpurun_out/r2y_bench.err').read()[-2000:])
PY
