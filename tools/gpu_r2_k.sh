#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 3 --warmup 3 --no-q100k --no-cpu-baseline ) > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
tail -3 gpurun_out/r2k_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2k_bench.json').read().strip().splitlines()[0])
    print('value',d['value'],'ms',d['ms_per_step'],'parity',d['parity_ok'])
    print('e2e',d['e2e'])
except Exception as e:
    print('ERR',e); print(open('gpurun_out/r2k_bench.err').read()[-3000:])
PY
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "packed or sketch" ) > gpurun_out/r2k_pytest.log 2>&1
tail -4 gpurun_out/r2k_pytest.log
