#!/bin/bash
# N GPUs (N = number visible): CLI --gpus byte-equality, contract bench
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2m${N}_gpus.txt
( time timeout 900 python -m pytest tests/test_cli.py -m gpu -x -q --timeout 800 -k "sharded" ) > gpurun_out/r2m${N}_pytest.log 2>&1
tail -6 gpurun_out/r2m${N}_pytest.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/r2m${N}_bench.json 2> gpurun_out/r2m${N}_bench.err
tail -5 gpurun_out/r2m${N}_bench.err
python - <<PY
import json
try:
    line=[l for l in open('gpurun_out/r2m${N}_bench.json').read().splitlines() if l.startswith('{')][0]
    d=json.loads(line)
    print('value',d['value'],'ms',d['ms_per_step'],'parity',d['parity_ok'])
    print('e2e',d['e2e'])
    print('q100k',json.dumps(d['query_100k'])[:2500])
    print('kernels',d['kernel_ms_per_step'])
except Exception as e:
    print('ERR',e); print(open('gpurun_out/r2m${N}_bench.err').read()[-3000:])
PY
