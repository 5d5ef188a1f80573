#!/bin/bash
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
( time timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N ) > gpurun_out/r2m${N}_bench.json 2> gpurun_out/r2m${N}_bench.err
tail -4 gpurun_out/r2m${N}_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2m${N}_bench.json').read().splitlines() if l.startswith('{')][0])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['genomes'],'parity',d['parity_ok'])
    q=d['query_100k']; print('q100k',q['value'],q['ms_per_step'],q['count_kernel_ms_per_step'],q['roofline']['frac'],q['parity_ok'],q['auto_layout'])
except Exception as e:
    print('ERR',e); print(open('gpurun_out/r2m${N}_bench.err').read()[-2500:])
PY
