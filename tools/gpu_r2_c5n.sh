#!/bin/bash
# N GPUs: configs[4] (--matrix grid tiled over the sharded index); with N >= 4 also the contract line
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload c5 --steps 2 ) > gpurun_out/r2c5_${N}gpu.json 2> gpurun_out/r2c5_${N}gpu.err
tail -3 gpurun_out/r2c5_${N}gpu.err
if [ "$N" -ge 4 ] && [ -z "$SKIP_CONTRACT" ]; then
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 3 --warmup 3 ) > gpurun_out/r2m${N}_bench.json 2> gpurun_out/r2m${N}_bench.err
tail -3 gpurun_out/r2m${N}_bench.err
fi
python - <<PY
import json,os
for f in ['r2c5_${N}gpu','r2m${N}_bench']:
    if not os.path.exists(f'gpurun_out/{f}.json'): continue
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json').read().splitlines() if l.startswith('{')][0])
        print(f,'value',d['value'],d['unit'],'ms',d['ms_per_step'],'e2e',(d.get('e2e') or {}).get('value'),'parity',d.get('parity_ok'),d.get('diag_is_F_mod_65536'),d.get('matrix_device_ms_per_step'))
        if d.get('query_100k'): print('  q100k',json.dumps(d['query_100k'])[:1800])
    except Exception as e:
        print(f,'ERR',e); print(open(f'gpurun_out/{f}.err').read()[-2000:])
PY
