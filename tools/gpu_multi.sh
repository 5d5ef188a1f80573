#!/bin/bash
# N GPUs of one box (N = $1): the driver's launch line for the contract bench, then the 100k-genome query workload
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc $?"; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --workload q100k --steps 3 --warmup 2 > gpurun_out/q100k_n$N.json 2> gpurun_out/q100k_n$N.err
echo "q100k rc $?"; tail -3 gpurun_out/q100k_n$N.err | cut -c1-300
python - $N <<'PY'
import json, sys
N=sys.argv[1]
for f in (f"gpurun_out/bench_n{N}.json", f"gpurun_out/q100k_n{N}.json"):
    try:
        j=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        r=j.get("roofline_query") or j["roofline"]
        print(f, "| value", round(j["value"],1), j["unit"], "| ms", round(j["ms_per_step"],2), "| e2e", round(j["e2e"]["value"],1), "| query ms", round(r["ms_per_launch"],3), "frac", round(r["frac"],3), "| hits", j.get("first_hits"))
    except Exception as e:
        print(f, "failed", e)
PY
