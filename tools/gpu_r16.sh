#!/bin/bash
# configs[2] shard shape (12.5k genomes, 10k queries): wave size and prefetch lead of the small-shard query kernel
mkdir -p gpurun_out
i=0
for v in "NQ_X=1" "NQ_QUERY_WAVE_SM=4" "NQ_QUERY_WAVE_SM=6" "NQ_QUERY_PF_AHEAD=1" "NQ_QUERY_PF_AHEAD=3" "NQ_QUERY_PF_AHEAD=0"; do
  i=$((i+1))
  env $v timeout 600 python bench.py --no-e2e --no-cpu-baseline --genomes 12500 --queries 10000 --steps 2 --warmup 1 > gpurun_out/c3s_$i.json 2> gpurun_out/c3s_$i.err
  python - gpurun_out/c3s_$i.json "$v" <<'PY'
import json, sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j["roofline_query"]
    print("c3shard", sys.argv[2], "| query ms", round(r["ms_per_launch"],3), "| frac", round(r["frac"],3), "| value", round(j["value"],1), "| hits", j["first_hits"][:3])
except Exception as e:
    print("failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
done
