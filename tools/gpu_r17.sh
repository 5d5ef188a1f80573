#!/bin/bash
mkdir -p gpurun_out
i=0
run() {  # env-string, bench args...
  v="$1"; shift; i=$((i+1))
  env $v timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 1 "$@" > gpurun_out/x_$i.json 2> gpurun_out/x_$i.err
  python - gpurun_out/x_$i.json "$v $*" <<'PY'
import json, sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j["roofline_query"]
    print(sys.argv[2], "| query ms", round(r["ms_per_launch"],3), "| frac", round(r["frac"],3), "| hits", j["first_hits"][:3])
except Exception as e:
    print("failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
}
A="--genomes 12500 --queries 10000"
run "NQ_QUERY_WAVE_SM=5" $A
run "NQ_QUERY_WAVE_SM=7" $A
run "NQ_QUERY_WAVE_SM=6 NQ_QUERY_FORM=stream" $A
run "NQ_QUERY_WAVE_SM=6 NQ_QUERY_FORM=seg16" $A
run "NQ_QUERY_WAVE_SM=6 NQ_QUERY_PF_AHEAD=1" $A
run "NQ_X=1" --genomes 10000 --queries 10000
run "NQ_QUERY_WAVE_SM=6" --genomes 10000 --queries 10000
run "NQ_QUERY_WAVE_SM=6" --genomes 10000 --queries 1000
run "NQ_QUERY_WAVE_SM=5" --genomes 10000 --queries 1000
