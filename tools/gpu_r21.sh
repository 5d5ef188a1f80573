#!/bin/bash
mkdir -p gpurun_out
i=0
run() {
  v="$1"; shift; i=$((i+1))
  env $v timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 1 "$@" > gpurun_out/x_$i.json 2> gpurun_out/x_$i.err
  python - gpurun_out/x_$i.json "$v $*" <<'PY'
import json, sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j["roofline_query"]
    print(sys.argv[2], "| query ms", round(r["ms_per_launch"],3), "| frac", round(r["frac"],3), "| value", round(j["value"],1), "| hits", j["first_hits"][:3])
except Exception as e:
    print("failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-600:])
PY
}
run "NQ_X=1" --genomes 12500 --queries 10000
run "NQ_X=1" --genomes 12500 --queries 1250
run "NQ_QUERY_NT=128" --genomes 12500 --queries 1250
run "NQ_X=1" --genomes 10000 --queries 1000
run "NQ_X=1" --genomes 10000 --queries 2000
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 280 -k "golden or c1 or sharded or full_size" > gpurun_out/pytest_q.log 2>&1; echo "pytest: $(tail -1 gpurun_out/pytest_q.log)"
