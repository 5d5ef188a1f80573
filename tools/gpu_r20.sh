#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? : $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head -5
NQ_QUERY_NT=256 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "query or golden or c1 or lines or sharded or device_pointer or matrix or synthetic" > gpurun_out/pytest_nt256.log 2>&1
echo "nt256 pytest exit $? : $(tail -1 gpurun_out/pytest_nt256.log)"
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
run "NQ_X=1" --genomes 10000 --queries 1000
run "NQ_X=1" --genomes 10000 --queries 2000
