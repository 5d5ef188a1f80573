#!/bin/bash
# Quick iteration: parity tests of the kernels, then full-size bench variants selected by env (VARIANTS="A=1 B=2;C=3")
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
i=0
IFS=';' read -ra VARS <<< "${VARIANTS:-NQ_NONE=1}"
for v in "${VARS[@]}"; do
  i=$((i+1))
  env $v timeout 600 python bench.py ${BENCH_ARGS:---no-e2e --no-cpu-baseline} > gpurun_out/bench_v$i.json 2> gpurun_out/bench_v$i.err
  python - "$v" gpurun_out/bench_v$i.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "| value", round(j["value"], 1), "| kernel_ms", {k: round(v, 3) for k, v in j["kernel_ms_per_step"].items()},
          "| q/s", round(j["query_sketches_per_s"]), "| query frac", round(j["roofline_query"]["frac"], 3), "| scan frac", round(j["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[1], "bench failed", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-1500:])
PY
done
