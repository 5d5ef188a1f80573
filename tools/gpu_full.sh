#!/bin/bash
# Full GPU regression (all -m gpu tests incl. the CLI) + the default bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/bench_full.json").read().strip().splitlines()[-1])
    print("value", round(j["value"], 1), "e2e", j["e2e"] and round(j["e2e"]["value"], 1), "kernel_ms", {k: round(v, 3) for k, v in j["kernel_ms_per_step"].items()},
          "q/s", round(j["query_sketches_per_s"]), "query frac", round(j["roofline_query"]["frac"], 3), "scan frac", round(j["roofline"]["frac"], 3))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_full.err").read()[-1500:])
PY
