#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "synthetic or large_n" > gpurun_out/pytest_split.log 2>&1
echo "pytest exit $? : $(tail -1 gpurun_out/pytest_split.log)"
timeout 600 python bench.py --workload q100k --steps 3 --warmup 2 > gpurun_out/q100k_v3.json 2> gpurun_out/q100k_v3.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/q100k_v3.json").read().strip().splitlines()[-1])
r=j["roofline"]; print("q100k value", round(j["value"]), "| query ms", round(r["ms_per_launch"],3), "| frac", round(r["frac"],3), "| e2e", round(j["e2e"]["value"]), "| hits", j["first_hits"])
PY
