#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? : $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head -5
for v in "NQ_QUERY_MID=1" "NQ_QUERY_MID=0"; do
for g in 25000 50000; do
  env $v timeout 600 python bench.py --workload q100k --genomes $g --steps 3 --warmup 2 > gpurun_out/q_$g.json 2> gpurun_out/q_$g.err
  python - "gpurun_out/q_$g.json" $g "$v" <<'PY'
import json, sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j["roofline"]
    print(sys.argv[3], "10k queries vs", sys.argv[2], "genomes | q/s", round(j["value"]), "| query ms", round(r["ms_per_launch"],3), "| frac", round(r["frac"],3), "| hits", j["first_hits"])
except Exception as e:
    print("failed", e); print(open(sys.argv[1].replace(".json",".err")).read()[-800:])
PY
done
done
timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/c2_default.json 2> gpurun_out/c2_default.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/c2_default.json").read().strip().splitlines()[-1]); r=j["roofline_query"]
print("c2 default | value", round(j["value"],1), "| query ms", round(r["ms_per_launch"],3), "| frac", round(r["frac"],3), r["kernel"])
PY
