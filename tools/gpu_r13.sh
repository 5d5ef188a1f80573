#!/bin/bash
# coarse filter of the S > 15 scan + mid-size query CTAs: parity, then c5 slice with/without the filter, 25k/50k-genome shard queries
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? : $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head -5
for v in "NQ_SCAN_FILTER=1" "NQ_SCAN_FILTER=0"; do
  env $v timeout 900 python bench.py --workload c5 --genomes 4000 --steps 1 > gpurun_out/c5_$v.json 2> gpurun_out/c5_$v.err
  python - "gpurun_out/c5_$v.json" "$v" <<'PY'
import json, sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("c5(4000)", sys.argv[2], "| sketch Gb/s", round(j["sketch_gbases_per_s"],1), "| matrix device ms", round(j["matrix_device_ms_per_step"],1), "| diag", j["diag_is_F_mod_65536"])
PY
done
for g in 25000 50000; do
  timeout 600 python bench.py --workload q100k --genomes $g --steps 3 --warmup 2 > gpurun_out/q_$g.json 2> gpurun_out/q_$g.err
  python - "gpurun_out/q_$g.json" $g <<'PY'
import json, sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=j["roofline"]
print("10k queries vs", sys.argv[2], "genomes | q/s", round(j["value"]), "| query ms", round(r["ms_per_launch"],3), "| frac", round(r["frac"],3), "| hits", j["first_hits"])
PY
done
