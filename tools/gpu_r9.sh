#!/bin/bash
# fused warp-per-read kernel + sliced finish of global-counter queries: parity, then configs[3] with and without the reads kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? : $(tail -1 gpurun_out/pytest_gpu.log)"; grep -E "FAILED|Error|assert" gpurun_out/pytest_gpu.log | head -10
for v in "NQ_READS_KERNEL=1" "NQ_READS_KERNEL=0"; do
env $v timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 --queries 2000 > gpurun_out/c4_$v.json 2> gpurun_out/c4_$v.err
echo "c4 $v exit $?"; python - "gpurun_out/c4_$v.json" <<'PY'
import json, sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("c4 value", round(j["value"],2), "Gbases/s | reads/s", round(j["reads_per_s"]/1e6,1), "M | ms", round(j["ms_per_step"],1), "| kernels", {k: round(v,1) for k,v in j["kernel_ms_per_step"].items()}, "| q/s", round(j["query_sketches_per_s"],1), "| e2e", round(j["e2e"]["value"],2), "| hits", j["first_hits"])
PY
tail -3 "gpurun_out/c4_$v.err"
done
