#!/bin/bash
# chunked (cell, chunk) counting sort for few-cell / many-entry index builds: parity + configs[3]
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? : $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 --queries ${C4Q:-2000} > gpurun_out/c4.json 2> gpurun_out/c4.err
echo "c4 exit $?"; python - <<'PY'
import json
j=json.loads(open("gpurun_out/c4.json").read().strip().splitlines()[-1])
print("c4 value", round(j["value"],2), "Gbases/s | reads/s", round(j["reads_per_s"]/1e6,1), "M | ms", round(j["ms_per_step"],1), "| kernels", {k: round(v,1) for k,v in j["kernel_ms_per_step"].items()}, "| q/s", round(j["query_sketches_per_s"],1), "| hits", j["first_hits"])
PY
tail -3 gpurun_out/c4.err
