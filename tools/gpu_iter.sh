#!/bin/bash
# Iteration run: parity tests, a mid-size bench, optional ncu capture of one kernel family.
#   NCU_KERNEL=sketch_scan|query_count|cell_sort (optional)   BENCH_ARGS="..." (optional)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
BA=${BENCH_ARGS:---genomes 2048 --queries 256 --no-e2e --no-cpu-baseline}
timeout 900 python bench.py $BA > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err
python - <<'PY'
import json
try:
    j = json.load(open("gpurun_out/bench_iter.json"))
    print("value", j["value"], "kernel_ms", j["kernel_ms_per_step"], "q/s", j["query_sketches_per_s"])
    print("roofline", j["roofline"]["frac"], j["roofline"]["gbases_per_s"], "query frac", j["roofline_query"]["frac"], "clocks", j["clocks"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench_iter.err").read()[-2000:])
PY
for KN in $NCU_KERNEL; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KN -s 1 -c 1 -o gpurun_out/prof_$KN -f \
    python bench.py ${NCU_BENCH_ARGS:---genomes 1024 --queries 128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline} > gpurun_out/ncu_$KN.out 2>&1
  tail -2 gpurun_out/ncu_$KN.out
done
