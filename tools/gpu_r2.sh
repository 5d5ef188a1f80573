#!/bin/bash
# Full GPU parity suite (matrix now runs through the query kernels), wave-sized query launches on the
# configs[2] shard shape, first measurements of configs[3] (reads) and configs[4] (S=18 matrix).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? : $(tail -1 gpurun_out/pytest_gpu.log)"
summ() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = j.get("roofline_query") or j["roofline"]
    print(sys.argv[1], "| value", round(j["value"], 1), "| query ms", round(r["ms_per_launch"], 3), "| frac", round(r["frac"], 3),
          "| kernel_ms", {k: round(v, 2) for k, v in j.get("kernel_ms_per_step", {}).items()}, "| first_hits", j.get("first_hits"))
except Exception as e:
    print(sys.argv[1], "bench failed", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-1500:])
PY
}
for v in "NQ_QUERY_FORM=stream NQ_QUERY_WAVES=0" "NQ_QUERY_FORM=stream" "NQ_QUERY_FORM=seg8" "NQ_QUERY_FORM=seg16"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 600 python bench.py --no-e2e --no-cpu-baseline --genomes 12500 --queries 10000 > gpurun_out/c3s_$tag.json 2> gpurun_out/c3s_$tag.err
  summ "c3shard $v" gpurun_out/c3s_$tag.json
done
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/c4.json 2> gpurun_out/c4.err
echo "c4 exit $?"; tail -c 2500 gpurun_out/c4.json; tail -5 gpurun_out/c4.err
timeout 1200 python bench.py --workload c5 --steps 1 > gpurun_out/c5.json 2> gpurun_out/c5.err
echo "c5 exit $?"; tail -c 2500 gpurun_out/c5.json; tail -5 gpurun_out/c5.err
