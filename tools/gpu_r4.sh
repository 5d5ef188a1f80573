#!/bin/bash
# dual-query counters + deeper ring: parity (default form = dual8 on small shards), then forms side by side
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? : $(tail -1 gpurun_out/pytest_gpu.log)"
NQ_QUERY_FORM=dual16 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "query or golden or c1 or lines or sharded or device_pointer or matrix" > gpurun_out/pytest_dual16.log 2>&1
echo "dual16 pytest exit $? : $(tail -1 gpurun_out/pytest_dual16.log)"
summ() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = j.get("roofline_query") or j["roofline"]
    print(sys.argv[1], "| value", round(j["value"], 1), "| query ms", round(r["ms_per_launch"], 3), "| frac", round(r["frac"], 3),
          "| first_hits", j.get("first_hits"))
except Exception as e:
    print(sys.argv[1], "bench failed", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-1500:])
PY
}
for form in dual8 dual16 seg8 stream; do
  NQ_QUERY_FORM=$form timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/c2_$form.json 2> gpurun_out/c2_$form.err
  summ "c2 $form" gpurun_out/c2_$form.json
done
for form in dual8 dual16 seg8; do
  NQ_QUERY_FORM=$form timeout 600 python bench.py --no-e2e --no-cpu-baseline --genomes 12500 --queries 10000 > gpurun_out/c3s_$form.json 2> gpurun_out/c3s_$form.err
  summ "c3shard $form" gpurun_out/c3s_$form.json
done
