#!/bin/bash
# Parity of every gather form of the query kernel, then the forms side by side on the two query workloads.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/seg_gpu.txt
for form in default seg8 seg16 seg32 stream; do
  if [ $form = default ]; then unset NQ_QUERY_FORM; else export NQ_QUERY_FORM=$form; fi
  sel=""
  [ $form != default ] && sel=1; selk="query or golden or c1 or large_n or lines or sharded or device_pointer"
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 ${sel:+-k "$selk"} > gpurun_out/pytest_$form.log 2>&1
  echo "form=$form pytest exit $? : $(tail -1 gpurun_out/pytest_$form.log)"
done
unset NQ_QUERY_FORM
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
for form in stream seg32 seg16; do
  NQ_QUERY_FORM=$form timeout 600 python bench.py --workload q100k --steps 3 --warmup 2 > gpurun_out/q100k_$form.json 2> gpurun_out/q100k_$form.err
  summ "q100k $form" gpurun_out/q100k_$form.json
done
for form in stream seg8 seg16 seg32; do
  NQ_QUERY_FORM=$form timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/c2_$form.json 2> gpurun_out/c2_$form.err
  summ "c2 $form" gpurun_out/c2_$form.json
done
for form in stream seg8 seg16; do
  NQ_QUERY_FORM=$form timeout 600 python bench.py --no-e2e --no-cpu-baseline --genomes 12500 --queries 10000 > gpurun_out/c3s_$form.json 2> gpurun_out/c3s_$form.err
  summ "c3shard $form" gpurun_out/c3s_$form.json
done
