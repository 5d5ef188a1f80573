#!/bin/bash
mkdir -p gpurun_out
summ() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = j.get("roofline_query") or j["roofline"]
    print(sys.argv[1], "| value", round(j["value"], 1), "| query ms", round(r["ms_per_launch"], 3), "| frac", round(r["frac"], 3), "| kernel_ms", {k: round(v,2) for k,v in j["kernel_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "bench failed", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-1500:])
PY
}
for form in stream seg8 stream; do
  NQ_QUERY_FORM=$form timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/c2_$form.json 2> gpurun_out/c2_$form.err
  summ "c2 $form" gpurun_out/c2_$form.json
done
timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/c2_default.json 2> gpurun_out/c2_default.err
summ "c2 default" gpurun_out/c2_default.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:query_count -s 1 -c 1 -o gpurun_out/prof_c2_stream_now -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_c2_stream_now.out 2>&1
tail -1 gpurun_out/ncu_c2_stream_now.out | cut -c1-150
