#!/bin/bash
# split16 gather (u16 copy of the postings for 65.6k..131k genomes): parity, then the 100k-genome query with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "synthetic or large_n or beyond or matrix" > gpurun_out/pytest_split.log 2>&1
echo "pytest exit $? : $(tail -1 gpurun_out/pytest_split.log)"
summ() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = j.get("roofline_query") or j["roofline"]
    print(sys.argv[1], "| value", round(j["value"], 1), "| query ms", round(r["ms_per_launch"], 3), "| frac", round(r["frac"], 3),
          "| e2e", round(j["e2e"]["value"], 1), "| build s", j.get("index_build_wall_s"), "| first_hits", j.get("first_hits"))
except Exception as e:
    print(sys.argv[1], "bench failed", e); print(open(sys.argv[2].replace(".json", ".err")).read()[-1500:])
PY
}
for v in "NQ_SPLIT16=1" "NQ_SPLIT16=0"; do
  env $v timeout 600 python bench.py --workload q100k --steps 3 --warmup 2 > gpurun_out/q100k_$v.json 2> gpurun_out/q100k_$v.err
  summ "q100k $v" gpurun_out/q100k_$v.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:query_count -s 2 -c 1 -o gpurun_out/prof_q100k_split16 -f \
    python bench.py --steps 1 --warmup 1 --workload q100k > gpurun_out/ncu_q100k_split16.out 2>&1
tail -1 gpurun_out/ncu_q100k_split16.out | cut -c1-200
