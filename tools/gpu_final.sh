#!/bin/bash
# Round-end record on one B200: smoke, both bench arms, launch list of the bench command, configs[4], memcheck of a test subset.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $? : $(tail -1 gpurun_out/final_smoke.log)"
timeout 900 python bench.py --impl reference > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "reference exit $?"
timeout 1200 python bench.py > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err; echo "bench exit $?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/final_c2.json").read().strip().splitlines()[-1])
print("c2 value", round(j["value"],1), "| ms", round(j["ms_per_step"],2), "| e2e", round(j["e2e"]["value"],2), "| scan frac", round(j["roofline"]["frac"],3),
      "| query frac", round(j["roofline_query"]["frac"],3), "| launches", j["gpu_launches"], "| cpu", j.get("cpu_baseline",{}).get("value"))
r=json.loads(open("gpurun_out/final_reference.json").read().strip().splitlines()[-1]); print("reference arm", round(r["value"],3), r["unit"], r["cpu_baseline"]["cores"], "cores")
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/final_launches.out 2>&1; echo "launch list exit $? ($(wc -l < gpurun_out/final_launches.csv) lines)"
timeout 1200 python bench.py --workload c5 --steps 1 > gpurun_out/final_c5.json 2> gpurun_out/final_c5.err; echo "c5 exit $?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/final_c5.json").read().strip().splitlines()[-1])
print("c5 pairs/s", round(j["value"]/1e6,1), "M | ms", round(j["ms_per_step"],1), "| device ms", round(j["matrix_device_ms_per_step"],1), "| sketch Gb/s", round(j["sketch_gbases_per_s"],1), "| diag", j["diag_is_F_mod_65536"])
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 800 \
  -k "golden_small or matrix_rows or reads_lines or reads_kernel or synthetic_sketches-5000 or sharded" > gpurun_out/final_memcheck.log 2>&1
echo "memcheck exit $? : $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/final_memcheck.log | tail -3 | tr '\n' ' ')"
