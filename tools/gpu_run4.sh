#!/bin/bash
# Round-1 re-entry measurement: regression tests, both bench arms at full size, ncu launch list of
# the bench command, full captures of the hot kernels.
mkdir -p gpurun_out
nproc > gpurun_out/env.txt; free -g >> gpurun_out/env.txt; nvidia-smi >> gpurun_out/env.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.out 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:query_count -s 1 -c 1 -o gpurun_out/prof_query -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_query.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sketch_scan -s 1 -c 1 -o gpurun_out/prof_scan -f \
  python bench.py --genomes 1024 --queries 128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_scan.out 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_ref.json; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
ls -la gpurun_out/
