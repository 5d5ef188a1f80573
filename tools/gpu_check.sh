#!/bin/bash
# First-contact GPU script: environment, smoke, parity tests, microbench, a small bench run.
mkdir -p gpurun_out
{
  echo "== env"; nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
  ls /root/reference 2>&1 | head -2
  echo "== smoke"
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15
} > gpurun_out/env_smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 tools/microbench > gpurun_out/microbench.txt 2>&1
timeout 900 python bench.py --genomes 1024 --queries 128 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
tail -5 gpurun_out/env_smoke.log; tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/microbench.txt; cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
