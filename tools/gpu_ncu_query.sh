#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-query_count} -s ${SKIP:-1} -c 1 -o gpurun_out/prof_${TAG:-q} -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ncu_${TAG:-q}.out 2>&1
tail -3 gpurun_out/ncu_${TAG:-q}.out
