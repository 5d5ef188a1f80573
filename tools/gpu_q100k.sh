#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload q100k --steps 3 --warmup 2 ${Q_ARGS} > gpurun_out/bench_q100k.json 2> gpurun_out/bench_q100k.err
echo "rc $?"; cat gpurun_out/bench_q100k.json; tail -5 gpurun_out/bench_q100k.err
