#!/bin/bash
# 2-GPU weak-scaling check of the bench (the driver's own launch line), plus the reference arm under torchrun.
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "rc $?"; cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
