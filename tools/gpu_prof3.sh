#!/bin/bash
# ncu --set full captures at the full C2 size: query kernel, index build kernel, scan kernel (index launch)
mkdir -p gpurun_out
for spec in "query_count:1:q5" "cell_build:1:cb5" "sketch_scan:2:scanfull"; do
  IFS=: read k skip tag <<< "$spec"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/prof_$tag -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_$tag.out 2>&1
  tail -1 gpurun_out/ncu_$tag.out
done
