#!/bin/bash
# ncu of the S=18 (global-memory sketch) scan kernel on a small configs[4] slice
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sketch_scan -s 0 -c 1 -o gpurun_out/prof_scan_s18 -f \
    python bench.py --workload c5 --genomes 592 --steps 1 > gpurun_out/ncu_scan_s18.out 2>&1
tail -2 gpurun_out/ncu_scan_s18.out | cut -c1-200
