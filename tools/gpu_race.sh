#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the kernels written this round: segment-table query forms,
# warp-per-read sketch + densification, chunked sort, matrix extraction, the filtered S > 15 scan
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 1400 \
  -k "golden_small or reads_lines or matrix_rows or densify_stall or (index_query_matrix and not ps4)" > gpurun_out/racecheck.log 2>&1
echo "racecheck exit $? : $(grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/racecheck.log | tail -3 | tr '\n' ' ')"
grep -E "Race reported|hazard" gpurun_out/racecheck.log | sort | uniq -c | head -20
