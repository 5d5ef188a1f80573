#!/bin/bash
mkdir -p gpurun_out
for v in "NQ_X=1" "NQ_SCAN_FILTER=0" "NQ_QUERY_FORM=stream" "NQ_SCAN_FILTER=0 NQ_QUERY_FORM=stream"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 500 -k "index_query_matrix_vs_oracle or large_S" > gpurun_out/pt_$tag.log 2>&1
  echo "$v : exit $? : $(tail -1 gpurun_out/pt_$tag.log)"
done
