#!/bin/bash
# New parity tests + ncu --set full of the query kernel: configs[2] shard shape (one wave), configs[1] in the seg8 form,
# the 100k-genome index in the seg32 form.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "matrix or beyond_32" > gpurun_out/pytest_new.log 2>&1
echo "pytest exit $? : $(tail -1 gpurun_out/pytest_new.log)"
cap() {  # tag skip env bench-args...
  tag=$1; skip=$2; envs=$3; shift 3
  env $envs timeout 900 ncu --set full --clock-control none --import-source on -k regex:query_count -s $skip -c 1 -o gpurun_out/prof_$tag -f \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline "$@" > gpurun_out/ncu_$tag.out 2>&1
  echo "$tag: $(tail -1 gpurun_out/ncu_$tag.out | cut -c1-200)"
}
cap c3s_stream 9 NQ_QUERY_FORM=stream --genomes 12500 --queries 10000
cap c2_seg8 1 NQ_QUERY_FORM=seg8
cap q100k_seg32 2 NQ_X=1 --workload q100k
ls -la gpurun_out/*.ncu-rep
