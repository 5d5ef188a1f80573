#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_pack.py 4e9 > gpurun_out/r2i_pack.log 2>&1
cat gpurun_out/r2i_pack.log
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 200 -k "packed" ) > gpurun_out/r2i_pytest.log 2>&1
tail -5 gpurun_out/r2i_pytest.log
