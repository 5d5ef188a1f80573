#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 200 -k "large_S or packed_path_random" ) > gpurun_out/r2s_pytest.log 2>&1; tail -2 gpurun_out/r2s_pytest.log
export NIQKI_B200_LIB=$PWD/niqki_b200/lib_tuning/libniqki_b200.so
for v in default 1; do
  if [ $v = default ]; then unset NQ_SCAN_GREAD; else export NQ_SCAN_GREAD=$v; fi
  timeout 300 python bench.py --workload c5 --genomes 4000 --steps 1 > gpurun_out/r2s_c5_gread_$v.json 2> gpurun_out/r2s_c5_gread_$v.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2s_c5_gread_$v.json').read().splitlines() if l.startswith('{')][0])
print('gread=$v sketch Gbases/s', d['sketch_gbases_per_s'], 'diag', d['diag_is_F_mod_65536'])
"
done
