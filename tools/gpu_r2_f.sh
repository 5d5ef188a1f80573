#!/bin/bash
mkdir -p gpurun_out
export NIQKI_B200_LIB=$PWD/niqki_b200/lib_tuning/libniqki_b200.so
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload q100k --genomes ${G:-10000} --queries ${Q:-1000} --steps 5 --warmup 3 > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err
}
run base X=1
run l2f32 NQ_L2_FETCH=32
run pred NQ_QUERY_EXP=4
run pred_l2f32 NQ_QUERY_EXP=4 NQ_L2_FETCH=32
run nrf0 NQ_SLAB_NRF=0
run nrf4 NQ_SLAB_NRF=4
run csr NQ_SLAB=0
G=12500 Q=10000 run base_c3 X=1
G=12500 Q=10000 run l2f32_c3 NQ_L2_FETCH=32
G=12500 Q=10000 run pred_c3 NQ_QUERY_EXP=4
G=12500 Q=10000 run nrf3_c3 NQ_SLAB_NRF=3
G=12500 Q=10000 run nrf0_c3 NQ_SLAB_NRF=0
G=12500 Q=10000 run csr_c3 NQ_SLAB=0
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline')
        print(f, 'ms',round(d['ms_per_step'],4),'frac',round(r.get('frac'),4), 'kern ms', round(r.get('ms_per_launch'),4))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-300:])
PY
