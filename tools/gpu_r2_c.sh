#!/bin/bash
# round 2: resolve + count split of the slab query; full parity suite
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2c_pytest.log 2>&1
tail -5 gpurun_out/r2c_pytest.log
export NIQKI_B200_LIB=$PWD/niqki_b200/lib_tuning/libniqki_b200.so
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload q100k --genomes ${G:-10000} --queries ${Q:-1000} --steps 5 --warmup 3 > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err
}
run base X=1
run nocount NQ_QUERY_EXP=1
run nogather NQ_QUERY_EXP=2
run neither NQ_QUERY_EXP=3
run nopf NQ_QUERY_PF_AHEAD=0
run nopf_neither NQ_QUERY_PF_AHEAD=0 NQ_QUERY_EXP=3
run nt256 NQ_QUERY_NT=256
G=12500 Q=10000 run base_c3 X=1
G=12500 Q=10000 run nopf_c3 NQ_QUERY_PF_AHEAD=0
G=12500 Q=10000 run nt128_c3 NQ_QUERY_NT=128
G=12500 Q=10000 run nt128_nopf_c3 NQ_QUERY_NT=128 NQ_QUERY_PF_AHEAD=0
G=25000 Q=10000 run base_25k X=1
G=25000 Q=10000 run csr_25k NQ_SLAB=0
G=50000 Q=10000 run base_50k X=1
G=50000 Q=10000 run csr_50k NQ_SLAB=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"query_slab|slab_resolve" -c 2 -o gpurun_out/r2c_slab \
  python bench.py --workload q100k --genomes 10000 --queries 1000 --steps 1 --warmup 0 > gpurun_out/r2c_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline')
        print(f, 'ms',round(d['ms_per_step'],4),'frac',r.get('frac'), 'build_s', d.get('index_build_wall_s'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-300:])
PY
