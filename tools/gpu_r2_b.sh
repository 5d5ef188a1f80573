#!/bin/bash
# round 2: what bounds the slab query kernel (measurement build: NQ_* knobs compiled in)
mkdir -p gpurun_out
export NIQKI_B200_LIB=$PWD/niqki_b200/lib_tuning/libniqki_b200.so
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload q100k --genomes ${G:-10000} --queries ${Q:-1000} --steps 5 --warmup 3 > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err
}
run base X=1
run nocount NQ_QUERY_EXP=1
run nogather NQ_QUERY_EXP=2
run neither NQ_QUERY_EXP=3
run nopf NQ_QUERY_PF_AHEAD=0
run nt256 NQ_QUERY_NT=256
run g16 NQ_SLAB=16
run csr NQ_SLAB=0
G=12500 Q=10000 run base_c3 X=1
G=12500 Q=10000 run nocount_c3 NQ_QUERY_EXP=1
G=12500 Q=10000 run neither_c3 NQ_QUERY_EXP=3
G=12500 Q=10000 run nt128_c3 NQ_QUERY_NT=128
timeout 600 ncu --set full --clock-control none --import-source on -k regex:query_slab -c 1 -o gpurun_out/r2b_slab \
  python bench.py --workload q100k --genomes 10000 --queries 1000 --steps 1 --warmup 0 > gpurun_out/r2b_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d.get('roofline')
        print(f, 'ms',round(d['ms_per_step'],4),'frac',r.get('frac'), 'build_s', d.get('index_build_wall_s'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-300:])
PY
