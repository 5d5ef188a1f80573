#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_pack.py 4e9 > gpurun_out/r2n_pack.log 2>&1; cat gpurun_out/r2n_pack.log
timeout 600 python tools/bench_cli_ingest.py 64 > gpurun_out/r2n_cli_ingest.json 2> gpurun_out/r2n_cli_ingest.err; cat gpurun_out/r2n_cli_ingest.json; tail -3 gpurun_out/r2n_cli_ingest.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2n_smoke.log 2>&1; tail -2 gpurun_out/r2n_smoke.log
( timeout 600 python bench.py --workload c5 --steps 2 ) > gpurun_out/r2n_c5_1gpu.json 2> gpurun_out/r2n_c5_1gpu.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2n_c5_1gpu.json').read().splitlines() if l.startswith('{')][0])
print('c5', d['value'], d['ms_per_step'], d['matrix_device_ms_per_step'], d['roofline']['frac'], d['diag_is_F_mod_65536'])
"
