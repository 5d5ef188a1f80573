#!/bin/bash
# Round-1 measurement run: regression tests, full-size bench (both arms), filter knob A/B,
# ncu launch list + full captures of the two hot kernels, PCIe bandwidth probe.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python - > gpurun_out/pcie.txt 2>&1 <<'PY'
import torch, time
h = torch.empty(4 << 30, dtype=torch.uint8, pin_memory=True)
d = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
for name, a, b in (("h2d", d, h), ("d2h", h, d)):
    a.copy_(b); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    print(name, "GB/s", 3 * 4.294967296 / (time.perf_counter() - t))
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
NQ_SCAN_NOFILTER=1 timeout 600 python bench.py --genomes 2048 --queries 256 --no-e2e --no-cpu-baseline > gpurun_out/bench_nofilter.json 2> gpurun_out/bench_nofilter.err
timeout 600 python bench.py --genomes 2048 --queries 256 --no-e2e --no-cpu-baseline > gpurun_out/bench_filter.json 2> gpurun_out/bench_filter.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
  python bench.py --genomes 2048 --queries 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sketch_scan -s 1 -c 1 -o gpurun_out/prof_scan -f \
  python bench.py --genomes 1024 --queries 128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_scan.out 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:query_count -s 1 -c 1 -o gpurun_out/prof_query -f \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_query.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cell_sort -s 1 -c 1 -o gpurun_out/prof_cellsort -f \
  python bench.py --genomes 2048 --queries 256 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_cellsort.out 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/pcie.txt; cat gpurun_out/bench_ref.json; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
echo; python - <<'PY'
import json
for f in ("bench_nofilter", "bench_filter"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json")); print(f, j["value"], j["kernel_ms_per_step"], j["query_sketches_per_s"], j["roofline_query"]["frac"])
    except Exception as e: print(f, "failed", e)
PY
ls -la gpurun_out/
