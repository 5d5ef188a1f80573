#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the few numbers the roofline discussion needs.

usage: python tools/ncu_summary.py gpurun_out/prof_scan.ncu-rep [...] > profiles/rNN_xxx.txt
Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and prints, per captured launch, the
duration, DRAM traffic, pipe utilisation, issue utilisation, occupancy, registers and the top
warp-stall reasons.
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"== {rep}: no data")
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            u = dict(zip(hdr, units))
            print(f"== {rep}: {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
            for k in WANT:
                if k in d and d[k] != "":
                    print(f"  {k:78s} {d[k]:>18s} {u.get(k, '')}")
            stalls = [(float(v), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and
                      k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
            for v, k in sorted(stalls, reverse=True)[:7]:
                name = k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
                print(f"  stall {name:72s} {v:18.3f} warps/issue")
            print()


if __name__ == "__main__":
    main()
