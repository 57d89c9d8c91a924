#!/usr/bin/env python
"""Compact per-kernel summary of an ncu report (read on the CPU box):  python tools/ncu_summary.py prof.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit smem (blocks)"),
    ("launch__occupancy_limit_registers", "occupancy limit regs (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA inst % of peak"),
    ("smsp__issue_active.avg.pct", "issue slots busy %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none summary of %s (per launch; cold-cache, serialised replays)" % rep)
    for d in data:
        print("\n== %s" % d[idx["Kernel Name"]][:160])
        for key, label in WANT:
            if key in idx:
                print("   %-34s %s %s" % (label, d[idx[key]], units[idx[key]]))


if __name__ == "__main__":
    main()
