#!/usr/bin/env python
"""Summarise `ncu --set full` reports (the numbers DESIGN.md / bench.py quote) as markdown.
usage: python tools/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/<name>.md"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "duration"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->L1 read"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
        ("launch__registers_per_thread", "registers / thread"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue")]

print("# ncu --set full summaries (--clock-control none; one launch each, replayed ~40x: times are cold-cache)\n")
for path in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## `{name[:110]}`  ({path.split('/')[-1]})\n")
        print("| metric | value |\n|---|---|")
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {label} (`{k}`) | {r[i]} {units[i]} |")
        print()
