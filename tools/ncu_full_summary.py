#!/usr/bin/env python
"""Markdown summary of `ncu --set full` reports: one table per report, one column per captured launch.
usage: tools/ncu_full_summary.py title report1.ncu-rep [report2.ncu-rep ...] > profiles/xxx.md"""
import csv, subprocess, sys
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"), ("launch__waves_per_multiprocessor", "waves / SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no instruction"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
print(f"# {sys.argv[1]}\n")
print("Captured with `ncu --set full --clock-control none --import-source on` (tools/gpu_round2_final.sh); numbers under the profiler are "
      "cold-cache and serialised, never bench values.\n")
for rep in sys.argv[2:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units, data = rows[0], rows[1], rows[2:]
    print(f"## `{rep.split('/')[-1]}`\n")
    names = [r[h.index("Kernel Name")].replace("void ", "").split("(")[0] for r in data]
    print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|" + "---:|" * len(names))
    for key, label in KEYS:
        if key not in h:
            continue
        i = h.index(key)
        vals = []
        for r in data:
            try:
                v = float(r[i].replace(",", ""))
                vals.append(f"{v:,.2f}".rstrip("0").rstrip(".") + (" " + units[i] if units[i] not in ("", "%") else ""))
            except ValueError:
                vals.append(r[i])
        print(f"| {label} | " + " | ".join(vals) + " |")
    print()
