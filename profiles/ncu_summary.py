#!/usr/bin/env python
"""profiles/ncu_summary.py REPORT.ncu-rep [regex ...] -- metric: value (unit) lines of an ncu report.

`ncu --page raw --csv` prints one wide row per kernel; this turns it into one line per metric and keeps
those whose name matches one of the regular expressions (default: the set the roofline discussion uses)."""
import csv
import io
import re
import subprocess
import sys

DEFAULT = [r"^gpu__time_duration\.sum", r"dram__bytes_(read|write)\.sum$", r"dram__throughput\.avg\.pct", r"gpu__dram_throughput",
           r"launch__(registers_per_thread|grid_size|block_size|occupancy_limit|shared_mem_per_block_dynamic|waves)",
           r"smsp__issue_active\.avg\.pct", r"smsp__inst_executed\.sum$", r"sm__warps_active\.avg\.pct",
           r"smsp__average_warps?_issue_stalled_.*_per_issue_active", r"lts__t_sector_hit_rate\.pct", r"lts__t_bytes\.sum$",
           r"lts__t_sectors_srcunit_tex_op_(read|write)\.sum$", r"l1tex__data_bank_conflicts_pipe_lsu\.sum$",
           r"sm__throughput\.avg\.pct", r"sm__cycles_elapsed\.max", r"smsp__cycles_active\.avg"]


def main():
    rep = sys.argv[1]
    pats = [re.compile(p) for p in (sys.argv[2:] or DEFAULT)]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units = rows[0], rows[1]
    for row in rows[2:]:
        print(f"# kernel: {row[names.index('Kernel Name')]}  grid {row[names.index('Grid Size')]}  block {row[names.index('Block Size')]}")
        for n, u, v in zip(names, units, row):
            short = n.split(".", 2)[-1] if n.count(".") >= 2 and n.split(".")[1].startswith("Triage") else n
            if any(p.search(n) or p.search(short) for p in pats):
                print(f"{n}: {v} {u}".rstrip())


if __name__ == "__main__":
    main()
