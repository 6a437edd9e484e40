"""Dump selected raw metrics of every kernel in an .ncu-rep (memory-system view).
Usage: python benchmarks/ncu_dump.py file.ncu-rep [regex]"""
import csv
import io
import re
import subprocess
import sys

PAT = sys.argv[2] if len(sys.argv) > 2 else (
    r"gpu__time_duration.sum|launch__(grid|block)_size|registers_per_thread|sm__warps_active.avg.pct|"
    r"issue_active.avg.pct|smsp__inst_executed.sum$|l1tex__t_(sectors|requests)_pipe_lsu_mem_global_op_(ld|red|atom|st).sum$|"
    r"l1tex__t_sector_hit_rate|l1tex__throughput|lts__throughput|lts__t_sectors.sum$|lts__t_sectors_op_(read|write|red|atom).sum$|"
    r"lts__t_sector_hit_rate|lts__t_requests.sum$|dram__bytes_(read|write).sum$|breakdown|"
    r"l1tex__m_(xbar2l1tex_read_sectors|l1tex2xbar_write_sectors).sum$|l1tex__lsu_writeback|l1tex__data_pipe_lsu_wavefronts.sum$|"
    r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|"
    r"smsp__average_warps_issue_stalled.*per_issue_active|lts__d_sectors|lts__t_sectors_srcunit_tex.sum$|"
    r"lts__t_sectors_srcunit_tex_op_(read|red|atom|write).sum$|sm__inst_executed_pipe_lsu|lts__average_t_sector")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, rows = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
rx = re.compile(PAT)
for r in rows:
    print("=" * 100)
    print(r[col["Kernel Name"]][:160])
    for n, i in col.items():
        if rx.search(n) and r[i] not in ("", "n/a"):
            print("  %-82s %s %s" % (n, r[i], units[i]))
