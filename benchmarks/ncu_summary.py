"""Summarise an .ncu-rep: per kernel the headline metrics, stall breakdown and hottest source lines.
Usage: python benchmarks/ncu_summary.py file.ncu-rep [--lines N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_op_shared_atom.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
        "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warps_eligible.avg.per_cycle_active"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    col = {n: i for i, n in enumerate(hdr)}
    for r in rows:
        print("=" * 100)
        print(r[col["Kernel Name"]][:150])
        for k in KEYS:
            if k in col:
                print("  %-70s %s %s" % (k, r[col[k]], units[col[k]]))
        stalls = [(float(r[i].replace(",", "")), n) for n, i in col.items()
                  if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio") and r[i]]
        stalls.sort(reverse=True)
        print("  stall reasons (warps per issue):")
        for v, n in stalls[:8]:
            print("    %6.2f  %s" % (v, n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))


if __name__ == "__main__":
    main()
