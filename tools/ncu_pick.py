"""Prints selected metrics of every kernel in an ncu raw CSV page: ncu -i X.ncu-rep --page raw --csv | python tools/ncu_pick.py [substr ...]"""
import csv
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
           "gpu__dram_throughput", "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "issue_active.avg.pct",
           "sm__pipe_tensor_cycles_active", "pipe_fma_cycles_active.avg.pct", "pipe_alu_cycles_active.avg.pct",
           "launch__registers_per_thread", "launch__occupancy_limit", "lts__t_sector_hit_rate.pct",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "issue_stalled", "sm__inst_executed.sum ",
           "lts__t_bytes.sum ", "achieved_occupancy"]


def main():
    pats = sys.argv[1:] or DEFAULT
    rows = list(csv.reader(sys.stdin))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("---", r[ki][:60], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for i, h in enumerate(hdr):
            if any(p.strip() in h for p in pats) and r[i] not in ("", "0", "n/a"):
                print("  %-95s %s %s" % (h[-95:], r[i], units[i]))


if __name__ == "__main__":
    main()
