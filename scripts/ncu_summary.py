"""Condense an `ncu --page raw --csv` export into one line per launch with the metrics the roofline discussion needs."""
import csv
import re
import sys

KEEP = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB_rd"), ("dram__bytes_write.sum", "MB_wr"),
        ("lts__t_bytes.sum", "MB_l2"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "hmma%"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct", "st_longsb"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_longsb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_shortsb"),
        ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%")]


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in KEEP if m in idx]
    print("kernel".ljust(64), " ".join(n.rjust(9) for _, n in cols))
    for r in data:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("dgsct::", "")[:64]
        vals = []
        for m, n in cols:
            v = r[idx[m]].replace(",", "")
            try:
                f = float(v)
                u = units[idx[m]]
                if m.startswith("gpu__time"):
                    f = f / 1e3 if u == "ns" else (f * 1e3 if u == "ms" else f)
                if "bytes" in m:
                    f = f / 1e6 if u in ("byte", "B") else (f / 1e3 if u == "Kbyte" else (f * 1e3 if u == "Gbyte" else f))
                vals.append(f"{f:9.1f}")
            except ValueError:
                vals.append(v[:9].rjust(9))
        print(name.ljust(64), " ".join(vals))


if __name__ == "__main__":
    main()
