"""Developer tool: per-kernel summary table of `ncu --page raw --csv` exports of `--set full` captures
(tests/tools/ncu_full_round2.sh). Usage: summarize_ncu_full.py title=path.csv ... > profiles/rNN_ncu_full_summary.md"""
import csv
import sys
from collections import OrderedDict

COLS = [
    ("time us", "gpu__time_duration.sum", 1e-3),
    ("DRAM rd MB", "dram__bytes_read.sum", None),
    ("DRAM wr MB", "dram__bytes_write.sum", None),
    ("DRAM % peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("regs", "launch__registers_per_thread", 1),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    ("FP64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1),
    ("threads / inst", "smsp__thread_inst_executed_per_inst_executed.ratio", 1),
    ("stall long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1),
    ("stall short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", 1),
    ("stall math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1),
    ("stall wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1),
    ("stall barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1),
]


def to_mb(v, unit):
    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
    return v * f


def main():
    print("# `ncu --set full --clock-control none` of the round-2 kernels (per-kernel means over the captured launches)\n")
    print("Commands: `tests/tools/ncu_full_round2.sh` (each bench command plain first, then under ncu; `AAADMM_NO_GRAPH=1`). The "
          "`.ncu-rep` files exceed what gpurun brings back; their raw pages were exported on the box. Times are per launch under "
          "ncu (cold cache, serialised). Stall columns = warps stalled per issue-active cycle.\n")
    for arg in sys.argv[1:]:
        title, path = arg.split("=", 1)
        rows = list(csv.reader(open(path)))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
        ix = {h: i for i, h in enumerate(hdr)}
        groups = OrderedDict()
        for r in data:
            name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("aaadmm::<unnamed>::", "").replace("<unnamed>::", "")
            groups.setdefault(name, []).append(r)
        print("## %s\n" % title)
        print("| kernel | launches | " + " | ".join(c[0] for c in COLS) + " | grid x block |")
        print("|---|---:|" + "---:|" * len(COLS) + "---|")
        for name, rs in groups.items():
            cells = []
            for label, key, scale in COLS:
                if key not in ix:
                    cells.append("-")
                    continue
                vals = []
                for r in rs:
                    try:
                        v = float(r[ix[key]].replace(",", ""))
                    except ValueError:
                        continue
                    if scale is None:
                        v = to_mb(v, units[ix[key]])
                    elif key == "gpu__time_duration.sum":
                        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[ix[key]], 1e-3)
                    vals.append(v)
                cells.append("%.1f" % (sum(vals) / len(vals)) if vals else "-")
            g = rs[0][ix["Grid Size"]] if "Grid Size" in ix else ""
            b = rs[0][ix["Block Size"]] if "Block Size" in ix else ""
            print("| %s | %d | " % (name, len(rs)) + " | ".join(cells) + " | %s x %s |" % (g, b))
        print()


if __name__ == "__main__":
    main()
