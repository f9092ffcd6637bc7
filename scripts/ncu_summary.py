"""Summaries for profiles/ from ncu outputs (no GPU needed to run this).

    python scripts/ncu_summary.py rep  gpurun_out/x.ncu-rep  > profiles/rNN_x_ncu.txt      longest launch per kernel name
    python scripts/ncu_summary.py list gpurun_out/launches.csv > profiles/rNN_launch_list_summary.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict, defaultdict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(header)}
    seen = OrderedDict()  # per kernel name: the longest captured launch (gated kernels also have empty ones)
    dur = col.get("gpu__time_duration.sum")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        key = float(r[dur].replace(",", "")) if dur is not None and r[dur] else 0.0
        if name not in seen or key > seen[name][0]:
            seen[name] = (key, r)
    seen = OrderedDict((k, v[1]) for k, v in seen.items())
    for name, r in seen.items():
        print(f"===== {name}")
        for m in METRICS:
            if m in col:
                print(f"  {m:<86s} {r[col[m]]:>16s} {units[col[m]]}")
        if "dram__bytes_read.sum" in col:
            print()
    return 0


def launch_list(path):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit() or (r and r[0] == "ID")]
    header = rows[0]
    col = {h: i for i, h in enumerate(header)}
    agg = defaultdict(list)
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        t = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        t *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
        agg[r[col["Kernel Name"]]].append(t)
    total = sum(sum(v) for v in agg.values())
    print(f"{'kernel':<52s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'max us':>9s} {'share':>7s}")
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{name[:52]:<52s} {len(v):>8d} {sum(v):>10.1f} {sum(v) / len(v):>9.2f} {max(v):>9.1f} {100 * sum(v) / total:>6.1f}%")
    return 0


if __name__ == "__main__":
    if len(sys.argv) != 3 or sys.argv[1] not in ("rep", "list"):
        sys.exit(__doc__)
    sys.exit(rep(sys.argv[2]) if sys.argv[1] == "rep" else launch_list(sys.argv[2]))
