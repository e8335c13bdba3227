"""Summarise .ncu-rep captures into small text files for profiles/ (runs without a GPU).

    python -m stencil_benchmarks_b200.scripts.ncu_summary gpurun_out/hdiff_tma_r01.ncu-rep [...] --out profiles
"""

import argparse
import csv
import io
import pathlib
import subprocess

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__cycles_elapsed.max",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
    "launch__block_size",
    "launch__grid_size",
    "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def summarise(report):
    raw = subprocess.run(["ncu", "-i", str(report), "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    lines = []
    for values in rows[2:]:
        record = dict(zip(header, values))
        lines.append(f"kernel: {record.get('Kernel Name', '?')}")
        lines.append(f"grid {record.get('Grid Size', '?')} block {record.get('Block Size', '?')}")
        unit = dict(zip(header, units))
        for metric in METRICS:
            if metric in record:
                lines.append(f"  {metric:78s} {record[metric]:>16s} {unit[metric]}")
        read = float(record.get("dram__bytes_read.sum", "nan"))
        write = float(record.get("dram__bytes_write.sum", "nan"))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        total = read * scale.get(unit.get("dram__bytes_read.sum", ""), 1) + \
            write * scale.get(unit.get("dram__bytes_write.sum", ""), 1)
        lines.append(f"  {'dram traffic (read + write), bytes per launch':78s} {total:16.0f}")
        lines.append("")
    return "\n".join(lines)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("reports", nargs="+")
    parser.add_argument("--out", default="profiles")
    args = parser.parse_args()
    out = pathlib.Path(args.out)
    out.mkdir(exist_ok=True)
    for report in args.reports:
        text = summarise(report)
        target = out / (pathlib.Path(report).stem + ".txt")
        target.write_text(f"# ncu --set full --clock-control none, from {pathlib.Path(report).name}\n" + text)
        print(target)
        print(text)


if __name__ == "__main__":
    main()
