"""STREAM array-size sweep through the plugin class (BASELINE.json configs[1]: roofline calibration).

    python -m stencil_benchmarks_b200.scripts.stream_sweep --max-log2 30 --out gpurun_out/stream_sizes.csv

Runs `stream b200 native` (the counterpart of `sbench stream cuda-hip native`) for n = 2^20 ... 2^max
in float64 and float32 and writes one CSV row per (dtype, n, function): bandwidth in GB/s from the
minimum time over ntimes-1 rounds, like the reference's table.
"""

import argparse
import csv

from ..benchmarks_collection.stream import b200 as stream


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--min-log2", type=int, default=20)
    parser.add_argument("--max-log2", type=int, default=30)
    parser.add_argument("--ntimes", type=int, default=10)
    parser.add_argument("--dtypes", default="float64,float32")
    parser.add_argument("--out", default=None)
    args = parser.parse_args()
    rows = []
    for dtype in args.dtypes.split(","):
        for log2 in range(args.min_log2, args.max_log2 + 1):
            bench = stream.Native(array_size=1 << log2, ntimes=args.ntimes, dtype=dtype, verify=True)
            for result in bench.run():
                rows.append(dict(dtype=dtype, log2_n=log2, function=result["name"],
                                 gbs=result["bandwidth"] / 1e3, min_time_s=result["time"]))
            line = "  ".join(f"{r['function']} {r['gbs']:7.1f}" for r in rows[-4:])
            print(f"{dtype} 2^{log2}: {line} GB/s", flush=True)
    if args.out:
        with open(args.out, "w", newline="") as fh:
            writer = csv.DictWriter(fh, fieldnames=list(rows[0]))
            writer.writeheader()
            writer.writerows(rows)


if __name__ == "__main__":
    main()
