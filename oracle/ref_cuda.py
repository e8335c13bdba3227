"""Run the reference's own CUDA kernels, recompiled for sm_100 (oracle/_ref/cuda_*.so), on the GPU box.

BASELINE INFRASTRUCTURE ONLY ("the existing kernel to beat").
The libraries are built by oracle/build_ref.py from the reference's unmodified templates with the
block sizes of scripts/sbench_h100_collection.py; each exports the reference ABI
``int kernel(double* time, T* field0, ..., T* fieldN)`` with DEVICE pointers to the first interior
element (cuda_hip/templates/base.j2:68-74) and does ``dry_runs`` = 1 warm launch before the timed one.

    python -m oracle.ref_cuda --repeat 11 --out gpurun_out/reference_cuda.json
"""

import argparse
import ctypes
import json
import statistics

import numpy as np

from oracle import ref_cpu
from stencil_benchmarks_b200 import capi


def device_field(entry, fill):
    """Device buffer with the strides the kernel was rendered for; returns (buffer, interior pointer)."""
    dtype = np.dtype(entry["dtype"])
    shape = [d + 2 * h for d, h in zip(entry["domain"], entry["halo"])]
    total = sum((n - 1) * s for n, s in zip(shape, entry["strides"])) + 1
    interior = sum(s * h for s, h in zip(entry["strides"], entry["halo"]))
    buffer = capi.DeviceBuffer(total * dtype.itemsize + 512)
    first = buffer.ptr + (-(buffer.ptr + interior * dtype.itemsize) % 128)
    host = (np.random.default_rng(int(fill * 1000)).random(min(total, 1 << 22)) * 0.5 + fill).astype(dtype)
    done = 0
    while done < total:
        n = min(host.size, total - done)
        capi.memcpy_h2d(first + done * dtype.itemsize, host.ctypes.data, n * dtype.itemsize)
        done += n
    return buffer, first + interior * dtype.itemsize


def algorithmic_bytes(name, entry):
    nx, ny, nz = entry["domain"]
    size = np.dtype(entry["dtype"]).itemsize
    if "hdiff" in name:
        return (2 * nx * ny * nz + (nx + 4) * (ny + 4) * nz) * size
    if "vadv" in name:
        return 6 * nx * ny * nz * size
    return 2 * nx * ny * nz * size


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--repeat", type=int, default=11)
    parser.add_argument("--out", default=None)
    parser.add_argument("--only", default="")
    args = parser.parse_args()
    capi.require_device()
    results = []
    for name, entry in ref_cpu.manifest().items():
        if not name.startswith("cuda_") or args.only not in name:
            continue
        library = ctypes.CDLL(str(ref_cpu.REF / entry["libraries"]["sm_100"]))
        library.kernel.restype = ctypes.c_int
        fields = [device_field(entry, 0.1 + 0.05 * i) for i, _ in enumerate(entry["args"])]
        pointers = [ctypes.c_void_p(ptr) for _, ptr in fields]
        elapsed = ctypes.c_double()
        times = []
        failed = False
        for _ in range(args.repeat):
            if library.kernel(ctypes.byref(elapsed), *pointers) != 0:
                failed = True
                break
            times.append(elapsed.value)
        if failed:
            print(f"{name:36s} failed")
            continue
        median = statistics.median(times)
        nbytes = algorithmic_bytes(name, entry)
        row = dict(kernel=name, reference_class=entry["reference_class"], block_size=entry["kwargs"].get("block_size"),
                   median_s=median, min_s=min(times), gbs_algorithmic=nbytes / median / 1e9,
                   gbs_sbench=entry["data_size"] / median / 1e9)
        results.append(row)
        print(f"{name:36s} {median * 1e3:9.4f} ms  {row['gbs_algorithmic']:8.1f} GB/s algorithmic "
              f"({row['gbs_sbench']:8.1f} sbench)", flush=True)
        del fields
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(dict(device=capi.device_info(), results=results), fh, indent=1)


if __name__ == "__main__":
    main()
