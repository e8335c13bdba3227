"""Run the reference's own CUDA kernels, recompiled for sm_100 (oracle/_ref/cuda_*.so), on the GPU box.

BASELINE INFRASTRUCTURE ONLY ("the existing kernel to beat").
The libraries are built by oracle/build_ref.py from the reference's unmodified templates -- the
block sizes of scripts/sbench_h100_collection.py plus a grid around them -- and each exports the
reference ABI ``int kernel(double* time, T* field0, ..., T* fieldN)`` with DEVICE pointers to the
first interior element (cuda_hip/templates/base.j2:68-74); every call does ``dry_runs`` = 1 warm
launch before the timed one.

    python -m oracle.ref_cuda --repeat 11 --check --out gpurun_out/reference_cuda.json

``--check`` cross-checks what every variant wrote against the C oracle (oracle/oracle.c) on the
same seeded fields, at the reference's own tolerance (tools/validation.py:98-106).
"""

import argparse
import ctypes
import json
import statistics

import numpy as np

from oracle import ref_cpu
from stencil_benchmarks_b200 import capi


def algorithmic_bytes(name, entry):
    nx, ny, nz = entry["domain"]
    size = np.dtype(entry["dtype"]).itemsize
    if "hdiff" in name:
        return (2 * nx * ny * nz + (nx + 4) * (ny + 4) * nz) * size
    if "vadv" in name:
        return (16 if entry["kwargs"].get("all_components") else 6) * nx * ny * nz * size
    return 2 * nx * ny * nz * size


class FieldSet:
    """Seeded host fields with the strides a group of kernels was rendered for, and their device
    mirrors.  One set serves every variant of a (stencil, geometry) group."""

    def __init__(self, entry, seed=0):
        self.entry = entry
        self.names = list(entry["args"])
        self.halo = tuple(entry["halo"])
        self.host = []
        for index, _ in enumerate(self.names):
            field = ref_cpu.alloc_field(entry)
            rng = np.random.default_rng([seed, index])
            plane = rng.random(field.shape[:2])
            levels = field.shape[2]
            for k in range(levels):  # one random plane scaled per level: values in (0, 1), all levels differ
                np.multiply(plane, 0.5 + 0.5 * (k + 1) / levels, out=field[:, :, k])
            self.host.append(field)
        self.device = []
        for field in self.host:
            extent = self.extent(field)
            buffer = capi.DeviceBuffer(extent + 512)
            interior = self.interior_offset(field)
            first = buffer.ptr + (-(buffer.ptr + interior) % 128)
            self.device.append((buffer, first))
        self.upload()

    @staticmethod
    def extent(field):
        return sum((n - 1) * s for n, s in zip(field.shape, field.strides)) + field.itemsize

    def interior_offset(self, field):
        return sum(s * h for s, h in zip(field.strides, self.halo))

    def upload(self, only=None):
        for name, field, (_, first) in zip(self.names, self.host, self.device):
            if only is None or name in only:
                capi.memcpy_h2d(first, field.ctypes.data, self.extent(field))

    def download(self, name):
        index = self.names.index(name)
        field = self.host[index]
        result = ref_cpu.alloc_field(self.entry)
        capi.memcpy_d2h(result.ctypes.data, self.device[index][1], self.extent(field))
        return result

    def pointers(self):
        return [ctypes.c_void_p(first + self.interior_offset(field))
                for field, (_, first) in zip(self.host, self.device)]


def group_key(entry):
    return (tuple(entry["args"]), tuple(entry["domain"]), tuple(entry["halo"]), tuple(entry["strides"]),
            entry["dtype"])


def expected_outputs(name, fields):
    """{field name: expected array} after TWO sweeps (the warm launch and the timed one, base.j2:137-184)
    from the C oracle -- only vertical advection (in-out utensstage) is not idempotent."""
    from oracle import native

    entry, halo = fields.entry, fields.halo
    host = dict(zip(fields.names, fields.host))
    if "hdiff" in name:
        out = ref_cpu.alloc_field(entry)
        out[...] = 0
        native.hdiff(host["inp"], host["coeff"], out, halo)
        return {"out": out}
    if "vadv" in name:
        result = {}
        components = [("u", 1, 0)]
        if entry["kwargs"].get("all_components"):
            components += [("v", 0, 1), ("w", 0, 0)]
        ccol, dcol = ref_cpu.alloc_field(entry), ref_cpu.alloc_field(entry)
        for c, ishift, jshift in components:
            work = ref_cpu.alloc_field(entry)
            work[...] = host[c + "tensstage"]
            for _ in range(2):
                native.vadv(host[c + "stage"], host[c + "pos"], host[c + "tens"], work, host["wcon"], ccol, dcol,
                            halo, ishift, jshift)
            result[c + "tensstage"] = work
        return result
    out = ref_cpu.alloc_field(entry)
    out[...] = 0
    kwargs = entry["kwargs"]
    if "copy" in name:
        native.copy(host["inp"], out, halo)
    elif "avg" in name:
        native.average(host["inp"], out, halo, kwargs.get("axis", 0), symmetric=False)
    else:
        native.laplacian(host["inp"], out, halo, (True, True, False))
    return {"out": out}


def run(only="", repeat=11, check=False, verbose=True):
    """Time (and optionally cross-check) every oracle/_ref CUDA kernel whose name contains `only`."""
    capi.require_device()
    entries = {n: e for n, e in ref_cpu.manifest().items() if n.startswith("cuda_") and only in n}
    groups = {}
    for name, entry in entries.items():
        groups.setdefault(group_key(entry), []).append(name)
    results = []
    for names in groups.values():
        fields = FieldSet(entries[names[0]])
        inner = tuple(slice(h, h + d) for d, h in zip(fields.entry["domain"], fields.halo))
        expected = {}
        for name in names:
            entry = entries[name]
            library = ctypes.CDLL(str(ref_cpu.REF / entry["libraries"]["sm_100"]))
            library.kernel.restype = ctypes.c_int
            pointers = fields.pointers()
            elapsed = ctypes.c_double()
            times = []
            for _ in range(repeat):
                if library.kernel(ctypes.byref(elapsed), *pointers) != 0:
                    times = []
                    break
                times.append(elapsed.value)
            if not times:
                if verbose:
                    print(f"{name:40s} failed", flush=True)
                results.append(dict(kernel=name, failed=True))
                continue
            median = statistics.median(times)
            nbytes = algorithmic_bytes(name, entry)
            row = dict(kernel=name, reference_class=entry["reference_class"],
                       block_size=entry["kwargs"].get("block_size"),
                       unroll_factor=entry["kwargs"].get("unroll_factor"), median_s=median, min_s=min(times),
                       gbs_algorithmic=nbytes / median / 1e9, gbs_sbench=entry["data_size"] / median / 1e9)
            if check:
                written = [n for n in fields.names if n == "out" or n.endswith("tensstage")]
                if "vadv" in name and not entry["kwargs"].get("all_components"):
                    written = ["utensstage"]
                fields.upload(only=written)
                if library.kernel(ctypes.byref(elapsed), *pointers) != 0:
                    row.update(check="failed to run")
                else:
                    # the variants of a group compute the same thing, except the basic stencils
                    kind = name if "basic" in name else "group"
                    if kind not in expected:
                        expected[kind] = expected_outputs(name, fields)
                    worst, ok = 0.0, True
                    for field_name, want in expected[kind].items():
                        got = fields.download(field_name)
                        worst = max(worst, float(np.max(np.abs(got[inner] - want[inner]))))
                        ok = ok and bool(np.allclose(got[inner], want[inner], rtol=1e-5, atol=1e-8))
                    row.update(check_ok=ok, check_max_abs_err=worst)
                fields.upload(only=written)
            results.append(row)
            if verbose:
                note = ""
                if check:
                    note = f"  check {'ok' if row.get('check_ok') else 'FAILED'} (max abs err {row.get('check_max_abs_err', float('nan')):.2e})"
                print(f"{name:40s} {median * 1e3:9.4f} ms  {row['gbs_algorithmic']:8.1f} GB/s algorithmic "
                      f"({row['gbs_sbench']:8.1f} sbench){note}", flush=True)
        del fields
    return results


def best(results, prefix):
    """Fastest variant whose name starts with `prefix` (and, if checked, passed the check)."""
    rows = [r for r in results if r["kernel"].startswith(prefix) and not r.get("failed")
            and r.get("check_ok", True)]
    return min(rows, key=lambda r: r["median_s"]) if rows else None


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--repeat", type=int, default=11)
    parser.add_argument("--out", default=None)
    parser.add_argument("--only", default="")
    parser.add_argument("--check", action="store_true")
    args = parser.parse_args()
    results = run(args.only, args.repeat, args.check)
    summary = {}
    for prefix in ("cuda_hdiff", "cuda_vadv_localmemmerged_uvw", "cuda_vadv", "cuda_basic_copy", "cuda_basic_lap_ij"):
        rows = [r for r in results if not (prefix == "cuda_vadv" and "uvw" in r["kernel"])]
        row = best(rows, prefix)
        if row is not None:
            summary[prefix] = dict(kernel=row["kernel"], ms=row["median_s"] * 1e3, gbs=row["gbs_algorithmic"])
            print(f"best {prefix:30s} {row['kernel']:40s} {row['median_s'] * 1e3:.4f} ms")
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(dict(device=capi.device_info(), best=summary, results=results), fh, indent=1)


if __name__ == "__main__":
    main()
