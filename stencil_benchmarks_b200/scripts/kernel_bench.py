"""Device-resident kernel timings at the BASELINE.json sizes (tuning aid, not bench.py).

    python -m stencil_benchmarks_b200.scripts.kernel_bench [--what stream,basic,hdiff,vadv]
        [--repeat 20] [--out gpurun_out/kernels.json]

Fields live in device memory only (filled by a STREAM init + scale so they hold
finite numbers); every kernel is timed `repeat` times through the C ABI's own
event timing (one launch per call, like the reference's `kernel()`), median and
min are reported with the algorithmic bytes of SURVEY.md §8d.
"""

import argparse
import ctypes
import json
import os
import statistics

import numpy as np

from .. import capi

_vp = ctypes.c_void_p


def padded_geometry(domain, halo, itemsize, alignment=128):
    nx, ny, nz = domain
    hx, hy, hz = halo
    row = -(-((nx + 2 * hx) * itemsize) // alignment) * alignment // itemsize
    sy = row
    sz = sy * (ny + 2 * hy)
    total = sz * (nz + 2 * hz)
    interior = hx + hy * sy + hz * sz
    return sy, sz, total, interior


class Field:
    def __init__(self, total, interior, itemsize, fill):
        self.buffer = capi.DeviceBuffer(total * itemsize + 256)
        base = self.buffer.ptr
        self.first = base + (-(base + interior * itemsize) % 128)
        self.interior = self.first + interior * itemsize
        host = np.full(min(total, 1 << 22), fill, dtype="float64" if itemsize == 8 else "float32")
        # tile a host block over the device buffer (values in (0,1), varying)
        host += np.random.default_rng(int(fill * 1000)).random(host.size).astype(host.dtype) * 0.5
        done = 0
        while done < total:
            n = min(host.size, total - done)
            capi.memcpy_h2d(self.first + done * itemsize, host.ctypes.data, n * itemsize)
            done += n


def time_call(call, repeat):
    times = []
    t = ctypes.c_double()
    for _ in range(repeat):
        call(ctypes.byref(t))
        times.append(t.value)
    return statistics.median(times), min(times)


def time_loop(lib, enqueue, steps, warmup=5):
    """Mean time of `steps` back-to-back launches between two events (what bench.py measures)."""
    start, stop = _vp(), _vp()
    lib.sb200_event_create(ctypes.byref(start))
    lib.sb200_event_create(ctypes.byref(stop))
    for _ in range(warmup):
        enqueue()
    lib.sb200_synchronize(None)
    lib.sb200_event_record(start, None)
    for _ in range(steps):
        enqueue()
    lib.sb200_event_record(stop, None)
    lib.sb200_synchronize(None)
    elapsed = ctypes.c_double()
    lib.sb200_event_elapsed(start, stop, ctypes.byref(elapsed))
    lib.sb200_event_destroy(start)
    lib.sb200_event_destroy(stop)
    return elapsed.value / steps


def environment_settings(spec):
    """'A=1,B=2;A=0' -> [{'A': '1', 'B': '2'}, {'A': '0'}]"""
    settings = []
    for group in (spec.split(";") if spec else []):
        settings.append(dict(item.split("=", 1) for item in group.split(",") if item))
    return settings


class environment:
    def __init__(self, values):
        self.values = values

    def __enter__(self):
        self.saved = {k: os.environ.get(k) for k in self.values}
        os.environ.update(self.values)

    def __exit__(self, *exc):
        for key, value in self.saved.items():
            if value is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = value


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--hdiff-sweep", default=None,
                        help="semicolon-separated SB200_HDIFF_CFG values to time on the same fields")
    parser.add_argument("--vadv-sweep", default=None,
                        help="semicolon-separated NAME=VALUE[,NAME=VALUE] environment settings")
    parser.add_argument("--loop", type=int, default=0, help="also time a loop of this many launches")
    parser.add_argument("--what", default="stream,basic,hdiff,vadv")
    parser.add_argument("--repeat", type=int, default=20)
    parser.add_argument("--out", default=None)
    parser.add_argument("--stream-log2", type=int, default=28)
    parser.add_argument("--dtypes", default="float64,float32")
    args = parser.parse_args()
    lib = capi.library()
    capi.require_device()
    what = args.what.split(",")
    results = []

    def report(name, dtype, nbytes, med, mn, **extra):
        row = dict(kernel=name, dtype=dtype, bytes=nbytes, median_s=med, min_s=mn,
                   gbs_median=nbytes / med / 1e9, gbs_best=nbytes / mn / 1e9, **extra)
        results.append(row)
        print(f"{name:28s} {dtype:8s} {med * 1e3:9.4f} ms  {row['gbs_median']:8.1f} GB/s "
              f"(best {row['gbs_best']:8.1f})", flush=True)

    for dtype in args.dtypes.split(","):
        code = capi.dtype_code(dtype)
        size = np.dtype(dtype).itemsize
        if "stream" in what:
            n = 1 << args.stream_log2
            bufs = [capi.DeviceBuffer(n * size) for _ in range(3)]
            ptrs = [b.ptr for b in bufs]
            lib.sb200_stream_op(capi.STREAM_INIT, code, *ptrs, n, 3.0, 0, None, None)
            for op, name, factor in [(capi.STREAM_COPY, "copy", 2), (capi.STREAM_SCALE, "scale", 2),
                                     (capi.STREAM_ADD, "add", 3), (capi.STREAM_TRIAD, "triad", 3)]:
                lib.sb200_stream_op(capi.STREAM_INIT, code, *ptrs, n, 3.0, 0, None, None)
                med, mn = time_call(lambda t: lib.sb200_stream_op(op, code, *ptrs, n, 1e-3, 1, t, None),
                                    args.repeat)
                report(f"stream_{name}_2^{args.stream_log2}", dtype, factor * n * size, med, mn)
            del bufs
        if "basic" in what:
            domain, halo = (1024, 1024, 80), (3, 3, 3)
            sy, sz, total, interior = padded_geometry(domain, halo, size)
            inp, out = Field(total, interior, size, 0.25), Field(total, interior, size, 0.5)
            nbytes = 2 * int(np.prod(domain)) * size
            variants = [("copy", capi.BASIC_COPY, 0, 0)]
            variants += [(f"onesided_ax{a}", capi.BASIC_ONESIDED_AVG, a, 0) for a in range(3)]
            variants += [(f"symmetric_ax{a}", capi.BASIC_SYMMETRIC_AVG, a, 0) for a in range(3)]
            variants += [("laplacian_xy", capi.BASIC_LAPLACIAN, 0, 3), ("laplacian_xyz", capi.BASIC_LAPLACIAN, 0, 7)]
            for name, kind, axis, mask in variants:
                med, mn = time_call(
                    lambda t: lib.sb200_basic(kind, code, _vp(inp.interior), _vp(out.interior), *domain,
                                              1, sy, sz, axis, mask, 1, t, None), args.repeat)
                report("basic_" + name, dtype, nbytes, med, mn)
            del inp, out
        if "hdiff" in what:
            domain, halo = (2048, 2048, 80), (3, 3, 3)
            sy, sz, total, interior = padded_geometry(domain, halo, size)
            f = [Field(total, interior, size, v) for v in (0.2, 0.4, 0.6)]
            nx, ny, nz = domain
            nbytes = (2 * nx * ny * nz + (nx + 4) * (ny + 4) * nz) * size
            med, mn = time_call(
                lambda t: lib.sb200_hdiff(code, _vp(f[0].interior), _vp(f[1].interior), _vp(f[2].interior),
                                          *domain, 1, sy, sz, 1, t, None), args.repeat)
            report("hdiff_2048x2048x80", dtype, nbytes, med, mn)
            for cfg in (args.hdiff_sweep.split(";") if args.hdiff_sweep else []):
                os.environ["SB200_HDIFF_CFG"] = cfg
                med, mn = time_call(
                    lambda t: lib.sb200_hdiff(code, _vp(f[0].interior), _vp(f[1].interior), _vp(f[2].interior),
                                              *domain, 1, sy, sz, 1, t, None), args.repeat)
                extra = {}
                for steps in ([20, args.loop] if args.loop else []):
                    extra[f"loop{steps}_ms"] = 1e3 * time_loop(
                        lib, lambda: lib.sb200_hdiff(code, _vp(f[0].interior), _vp(f[1].interior),
                                                     _vp(f[2].interior), *domain, 1, sy, sz, 0, None, None), steps)
                report(f"hdiff cfg={cfg}", dtype, nbytes, med, mn, **extra)
                if extra:
                    print("    " + "  ".join(f"{k}={v:.4f}" for k, v in extra.items()), flush=True)
            os.environ.pop("SB200_HDIFF_CFG", None)
            del f
        if "vadv" in what:
            domain, halo = (1024, 1024, 160), (3, 3, 3)
            sy, sz, total, interior = padded_geometry(domain, halo, size)
            f = [Field(total, interior, size, v) for v in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7)]
            nbytes = 6 * int(np.prod(domain)) * size
            med, mn = time_call(
                lambda t: lib.sb200_vadv(code, *[_vp(x.interior) for x in f], None, *domain, 1, sy, sz,
                                         1, 0, capi.VADV_AUTO, 0, t, None), args.repeat)
            report("vadv_1024x1024x160", dtype, nbytes, med, mn,
                   sbench_gbs=10 * int(np.prod(domain)) * size / med / 1e9)
            for setting in environment_settings(args.vadv_sweep):
                with environment(setting):
                    med, mn = time_call(
                        lambda t: lib.sb200_vadv(code, *[_vp(x.interior) for x in f], None, *domain, 1, sy, sz,
                                                 1, 0, capi.VADV_AUTO, 0, t, None), args.repeat)
                report(f"vadv {setting}", dtype, nbytes, med, mn)
            del f
        if "vadv3" in what:
            # all_components: u, v, w in one sweep sharing wcon (16 fields' worth of traffic)
            domain, halo = (1024, 1024, 160), (3, 3, 3)
            sy, sz, total, interior = padded_geometry(domain, halo, size)
            comp = [[Field(total, interior, size, 0.1 + 0.1 * c + 0.02 * n) for n in range(4)] for c in range(3)]
            wcon = Field(total, interior, size, 0.7)
            nbytes = 16 * int(np.prod(domain)) * size

            def table(n, count=3):
                return (_vp * count)(*[_vp(comp[c][n].interior) for c in range(count)])

            three = (ctypes.c_int * 3)
            med, mn = time_call(
                lambda t: lib.sb200_vadv_components(
                    code, 3, table(0), table(1), table(2), table(3), three(1, 0, 0), three(0, 1, 0),
                    _vp(wcon.interior), None, None, *domain, 1, sy, sz, capi.VADV_AUTO, 0, t, None), args.repeat)
            report("vadv_uvw_merged_1024x1024x160", dtype, nbytes, med, mn)
            for setting in environment_settings(args.vadv_sweep):
                with environment(setting):
                    med, mn = time_call(
                        lambda t: lib.sb200_vadv_components(
                            code, 3, table(0), table(1), table(2), table(3), three(1, 0, 0), three(0, 1, 0),
                            _vp(wcon.interior), None, None, *domain, 1, sy, sz, capi.VADV_AUTO, 0, t, None),
                        args.repeat)
                report(f"vadv_uvw {setting}", dtype, nbytes, med, mn)
            total_med = 0.0
            for c, (i, j) in enumerate([(1, 0), (0, 1), (0, 0)]):
                med, mn = time_call(
                    lambda t: lib.sb200_vadv(code, *[_vp(x.interior) for x in comp[c]], _vp(wcon.interior), None,
                                             None, None, *domain, 1, sy, sz, i, j, capi.VADV_AUTO, 0, t, None),
                    args.repeat)
                total_med += med
            report("vadv_uvw_three_sweeps", dtype, nbytes, total_med, total_med)
            del comp, wcon
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(dict(device=capi.device_info(), results=results), fh, indent=1)


if __name__ == "__main__":
    main()
