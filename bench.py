#!/usr/bin/env python
"""Headline benchmark: effective HBM GB/s of the B200 stencil backend (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hdiff|vadv|triad]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own OpenMP CPU kernels

Workload (BASELINE.json): horizontal diffusion 2048x2048x80 float64 per GPU (weak scaling:
the global domain is 2048 x 2048*N x 80, J-partitioned; `--scaling strong` splits the single
2048x2048x80 domain instead).  The width-2 halo rows a sweep needs from its neighbours are read
by the sweep itself from the neighbour GPU's memory (CUDA IPC over NVLink, `--exchange peer`,
default) or exchanged as width-3 halos with NCCL send/recv overlapped with the interior rows
(`--exchange nccl`).  One "step" = one sweep.
`value` counts the ALGORITHMIC bytes of SURVEY.md §8d, (2*N + (nx+4)(ny+4)nz)*8 per GPU,
not the larger sbench figure.  Fields (8.7 GB per GPU) are far larger than the 126 MB L2,
so no L2 flush is needed between steps.

Keys beyond the base contract: `roofline` (dominant kernel vs the measured HBM peak),
`cpu_baseline` (reference OpenMP kernels from oracle/_ref on this box's cores, bounded
sample), `e2e` (same metric through the plugin's run(): pinned host fields, H2D of the
inputs and D2H of the output inside the timed region), `also` (triad and vadv device-timed).
"""

import argparse
import ctypes
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

ROOT = pathlib.Path(__file__).parent.resolve()
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    "hdiff": dict(domain=(2048, 2048, 80), dtype="float64", halo=(3, 3, 3),
                  reference_kernels=["hdiff_otfvec_2048x2048x80_f64", "hdiff_otf_2048x2048x80_f64",
                                     "hdiff_minimummem_2048x2048x80_f64"]),
    "vadv": dict(domain=(1024, 1024, 160), dtype="float64", halo=(3, 3, 3),
                 reference_kernels=["vadv_kmiddlevec_1024x1024x160_f64",
                                    "vadv_kinnermostvec_1024x1024x160_f64"]),
}
TRIAD_N = 1 << 30  # BASELINE.json configs[1]: STREAM up to 2^30 float64 elements per array
METRIC = {"hdiff": "horizontal-diffusion effective HBM bandwidth",
          "vadv": "vertical-advection effective HBM bandwidth",
          "triad": "STREAM triad effective HBM bandwidth"}
FALLBACK_PEAK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md


def algorithmic_bytes(workload, domain, itemsize=8):
    if workload == "triad":
        return 3 * TRIAD_N * itemsize
    nx, ny, nz = domain
    if workload == "hdiff":
        return (2 * nx * ny * nz + (nx + 4) * (ny + 4) * nz) * itemsize
    return 6 * nx * ny * nz * itemsize


def measured_peak():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except (KeyError, ValueError):
            pass
    return FALLBACK_PEAK_GBS, "fallback (B200_PROFILING.md)"


def profiled_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu summary, or None."""
    path = ROOT / "profiles" / "traffic.json"
    if path.exists():
        try:
            return json.loads(path.read_text()).get(workload)
        except ValueError:
            pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._loop, daemon=True)

    def _sample(self):
        try:
            out = subprocess.run(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            parts = [p.strip() for p in out.strip().split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)
        except (OSError, subprocess.SubprocessError):
            pass

    def _loop(self):
        while not self._stop.is_set():
            self._sample()
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(self.samples[0][1]), "samples": len(self.samples),
                "reasons": reasons}


# ------------------------------------------------------------------------------------------
# reference arm: the reference's OpenMP kernels (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------
def time_reference_triad(steps, warmup, budget_s=None):
    """CPU baseline of kind "port": the C restatement of the STREAM kernels (oracle/oracle.c) with
    OpenMP on all host cores.  (The reference's CPU STREAM, stream/mc_calpin.py, is a separate
    benchmark family that oracle/_ref does not build.)  Bounded sample: 2^28 elements per array."""
    import numpy as np

    from oracle import native, ref_cpu

    threads = ref_cpu.use_all_cores()
    n = 1 << 28
    a, b, c = np.zeros(n), np.full(n, 2.0), np.full(n, 0.5)
    for _ in range(max(warmup, 1)):
        native.stream_triad(a, b, c)
    times = []
    start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        native.stream_triad(a, b, c)
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - start > budget_s:
            break
    mean = sum(times) / len(times)
    return dict(name="oracle_stream_triad_f64 (OpenMP port)", isa="native gcc -O2", mean_s=mean,
                sweeps=len(times), min_s=min(times), threads=threads, bytes=3 * n * 8, kind="port",
                sample=f"{len(times)} triads over 2^28 float64 elements per array")


def time_reference(workload, steps, warmup, budget_s=None):
    from oracle import ref_cpu

    if workload == "triad":
        return time_reference_triad(steps, warmup, budget_s)

    cfg = WORKLOADS[workload]
    if not ref_cpu.available():
        raise RuntimeError("oracle/_ref is not built (run oracle/build_ref.py in the dev container)")
    ref_cpu.use_all_cores()
    best = None
    for name in cfg["reference_kernels"]:
        kernel = ref_cpu.Kernel(name)
        fields = kernel.fields(seed=0, fast=True)
        for _ in range(max(warmup, 1)):
            kernel(fields)
        times = []
        start = time.perf_counter()
        for _ in range(steps):
            times.append(kernel(fields))
            if budget_s is not None and time.perf_counter() - start > budget_s:
                break
        result = dict(name=name, isa=kernel.isa, mean_s=sum(times) / len(times), sweeps=len(times),
                      min_s=min(times))
        if best is None or result["mean_s"] < best["mean_s"]:
            best = result
        del fields, kernel
    best["threads"] = ref_cpu.threads()
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    best = time_reference(args.workload, args.steps, args.warmup)
    if args.workload == "triad":
        value = best["bytes"] / best["mean_s"] / 1e9
        sample = best["sample"]
    else:
        cfg = WORKLOADS[args.workload]
        nbytes = algorithmic_bytes(args.workload, cfg["domain"])
        value = nbytes / best["mean_s"] / 1e9
        sample = (f"{best['sweeps']} full sweeps of {'x'.join(map(str, cfg['domain']))} float64, reference "
                  f"OpenMP kernel {best['name']} ({best['isa']}), best of {len(cfg['reference_kernels'])} variants")
    line = {
        "impl": "reference",
        "metric": METRIC[args.workload], "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": best["sweeps"], "warmup": max(args.warmup, 1), "ms_per_step": best["mean_s"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.workload, 1),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": best["threads"],
                         "kind": best.get("kind", "reference"), "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


EXCHANGE_TEXT = {
    "peer": "fused into the sweep: halo rows read by TMA from the neighbour GPU's HBM (CUDA IPC peer "
            "memory over NVLink), every sweep, one kernel launch per sweep",
    "nccl": "NCCL send/recv, width 3, every sweep, on a high-priority stream overlapped with the interior kernel",
}


def local_domain(workload, n_gpus, rank, scaling):
    """Domain one rank sweeps: the full BASELINE domain (weak) or its J slab of it (strong)."""
    nx, ny, nz = WORKLOADS[workload]["domain"]
    if scaling == "strong" and n_gpus > 1:
        from stencil_benchmarks_b200 import distributed

        ny = distributed.split_rows(ny, n_gpus)[rank][1]
    return nx, ny, nz


def workload_config(workload, n_gpus, exchange=None, scaling="weak"):
    if workload == "triad":
        return {
            "workload": f"STREAM triad a = b + 3c, {TRIAD_N} float64 elements per array per GPU",
            "partition": "independent arrays per GPU, no communication",
            "bytes_per_step_per_gpu": algorithmic_bytes("triad", None),
            "l2": "arrays (3 x 8.6 GB per GPU) exceed the 126 MB L2; no flush between steps",
        }
    cfg = WORKLOADS[workload]
    nx, ny, nz = cfg["domain"]
    strong = scaling == "strong" and n_gpus > 1
    per_gpu = local_domain(workload, n_gpus, 0, scaling)
    resident_gb = (3 if workload == "hdiff" else 8) * (nx + 6) * (per_gpu[1] + 6) * (nz + 6) * 8 / 1e9
    return {
        "workload": (f"{workload} {nx}x{ny}x{nz} float64 " + ("global" if strong else "per GPU")
                     + ", halo 3, alignment 128"),
        "global_domain": [nx, ny if strong else ny * n_gpus, nz],
        "per_gpu_domain": list(per_gpu),
        "partition": "J slabs, one per GPU" if n_gpus > 1 else "single GPU",
        "halo_exchange": (EXCHANGE_TEXT.get(exchange, "none") if workload == "hdiff" and n_gpus > 1 else "none"),
        "bytes_per_step_per_gpu": algorithmic_bytes(workload, per_gpu),
        "l2": f"fields ({resident_gb:.1f} GB per GPU) exceed the 126 MB L2; no flush between steps",
    }


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch

    from stencil_benchmarks_b200 import capi, distributed
    from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import (
        horizontal_diffusion,
        vertical_advection,
    )

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        # the exchange must not queue behind the interior sweep's CTAs: NCCL's own stream and the
        # communication stream below get high priority
        options = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=options)
    else:
        dist = None

    lib = capi.library()
    cfg = WORKLOADS[args.workload]
    domain = local_domain(args.workload, world, rank, args.scaling)
    nx, ny, nz = domain
    cls = horizontal_diffusion.Fused if args.workload == "hdiff" else vertical_advection.Thomas
    bench = cls(domain=domain, halo=cfg["halo"], dtype=cfg["dtype"], verify=False,
                device=local_rank, seed=100 + rank, dry_runs=0)
    data = bench.data()
    mirrors = bench._device_fields(data)
    bench.upload(data, mirrors)
    pointers = {name: bench.interior_ptr(mirrors[name][1], host) for name, host in zip(bench.args, data)}
    geometry = bench.geometry()
    sy, sz = geometry[4], geometry[5]
    code = capi.dtype_code(cfg["dtype"])
    raw = lib.raw

    main_stream = torch.cuda.current_stream()
    comm_stream = torch.cuda.Stream(priority=-1)
    exchange = None
    peers = None
    if args.workload == "hdiff" and world > 1 and args.exchange == "peer":
        # fused exchange: neighbours' inp slabs mapped through CUDA IPC, halo rows read by TMA.
        # If any rank cannot map its neighbours (IPC disabled in the container), every rank
        # switches to the NCCL exchange -- still a GPU path -- and the line says so.
        try:
            peers = distributed.PeerSlabs(dist, rank, world, mirrors["inp"][0].ptr, pointers["inp"].value,
                                          ny, sz)
            mapped = 1
        except Exception as error:  # noqa: BLE001 - reported, not swallowed
            print(f"rank {rank}: peer mapping failed ({error}); using the NCCL exchange", file=sys.stderr)
            mapped = 0
        flag = torch.tensor([mapped], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if peers is not None:
                peers.close()
                peers = None
            args.exchange = "nccl"
    if args.workload == "hdiff" and world > 1 and peers is None:
        exchange = distributed.cuda_halo_exchange(rank, world, cfg["dtype"], nx, ny, nz, cfg["halo"][0],
                                                  sy, sz, width=cfg["halo"][1])
        (lo, hi), strips = distributed.interior_and_boundary_rows(
            ny, 2, exchange.lower is not None, exchange.upper is not None)

    def vp(value):
        return ctypes.c_void_p(value)

    def hdiff_rows(j0, j1, stream):
        offset = j0 * sy * 8
        raw.sb200_hdiff(code, vp(pointers["inp"].value + offset), vp(pointers["coeff"].value + offset),
                        vp(pointers["out"].value + offset), nx, j1 - j0, nz, 1, sy, sz, 0, None,
                        vp(stream.cuda_stream))

    def step():
        if args.workload == "vadv":
            bench.launch(pointers, 0, None, main_stream.cuda_stream)
        elif peers is not None:
            raw.sb200_hdiff_peer(code, pointers["inp"], pointers["coeff"], pointers["out"],
                                 vp(peers.lower), peers.ny_lower, peers.sz_lower,
                                 vp(peers.upper), peers.ny_upper, peers.sz_upper,
                                 nx, ny, nz, 1, sy, sz, 0, None, vp(main_stream.cuda_stream))
        elif exchange is None:
            hdiff_rows(0, ny, main_stream)
        else:
            # halo exchange on the communication stream, interior rows meanwhile
            comm_stream.wait_stream(main_stream)
            with torch.cuda.stream(comm_stream):
                requests = exchange.start(pointers["inp"].value)
            hdiff_rows(lo, hi, main_stream)
            with torch.cuda.stream(comm_stream):
                exchange.finish(pointers["inp"].value, requests)
                for j0, j1 in strips:
                    hdiff_rows(j0, j1, comm_stream)
            main_stream.wait_stream(comm_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()  # every rank has uploaded its fields before a neighbour reads them
    for _ in range(args.warmup):
        step()
    barrier()
    launches_before = capi.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        start.record(main_stream)
        for _ in range(args.steps):
            step()
        stop.record(main_stream)
        barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = capi.launch_count() - launches_before
    if dist is not None:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    # bytes of one rank (rank 0's slab is the largest); whole job = every rank's slab
    nbytes = algorithmic_bytes(args.workload, local_domain(args.workload, world, 0, args.scaling))
    job_bytes = sum(algorithmic_bytes(args.workload, local_domain(args.workload, world, r, args.scaling))
                    for r in range(world))
    value = job_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the plugin API: H2D inputs + kernel + D2H outputs per step ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    bench.chunks = args.e2e_chunks  # slab-pipelined upload / sweep / download on three streams
    h2d, d2h = bench.transfer_bytes()
    bench.run()  # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        bench.run()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = job_bytes / e2e_s / 1e9

    peak, peak_source = measured_peak()
    achieved = nbytes / (ms_per_step * 1e-3) / 1e9
    line = {
        "metric": METRIC[args.workload], "value": value, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.workload, world, args.exchange, args.scaling),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": profiled_traffic(args.workload),
                     "peak_source": peak_source,
                     "kernel": "hdiff_tma_kernel<double>" if args.workload == "hdiff" else "vadv kernel",
                     "note": "per GPU; achieved = algorithmic bytes / mean step time (CUDA events)"},
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": d2h * world, "steps": e2e_steps, "ms_per_step": e2e_s * 1e3,
                "how": f"plugin run() with chunks={args.e2e_chunks}: pinned host fields, per step H2D of the "
                       "fields the sweep reads, the sweep, D2H of the field it writes, slab-pipelined"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "pct_of_nominal_8TBs": achieved / 8000.0,
    }

    if rank == 0 and world == 1 and not args.no_extras:
        del bench, mirrors, data
        line["also"] = extra_kernels(lib, capi, args)
        if not args.no_cpu_baseline:
            try:
                best = time_reference(args.workload, steps=5, warmup=1, budget_s=20.0)
                line["cpu_baseline"] = {
                    "value": nbytes / best["mean_s"] / 1e9, "unit": "GB/s", "cores": best["threads"],
                    "kind": "reference",
                    "sample": f"{best['sweeps']} full sweeps, reference OpenMP kernel {best['name']} "
                              f"({best['isa']})"}
            except Exception as error:  # the baseline must not lose the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": os.cpu_count(),
                                        "kind": "reference", "sample": f"failed: {error}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        if peers is not None:
            peers.close()
        dist.destroy_process_group()
    return 0


def run_triad(args):
    """STREAM triad, one independent set of arrays per GPU (no communication)."""
    import torch

    from stencil_benchmarks_b200 import capi
    from stencil_benchmarks_b200.benchmarks_collection.stream import b200 as stream

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = capi.library()
    raw = lib.raw
    n = TRIAD_N
    buffers = [capi.DeviceBuffer(8 * n) for _ in range(3)]
    ptrs = [ctypes.c_void_p(b.ptr) for b in buffers]
    lib.sb200_stream_op(capi.STREAM_INIT, capi.F64, *ptrs, n, 3.0, 0, None, None)
    stream_handle = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        raw.sb200_stream_op(capi.STREAM_TRIAD, capi.F64, *ptrs, n, 3.0, 0, None, stream_handle)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches_before = capi.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        start.record()
        for _ in range(args.steps):
            step()
        stop.record()
        barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = capi.launch_count() - launches_before
    if dist is not None:
        t = torch.tensor([elapsed_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    nbytes = algorithmic_bytes("triad", None)
    achieved = nbytes / (ms_per_step * 1e-3) / 1e9
    del buffers
    # end to end = the plugin call: the reference's STREAM interface has no host data
    # (stream/cuda_hip.py:121-144 runs init + kernels + verification on the device)
    results = stream.Native(array_size=n, ntimes=5, dtype="float64", device=local_rank).run()
    triad = next(r for r in results if r["name"] == "triad")
    e2e_value = triad["bandwidth"] / 1e3
    if dist is not None:
        t = torch.tensor([e2e_value], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        e2e_value = float(t.item())
    peak, peak_source = measured_peak()
    line = {
        "metric": METRIC["triad"], "value": world * achieved, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config("triad", world),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": profiled_traffic("triad"),
                     "peak_source": peak_source, "kernel": "stream_kernel<double, TRIAD>",
                     "note": "per GPU; the denominator is a torch copy, a 2:1 read:write kernel reads above 1"},
        "e2e": {"value": world * e2e_value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "how": "plugin call stream.b200.Native.run() (McCalpin table, min time over 4 rounds); "
                       "the reference's STREAM interface keeps all data on the device"},
        "gpu_launches": int(launches), "clocks": clocks.summary(),
        "pct_of_nominal_8TBs": achieved / 8000.0,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        best = time_reference_triad(steps=20, warmup=1, budget_s=15.0)
        line["cpu_baseline"] = {"value": best["bytes"] / best["mean_s"] / 1e9, "unit": "GB/s",
                                "cores": best["threads"], "kind": "port", "sample": best["sample"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def extra_kernels(lib, capi, args):
    """Device-timed STREAM triad (2^28 f64; 2^30 with --full-stream) and the other stencil."""
    import numpy as np

    result = {}
    n = 1 << (30 if args.full_stream else 28)
    buffers = [capi.DeviceBuffer(8 * n) for _ in range(3)]
    ptrs = [b.ptr for b in buffers]
    lib.sb200_stream_op(capi.STREAM_INIT, capi.F64, *ptrs, n, 3.0, 0, None, None)
    t = ctypes.c_double()
    times = []
    for _ in range(10):
        lib.sb200_stream_op(capi.STREAM_TRIAD, capi.F64, *ptrs, n, 1e-3, 0, ctypes.byref(t), None)
        times.append(t.value)
    result["stream_triad_f64"] = {"n": n, "gbs": 3 * 8 * n / statistics.median(times[1:]) / 1e9}
    del buffers
    return result


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=200)
    parser.add_argument("--warmup", type=int, default=5)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--workload", default="hdiff", choices=sorted(WORKLOADS) + ["triad"])
    parser.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                        help="hdiff halo exchange at N > 1: fused into the sweep over peer memory, "
                             "or NCCL send/recv on a second stream")
    parser.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                        help="weak: the BASELINE domain per GPU (default, what the driver measures); "
                             "strong: the BASELINE domain split into J slabs over the GPUs")
    parser.add_argument("--e2e-steps", type=int, default=3)
    parser.add_argument("--e2e-chunks", type=int, default=8)
    parser.add_argument("--no-extras", action="store_true")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--full-stream", action="store_true")
    args = parser.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "triad":
        return run_triad(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
