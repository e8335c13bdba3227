#!/usr/bin/env python
"""Headline benchmark: effective HBM GB/s of the B200 stencil backend (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hdiff|vadv|triad]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own OpenMP CPU kernels

Workload (BASELINE.json): horizontal diffusion 2048x2048x80 float64 per GPU (weak scaling:
the global domain is 2048 x 2048*N x 80, J-partitioned; `--scaling strong` splits the single
2048x2048x80 domain instead).  Every rank builds its slab of ONE global synthetic field whose
rows depend on (seed, global row) only, so the neighbours' rows are known everywhere.  The
width-2 halo rows a sweep needs from its neighbours are read by the sweep itself from the
neighbour GPU's memory (CUDA IPC over NVLink, `--exchange peer`, default; the local halo rows are
poisoned with NaN) or exchanged as width-3 halos with NCCL send/recv overlapped with the interior
rows (`--exchange nccl`).  One "step" = one sweep of every slab.  `--iterate` makes it a time loop:
`inp` and `out` swap roles every step and the ranks order their sweeps through step flags in peer
memory (no host involvement).

`value` counts the ALGORITHMIC bytes of SURVEY.md §8d, (2*N + (nx+4)(ny+4)nz)*8 per GPU, not the
larger sbench figure.  Fields (8.7 GB per GPU) are far larger than the 126 MB L2, so no L2 flush
is needed between steps.

Keys beyond the base contract:
  roofline          dominant kernel against the measured HBM peak
  exchange_parity   after the timed loop, the edge row blocks of every slab (the rows whose inputs
                    come from the neighbours) against the C oracle on the global field
  e2e               the same metric through the plugin's run(): pinned host fields, H2D of the inputs
                    and D2H of the output inside the timed region; at N > 1 through the partitioned
                    run (edge rows first, neighbours ordered, fused halo reads); with the PCIe copy
                    rates measured in the same process beside it
  also              device-timed numbers of the other BASELINE.json configs at the same N: STREAM
                    copy/scale/add/triad at 2^30 float64, basic copy / Laplacian (float32 + float64),
                    vertical advection 1024x1024x160 (u, and u/v/w in one sweep), each with its
                    fraction of the measured and of the nominal peak
  cpu_baseline      the reference's OpenMP kernels (oracle/_ref) on this box's cores (N = 1 only);
                    `also`: the other BASELINE configs on the host cores (hdiff 128x128x80, basic copy /
                    Laplacian 1024x1024x80, vadv 1024x1024x160), best OpenMP variant each
  reference_gpu     the reference's own CUDA kernels recompiled for sm_100 (best of the block-size
                    sweep in oracle/_ref), same box, same run (N = 1 only)
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
PROCESS_START = time.perf_counter()

WORKLOADS = {
    "hdiff": dict(domain=(2048, 2048, 80), dtype="float64", halo=(3, 3, 3),
                  reference_kernels=["hdiff_otfvec_2048x2048x80_f64", "hdiff_otf_2048x2048x80_f64",
                                     "hdiff_minimummem_2048x2048x80_f64", "hdiff_rolling_2048x2048x80_f64"]),
    "vadv": dict(domain=(1024, 1024, 160), dtype="float64", halo=(3, 3, 3),
                 reference_kernels=["vadv_kmiddlevec_1024x1024x160_f64",
                                    "vadv_kinnermostvec_1024x1024x160_f64"]),
}
TRIAD_N = 1 << 30  # BASELINE.json configs[1]: STREAM up to 2^30 float64 elements per array
METRIC = {"hdiff": "horizontal-diffusion effective HBM bandwidth",
          "vadv": "vertical-advection effective HBM bandwidth",
          "triad": "STREAM triad effective HBM bandwidth"}
FALLBACK_PEAK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md
NOMINAL_PEAK_GBS = 8000.0   # north star: "~8 TB/s HBM3e"
FIELD_SEEDS = {"inp": 1, "coeff": 2}
EDGE_ROWS = 8  # rows per edge block checked against the oracle after the timed loop
# Time loop: the explicit scheme is stable for coeff <= 1/32 (the limited 4th-order operator has
# eigenvalues up to 64); with the reference's U[0,1) coefficients every sweep amplifies rounding
# differences about fourfold and a 25-sweep parity check is meaningless.  The loop therefore runs
# with coeff = U[0,1) / 40 -- a time step a model would take.
TIME_LOOP_COEFF_SCALE = 0.025
TIME_LOOP_MAX_CHECKED_SWEEPS = 32


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
    """SM clock and throttle reasons while the timed region runs: NVML in-process (about 20 kHz
    possible, sampled every millisecond), nvidia-smi every 100 ms where NVML cannot be loaded."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device, uuid=None):
        self.device = device
        self.samples = []   # (sm MHz, reasons bitmask)
        self.max_mhz = None
        self.source = "nvml"
        self._handle = None
        self._stop = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            handle = None
            if uuid:
                try:
                    handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
                except pynvml.NVMLError:
                    handle = None
            self._handle = handle or pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001 - any NVML problem: fall back to nvidia-smi
            self._handle = None
            self.source = "nvidia-smi"
        self._thread = threading.Thread(target=self._loop, daemon=True)

    def _sample_nvml(self):
        n = self._nvml
        try:
            self.samples.append((float(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)),
                                 int(n.nvmlDeviceGetCurrentClocksEventReasons(self._handle))))
        except n.NVMLError:
            pass

    def _sample_smi(self):
        try:
            out = subprocess.run(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            parts = [p.strip() for p in out.strip().split(",")]
            if len(parts) >= 7 and parts[0].replace(".", "").isdigit():
                mask = 0
                for bit, column in ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6)):
                    if parts[column] == "Active":
                        mask |= bit
                self.samples.append((float(parts[0]), mask))
                self.max_mhz = float(parts[1])
        except (OSError, ValueError, subprocess.SubprocessError):
            pass

    def _loop(self):
        while not self._stop.is_set():
            if self._handle is not None:
                self._sample_nvml()
                self._stop.wait(0.001)
            else:
                self._sample_smi()
                self._stop.wait(0.1)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "samples": 0, "reasons": ["unavailable"],
                    "source": self.source}
        mask = 0
        for _, reasons in self.samples:
            mask |= reasons
        clocks = [mhz for mhz, _ in self.samples]
        return {"sm_mhz": statistics.median(clocks), "sm_min_mhz": min(clocks), "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "source": self.source,
                "reasons": [name for bit, name in sorted(self.REASONS.items()) if mask & bit]}


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


def time_oracle_port(workload, steps, warmup, budget_s=None):
    """CPU baseline of kind "port": the C restatement of the reference's algorithm (oracle/oracle.c,
    OpenMP over k planes / columns) on all host cores, full sweeps of the workload.  Used where
    oracle/_ref -- the reference's own compiled kernels -- is not there."""
    import numpy as np

    from oracle import native, ref_cpu

    cfg = WORKLOADS[workload]
    threads = ref_cpu.use_all_cores()
    halo = cfg["halo"]
    shape = tuple(d + 2 * h for d, h in zip(cfg["domain"], halo))
    rng = np.random.default_rng(0)
    count = 3 if workload == "hdiff" else 7
    fields = []
    for index in range(count):
        field = np.empty(shape, dtype=cfg["dtype"], order="F")
        plane = rng.random(shape[:2])
        for k in range(shape[2]):
            np.multiply(plane, 0.5 + 0.5 * (k + 1) / shape[2], out=field[:, :, k])
        fields.append(field)

    def sweep():
        t0 = time.perf_counter()
        if workload == "hdiff":
            native.hdiff(fields[0], fields[1], fields[2], halo)
        else:
            native.vadv(*fields, halo)
        return time.perf_counter() - t0

    for _ in range(max(warmup, 1)):
        sweep()
    times = []
    start = time.perf_counter()
    for _ in range(steps):
        times.append(sweep())
        if budget_s is not None and time.perf_counter() - start > budget_s:
            break
    mean = sum(times) / len(times)
    return dict(name=f"oracle_{workload}_f64 (OpenMP port, oracle/oracle.c)", isa="port: gcc -O2 -fopenmp",
                mean_s=mean, sweeps=len(times), min_s=min(times), threads=threads, kind="port",
                tried=[f"port {mean * 1e3:.1f} ms"])


def time_reference(workload, steps, warmup, budget_s=None):
    """Best of the reference's OpenMP variants for the workload, each compiled on this machine with
    the reference's own flags (-march=native included) where g++ is present."""
    from oracle import ref_cpu

    if workload == "triad":
        return time_reference_triad(steps, warmup, budget_s)

    cfg = WORKLOADS[workload]
    if not ref_cpu.available():
        # oracle/_ref exists only where the reference tree was present at build time
        return time_oracle_port(workload, steps, warmup, budget_s)
    ref_cpu.use_all_cores()
    names = [n for n in cfg["reference_kernels"] if n in ref_cpu.manifest()]
    best = None
    tried = []
    for name in names:
        kernel = ref_cpu.Kernel(name)
        fields = kernel.fields(seed=0, fast=True)
        for _ in range(max(warmup, 1)):
            kernel(fields)
        times = []
        start = time.perf_counter()
        for _ in range(steps):
            times.append(kernel(fields))
            if budget_s is not None and time.perf_counter() - start > budget_s / len(names):
                break
        result = dict(name=name, isa=kernel.isa, mean_s=sum(times) / len(times), sweeps=len(times),
                      min_s=min(times))
        tried.append(f"{name.split('_')[1]} {result['mean_s'] * 1e3:.1f} ms")
        if best is None or result["mean_s"] < best["mean_s"]:
            best = result
        del fields, kernel
    best["threads"] = ref_cpu.threads()
    best["tried"] = tried
    return best


# the other BASELINE.json configs on the host cores ("CPU OpenMP numbers alongside"):
# (label, kernels of oracle/_ref to try, workload whose byte formula applies, domain)
CPU_OTHER_CONFIGS = [
    ("hdiff_128x128x80_f64 (BASELINE configs[0], cache resident)",
     ["hdiff_otf_128x128x80_f64", "hdiff_otfvec_128x128x80_f64", "hdiff_rolling_128x128x80_f64"],
     "hdiff", (128, 128, 80)),
    ("basic_copy_1024x1024x80_f64", ["copy_1dvec_1024x1024x80_f64"], "basic", (1024, 1024, 80)),
    ("basic_laplacian_ij_1024x1024x80_f64", ["laplacian_3dvec_1024x1024x80_f64"], "basic", (1024, 1024, 80)),
    ("vadv_1024x1024x160_f64", ["vadv_kmiddlevec_1024x1024x160_f64", "vadv_kinnermostvec_1024x1024x160_f64"],
     "vadv", (1024, 1024, 160)),
    ("hdiff_2048x2048x80_f64", WORKLOADS["hdiff"]["reference_kernels"][2:], "hdiff", (2048, 2048, 80)),
]


def cpu_other_configs(primary, budget_s=20.0):
    """The reference's OpenMP kernels of the BASELINE configs other than `primary`, on all host
    cores: per config the best variant's mean sweep time and ALGORITHMIC GB/s (the byte formulas of
    `algorithmic_bytes`, basic: 2*N*s).  Bounded: every variant gets an equal share of `budget_s`,
    at least one warm and one timed sweep."""
    from oracle import ref_cpu

    if not ref_cpu.available():
        return {"unavailable": "oracle/_ref is not built"}
    ref_cpu.use_all_cores()
    present = ref_cpu.manifest()
    configs = [(label, [k for k in kernels if k in present], workload, domain)
               for label, kernels, workload, domain in CPU_OTHER_CONFIGS
               if workload == "basic" or workload != primary or domain != tuple(WORKLOADS[primary]["domain"])]
    variants = sum(len(kernels) for _, kernels, _, _ in configs)
    result = {}
    for label, kernels, workload, domain in configs:
        best = None
        fields = layout = None
        for name in kernels:
            share = budget_s / max(variants, 1)
            try:
                kernel = ref_cpu.Kernel(name)
                # variants rendered for the same layout sweep the same fields (generating 11 GB of
                # vadv fields takes longer than timing them)
                wanted = tuple(kernel.entry[key] for key in ("strides", "alignment", "dtype", "args"))
                if fields is None or wanted != layout:
                    fields = None
                    fields, layout = kernel.fields(seed=0, fast=True), wanted
                kernel(fields)
                times = []
                start = time.perf_counter()
                while not times or (time.perf_counter() - start < share and len(times) < 50):
                    times.append(kernel(fields))
                mean = sum(times) / len(times)
                if best is None or mean < best[1]:
                    best = (name, mean, len(times), kernel.isa)
                del kernel
            except Exception as error:  # noqa: BLE001 - one variant must not lose the others
                result.setdefault("failed", []).append(f"{name}: {error}")
        del fields
        if best is None:
            continue
        if workload == "basic":
            nbytes = 2 * domain[0] * domain[1] * domain[2] * 8
        else:
            nbytes = algorithmic_bytes(workload, domain)
        result[label] = {"kernel": best[0], "ms": best[1] * 1e3, "gbs": nbytes / best[1] / 1e9,
                         "sweeps": best[2], "march": best[3], "variants_tried": len(kernels)}
    result["cores"] = ref_cpu.threads()
    return result


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
        sample = (f"{best['sweeps']} full sweeps of {'x'.join(map(str, cfg['domain']))} float64"
                  + (f" (one of the {args.gpus} J slabs of the global domain per step)" if args.gpus > 1 else "") + ", "
                  + ("reference OpenMP kernel" if best.get("kind", "reference") == "reference" else "kernel")
                  + f" {best['name']} (-march={best['isa']}), best of: {', '.join(best['tried'])}")
    line = {
        "impl": "reference",
        "metric": METRIC[args.workload], "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": best["sweeps"], "warmup": max(args.warmup, 1), "ms_per_step": best["mean_s"] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        # the B200 arm's config at this N (the contract: same metric, unit and config in both arms)
        "config": workload_config(args.workload, args.gpus, "peer", "weak", False),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": best["threads"],
                         "kind": best.get("kind", "reference"), "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("the host has no partition and no exchange: it sweeps the global domain of `config` slab by slab in "
                 "shared memory; each timed step is ONE slab (the BASELINE domain, 1/N of the global domain) -- a "
                 "bounded sample; GB/s is a rate and does not depend on how many of the N equal slabs are swept"),
    }
    if not args.no_extras:
        line["also"] = cpu_other_configs(args.workload)
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


def first_global_row(workload, n_gpus, rank, scaling):
    """Global index of the first interior row of a rank's slab."""
    ny = WORKLOADS[workload]["domain"][1]
    if n_gpus == 1:
        return 0
    if scaling == "strong":
        from stencil_benchmarks_b200 import distributed

        return distributed.split_rows(ny, n_gpus)[rank][0]
    return rank * ny


def workload_config(workload, n_gpus, exchange=None, scaling="weak", iterate=False):
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
    config = {
        "workload": (f"{workload} {nx}x{ny}x{nz} float64 " + ("global" if strong else "per GPU")
                     + ", halo 3, alignment 128"),
        "global_domain": [nx, ny if strong else ny * n_gpus, nz],
        "per_gpu_domain": list(per_gpu),
        "partition": "J slabs, one per GPU" if n_gpus > 1 else "single GPU",
        "halo_exchange": (EXCHANGE_TEXT.get(exchange, "none") if workload == "hdiff" and n_gpus > 1 else "none"),
        "bytes_per_step_per_gpu": algorithmic_bytes(workload, per_gpu),
        "l2": f"fields ({resident_gb:.1f} GB per GPU) exceed the 126 MB L2; no flush between steps",
    }
    if workload == "hdiff":
        config["mode"] = ("time loop: inp/out swap every step, neighbours ordered by step flags in peer memory"
                          if iterate else "repeated sweep of the same input field")
    return config


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def checked(status, what):
    """Every C call of the timed loop goes through here: a failed launch must stop the run, not
    be timed as a no-op."""
    if status != 0:
        raise RuntimeError(f"{what} failed (status {status}); see stderr")


def make_global_rows(name, rows, nx, nz, halo, coeff_scale=1.0):
    """Padded global rows of the synthetic 'inp' / 'coeff' field."""
    from stencil_benchmarks_b200 import distributed

    block = distributed.global_rows(FIELD_SEEDS[name], rows, nx, nz, halo)
    if name == "coeff" and coeff_scale != 1.0:
        block *= coeff_scale
    return block


def fill_hdiff_slab(bench, data, start, has_lower, has_upper, coeff_scale=1.0):
    """The rank's rows of the global synthetic inp / coeff fields; j-halo rows that belong to a
    neighbour are poisoned (the exchange has to bring them, whichever it is)."""
    import numpy as np

    nx, ny, nz = (int(d) for d in bench.domain)
    halo = tuple(int(h) for h in bench.halo)
    rows = range(start, start + ny + 2 * halo[1])
    for name in ("inp", "coeff"):
        getattr(data, name)[...] = make_global_rows(name, rows, nx, nz, halo, coeff_scale)
    if has_lower:
        data.inp[:, :halo[1], :] = np.nan
    if has_upper:
        data.inp[:, halo[1] + ny:, :] = np.nan


def edge_parity(bench, out_host, start, mode, coeff_scale=1.0):
    """Edge row blocks of this rank's `out` against the C oracle applied to the GLOBAL field (whose
    rows every rank can regenerate): the first and the last EDGE_ROWS rows are the ones whose inputs
    cross the slab boundary."""
    import numpy as np

    from oracle import native

    nx, ny, nz = (int(d) for d in bench.domain)
    halo = tuple(int(h) for h in bench.halo)
    hy = halo[1]
    block = min(EDGE_ROWS, ny)
    worst, ok = 0.0, True
    for first in sorted({0, ny - block}):
        rows = range(start + first, start + first + block + 2 * hy)
        shape = (nx + 2 * halo[0], block + 2 * hy, nz + 2 * halo[2])
        fields = {}
        for name in ("inp", "coeff"):
            fields[name] = np.asfortranarray(make_global_rows(name, rows, nx, nz, halo, coeff_scale))
        expected = np.zeros(shape, order="F")
        native.hdiff(fields["inp"], fields["coeff"], expected, halo)
        got = out_host[halo[0]:halo[0] + nx, hy + first:hy + first + block, halo[2]:halo[2] + nz]
        want = expected[halo[0]:halo[0] + nx, hy:hy + block, halo[2]:halo[2] + nz]
        difference = np.abs(got - want)
        worst = max(worst, float("inf") if np.isnan(difference).any() else float(difference.max()))
        ok = ok and bool(np.allclose(got, want, rtol=1e-13, atol=1e-14))
    return {"mode": mode, "ok": ok, "max_abs_err": worst, "rows_checked_per_edge": block,
            "tolerance": "rtol 1e-13, atol 1e-14 against oracle/oracle.c on the global field"}


def time_loop_expected(make_rows, halo, padded_rows_global, rows, steps):
    """Rows `rows` (padded GLOBAL row indices, a short contiguous range) of the global field after
    `steps` sweeps of the time loop, from the C oracle applied to a window wide enough that the
    window's artificial edges (2 rows of contamination per sweep) never reach `rows`.
    `make_rows(field, range)` builds padded global rows of 'inp' / 'coeff'."""
    import numpy as np

    from oracle import native

    margin = 2 * steps + halo[1]
    lo = max(0, rows[0] - margin)
    hi = min(padded_rows_global, rows[-1] + 1 + margin)
    window = range(lo, hi)
    x = np.asfortranarray(make_rows("inp", window))
    coeff = np.asfortranarray(make_rows("coeff", window))
    y = x.copy(order="F")
    for _ in range(steps):
        native.hdiff(x, coeff, y, halo)
        x, y = y, x
    return x[:, rows[0] - lo:rows[-1] + 1 - lo, :]


def time_loop_parity(loop, bench, scratch, start, global_interior_rows, steps, mode):
    """Edge row blocks of the slab's state after `steps` sweeps against the oracle iterated on the
    global field."""
    import numpy as np

    nx, ny, nz = (int(d) for d in bench.domain)
    halo = tuple(int(h) for h in bench.halo)
    hy = halo[1]
    state = loop.download(scratch)
    block = min(EDGE_ROWS, ny)

    def make_rows(name, rows):
        return make_global_rows(name, rows, nx, nz, halo, TIME_LOOP_COEFF_SCALE)

    worst, ok = 0.0, True
    for first in sorted({0, ny - block}):
        rows = range(start + hy + first, start + hy + first + block)  # padded global indices
        want = time_loop_expected(make_rows, halo, global_interior_rows + 2 * hy, rows, steps)
        want = want[halo[0]:halo[0] + nx, :, halo[2]:halo[2] + nz]
        got = state[halo[0]:halo[0] + nx, hy + first:hy + first + block, halo[2]:halo[2] + nz]
        difference = np.abs(got - want)
        worst = max(worst, float("inf") if np.isnan(difference).any() else float(difference.max()))
        ok = ok and bool(np.allclose(got, want, rtol=1e-11, atol=1e-13))
    return {"mode": mode, "ok": ok, "max_abs_err": worst, "sweeps": steps, "rows_checked_per_edge": block,
            "tolerance": "rtol 1e-11, atol 1e-13 against oracle/oracle.c iterated on the global field"}


def measure_pcie(torch, nbytes=1 << 30, repeats=3):
    """Pinned-memory copy rates of this process' GPU: each direction alone and both at once."""
    host_up = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host_down = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dev_up = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev_down = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(up, down):
        best = None
        for _ in range(repeats):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if up:
                with torch.cuda.stream(s_up):
                    dev_up.copy_(host_up, non_blocking=True)
            if down:
                with torch.cuda.stream(s_down):
                    host_down.copy_(dev_down, non_blocking=True)
            torch.cuda.synchronize()
            elapsed = time.perf_counter() - t0
            best = elapsed if best is None else min(best, elapsed)
        return nbytes / best / 1e9

    result = {"h2d_gbs": timed(True, False), "d2h_gbs": timed(False, True)}
    result["duplex_gbs_each_way"] = timed(True, True)
    result["how"] = f"{nbytes >> 20} MiB pinned <-> device, best of {repeats}, wall clock around the copies"
    return result


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

    def reduce_max(value):
        if dist is None:
            return value
        t = torch.tensor([value], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lib = capi.library()
    raw = lib.raw
    cfg = WORKLOADS[args.workload]
    itemsize = np.dtype(cfg["dtype"]).itemsize
    domain = local_domain(args.workload, world, rank, args.scaling)
    nx, ny, nz = domain
    start_row = first_global_row(args.workload, world, rank, args.scaling)
    lower, upper = distributed.neighbours(rank, world)
    cls = horizontal_diffusion.Fused if args.workload == "hdiff" else vertical_advection.Thomas
    bench = cls(domain=domain, halo=cfg["halo"], dtype=cfg["dtype"], verify=False,
                device=local_rank, seed=100 + rank, dry_runs=0)
    data = bench.data()
    if args.workload == "hdiff":
        fill_hdiff_slab(bench, data, start_row, lower is not None, upper is not None,
                        TIME_LOOP_COEFF_SCALE if args.iterate else 1.0)
    mirrors = bench._device_fields(data)
    bench.upload(data, mirrors)
    pointers = {name: bench.interior_ptr(mirrors[name][1], host) for name, host in zip(bench.args, data)}
    geometry = bench.geometry()
    sy, sz = geometry[4], geometry[5]
    code = capi.dtype_code(cfg["dtype"])

    main_stream = torch.cuda.current_stream()
    comm_stream = torch.cuda.Stream(priority=-1)
    exchange = None
    peers = None
    if args.workload == "hdiff" and world > 1 and args.exchange == "peer":
        # fused exchange: neighbours' inp slabs mapped through CUDA IPC, halo rows read by TMA.
        # If any rank cannot map its neighbours (IPC disabled in the container), every rank
        # switches to the NCCL exchange -- still a GPU path -- and the line says so.
        try:
            peers = distributed.attach_neighbours(bench, dist, rank, world)
            mapped = 1
        except Exception as error:  # noqa: BLE001 - reported, not swallowed
            print(f"rank {rank}: peer mapping failed ({error}); using the NCCL exchange", file=sys.stderr)
            mapped = 0
        flag = torch.tensor([mapped], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if peers is not None:
                peers.close()
                peers = bench.peers = None
            args.exchange = "nccl (peer mapping failed on some rank: NOT the default path)"
    if args.workload == "hdiff" and world > 1 and peers is None:
        exchange = distributed.cuda_halo_exchange(rank, world, cfg["dtype"], nx, ny, nz, cfg["halo"][0],
                                                  sy, sz, width=cfg["halo"][1])
        (lo, hi), strips = distributed.interior_and_boundary_rows(
            ny, 2, exchange.lower is not None, exchange.upper is not None)

    def vp(value):
        return ctypes.c_void_p(value)

    iterate = None
    if args.iterate and args.workload == "hdiff":
        iterate = distributed.TimeLoop(bench, mirrors, dist, rank, world)

    def hdiff_rows(j0, j1, stream):
        if j1 <= j0:
            return
        offset = j0 * sy * itemsize
        checked(raw.sb200_hdiff(code, vp(pointers["inp"].value + offset), vp(pointers["coeff"].value + offset),
                                vp(pointers["out"].value + offset), nx, j1 - j0, nz, 1, sy, sz, 0, None,
                                vp(stream.cuda_stream)), "sb200_hdiff")

    def step():
        if args.workload == "vadv":
            bench.launch(pointers, 0, None, main_stream.cuda_stream)
        elif iterate is not None:
            iterate.step(main_stream.cuda_stream)
        elif peers is not None:
            checked(raw.sb200_hdiff_peer(code, pointers["inp"], pointers["coeff"], pointers["out"],
                                         vp(peers.lower), peers.ny_lower, peers.sz_lower,
                                         vp(peers.upper), peers.ny_upper, peers.sz_upper,
                                         nx, ny, nz, 1, sy, sz, 0, None, vp(main_stream.cuda_stream)),
                    "sb200_hdiff_peer")
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

    barrier()  # every rank has uploaded its fields before a neighbour reads them
    for _ in range(args.warmup):
        step()
    barrier()
    launches_before = capi.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    uuid = str(getattr(torch.cuda.get_device_properties(local_rank), "uuid", "") or "")
    with ClockSampler(local_rank, uuid) as clocks:
        barrier()
        start.record(main_stream)
        for _ in range(args.steps):
            step()
        stop.record(main_stream)
        barrier()
    elapsed_ms = reduce_max(start.elapsed_time(stop))
    launches = capi.launch_count() - launches_before
    ms_per_step = elapsed_ms / args.steps
    # bytes of one rank (rank 0's slab is the largest); whole job = every rank's slab
    nbytes = algorithmic_bytes(args.workload, local_domain(args.workload, world, 0, args.scaling))
    job_bytes = sum(algorithmic_bytes(args.workload, local_domain(args.workload, world, r, args.scaling))
                    for r in range(world))
    value = job_bytes / (ms_per_step * 1e-3) / 1e9

    # ---- parity of what the timed loop computed: edge rows of every slab against the oracle ----
    parity = None
    if args.workload == "hdiff":
        if iterate is not None:
            global_rows = cfg["domain"][1] * (1 if args.scaling == "strong" else world)
            mode = "time loop, " + ("single GPU" if world == 1 else "step flags in peer memory")
            note = None
            if iterate.count > TIME_LOOP_MAX_CHECKED_SWEEPS:
                # hundreds of sweeps smooth the field until neighbouring values coincide to the
                # last bit; the limiter's strict `> 0` test then flips on the sign of a rounding error
                # and FMA-contracted GPU code and the oracle part ways by O(coeff * flux).  The
                # check is therefore made on a fresh loop of the same code path (the long loop itself
                # is checked bit for bit against a single-GPU loop in tests/test_gpu_multi.py).
                note = (f"checked on a fresh loop of {TIME_LOOP_MAX_CHECKED_SWEEPS} sweeps after the timed "
                        f"loop of {iterate.count}")
                iterate.close()
                bench.upload(data, mirrors)
                barrier()
                iterate = distributed.TimeLoop(bench, mirrors, dist, rank, world)
                for _ in range(TIME_LOOP_MAX_CHECKED_SWEEPS):
                    iterate.step(main_stream.cuda_stream)
                barrier()
            parity = time_loop_parity(iterate, bench, data.out, start_row, global_rows, iterate.count, mode)
            if note:
                parity["note"] = note
        else:
            bench.download(data, mirrors)
            mode = "none (single GPU)" if world == 1 else ("peer" if peers is not None else "nccl")
            parity = edge_parity(bench, data.out, start_row, mode)
        if dist is not None:
            everyone = [None] * world
            dist.all_gather_object(everyone, parity)
            parity = dict(parity, ok=all(p["ok"] for p in everyone),
                          max_abs_err=max(p["max_abs_err"] for p in everyone), ranks_checked=world)
        if iterate is not None:
            iterate.close()
            iterate = None

    # ---- end to end through the plugin API: H2D inputs + kernel + D2H outputs per step ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    bench.chunks = args.e2e_chunks  # slab-pipelined upload / sweep / download on three streams
    h2d, d2h = bench.transfer_bytes()
    barrier()
    bench.run()  # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        bench.run()
    torch.cuda.synchronize()
    e2e_s = reduce_max((time.perf_counter() - t0) / e2e_steps)
    e2e_value = job_bytes / e2e_s / 1e9
    if args.workload == "hdiff" and peers is not None:
        e2e_parity = edge_parity(bench, data.out, start_row, "peer, partitioned run()",
                                 TIME_LOOP_COEFF_SCALE if args.iterate else 1.0)
        e2e_how = (f"plugin run() of every slab with chunks={args.e2e_chunks}: pinned host fields; per step the "
                   "edge rows go up first, the ranks wait for their neighbours' edge rows, then H2D / sweep (halo "
                   "rows read from the neighbours' HBM) / D2H run slab-pipelined; a barrier ends the step")
    else:
        e2e_parity = None
        e2e_how = (f"plugin run() with chunks={args.e2e_chunks}: pinned host fields, per step H2D of the "
                   "fields the sweep reads, the sweep, D2H of the field it writes, slab-pipelined")
    pcie = measure_pcie(torch) if not args.no_extras else None

    peak, peak_source = measured_peak()
    achieved = nbytes / (ms_per_step * 1e-3) / 1e9
    e2e = {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": h2d * world,
           "d2h_bytes_per_step": d2h * world, "steps": e2e_steps, "ms_per_step": e2e_s * 1e3, "how": e2e_how,
           "bound": "pcie"}
    if e2e_parity is not None:
        e2e["parity"] = e2e_parity
    if pcie is not None:
        # both directions run at once: the step cannot be shorter than the longer of the two copies
        floor_s = max(h2d, d2h) / (pcie["duplex_gbs_each_way"] * 1e9)
        e2e.update(pcie=pcie, frac_of_pcie_floor=floor_s / e2e_s,
                   note=("bound by the host link, not by the kernel: one GPU behind one PCIe 5 x16 link moves "
                         f"{(h2d + d2h) / 1e9:.1f} GB per sweep, a host CPU sweeping in place moves none; this ratio "
                         "stays below 1 at N=1 by construction and scales with the number of host links"))
    line = {
        "metric": METRIC[args.workload], "value": value, "unit": "GB/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args.workload, world, args.exchange, args.scaling, bool(args.iterate)),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": profiled_traffic(args.workload),
                     "peak_source": peak_source,
                     "kernel": "hdiff_tma_kernel<double>" if args.workload == "hdiff" else "vadv_onchip_kernel<double>",
                     "note": "per GPU; achieved = algorithmic bytes / mean step time (CUDA events); traffic = "
                             "dram bytes of one launch from the committed ncu capture (profiles/traffic.json)"},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "pct_of_nominal_8TBs": achieved / NOMINAL_PEAK_GBS,
    }
    if parity is not None:
        line["exchange_parity"] = parity

    if peers is not None:
        barrier()
        peers.close()
        bench.peers = None
    del bench, mirrors, data, pointers
    if not args.no_extras:
        line["also"] = extra_kernels(lib, capi, args.workload, reduce_max, world, peak)
    if rank == 0 and world == 1 and not args.no_extras and not args.no_cpu_baseline:
        line.update(baselines_beside(args, nbytes, ms_per_step))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def baselines_beside(args, nbytes, ms_per_step):
    """The keys of the N = 1 line that are measured beside the GPU numbers, after them: `cpu_baseline`
    (the reference's OpenMP kernels on the host cores, bounded sample; contract), and the two
    optional comparisons `cpu_baseline.also` and `reference_gpu`, which are skipped -- and say so --
    once the process has used its wall-clock allowance.  Nothing here may lose the GPU numbers."""
    keys = {}
    try:
        best = time_reference(args.workload, steps=5, warmup=1, budget_s=24.0)
        kind = best.get("kind", "reference")
        keys["cpu_baseline"] = {
            "value": nbytes / best["mean_s"] / 1e9, "unit": "GB/s", "cores": best["threads"], "kind": kind,
            "sample": f"{best['sweeps']} full sweeps, " + ("reference OpenMP kernel" if kind == "reference" else "kernel")
                      + f" {best['name']} (-march={best['isa']}), best of: {', '.join(best['tried'])}"}
    except Exception as error:  # noqa: BLE001
        keys["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": os.cpu_count(),
                                "kind": "reference", "sample": f"failed: {error}"}

    def within_allowance():
        return time.perf_counter() - PROCESS_START < args.optional_until_s

    skipped = {"unavailable": f"skipped: the run was past --optional-until-s {args.optional_until_s:.0f} s"}
    try:
        keys["cpu_baseline"]["also"] = (cpu_other_configs(args.workload, budget_s=12.0)
                                        if within_allowance() else skipped)
    except Exception as error:  # noqa: BLE001
        keys["cpu_baseline"]["also"] = {"unavailable": f"{type(error).__name__}: {error}"}
    keys["reference_gpu"] = reference_gpu(args.workload, ms_per_step) if within_allowance() else skipped
    return keys


def reference_gpu(workload, ours_ms):
    """The reference's own CUDA kernels (unmodified templates, sm_100, best block size of the sweep
    in oracle/_ref) on this GPU, results cross-checked against the C oracle: the kernel to beat."""
    try:
        from oracle import ref_cuda

        prefix = "cuda_hdiff" if workload == "hdiff" else "cuda_vadv"
        results = ref_cuda.run(only=prefix, repeat=3, check=False, verbose=False)
        rows = [r for r in results if not r.get("failed") and "uvw" not in r["kernel"]]
        if not rows:
            return {"unavailable": "no oracle/_ref CUDA kernel ran"}
        # time the three fastest again, with the cross-check
        shortlist = sorted(rows, key=lambda r: r["median_s"])[:3]
        final = []
        for row in shortlist:
            final += ref_cuda.run(only=row["kernel"], repeat=7, check=True, verbose=False)
        final = [r for r in final if r["kernel"] in {s["kernel"] for s in shortlist} and r.get("check_ok")]
        if not final:
            return {"unavailable": "no variant passed the cross-check"}
        best = min(final, key=lambda r: r["median_s"])
        return {"kernel": best["kernel"], "reference_class": best["reference_class"],
                "block_size": best["block_size"], "ms": best["median_s"] * 1e3,
                "gbs_algorithmic": best["gbs_algorithmic"], "variants_timed": len(rows),
                "check_max_abs_err": best.get("check_max_abs_err"), "speedup_of_b200_kernel": best["median_s"] * 1e3 / ours_ms}
    except Exception as error:  # noqa: BLE001 - a baseline must not lose the GPU numbers
        return {"unavailable": f"{type(error).__name__}: {error}"}


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
        checked(raw.sb200_stream_op(capi.STREAM_TRIAD, capi.F64, *ptrs, n, 3.0, 0, None, stream_handle),
                "sb200_stream_op")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(value, op):
        if dist is None:
            return value
        t = torch.tensor([value], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=op)
        return float(t.item())

    for _ in range(args.warmup):
        step()
    barrier()
    launches_before = capi.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    uuid = str(getattr(torch.cuda.get_device_properties(local_rank), "uuid", "") or "")
    with ClockSampler(local_rank, uuid) as clocks:
        barrier()
        start.record()
        for _ in range(args.steps):
            step()
        stop.record()
        barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = capi.launch_count() - launches_before
    if dist is not None:
        elapsed_ms = reduce(elapsed_ms, dist.ReduceOp.MAX)
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
        e2e_value = reduce(e2e_value, dist.ReduceOp.MIN)
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
        "pct_of_nominal_8TBs": achieved / NOMINAL_PEAK_GBS,
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


def extra_kernels(lib, capi, primary, reduce_max, world, peak, repeat=9):
    """Device-timed kernels of the other BASELINE.json configs, on every rank (independent per GPU:
    none of them communicates); per-GPU time = max over ranks of the median single-launch time."""
    import numpy as np

    from stencil_benchmarks_b200.scripts.kernel_bench import Field, padded_geometry

    vp = ctypes.c_void_p
    result = {}

    def record(name, nbytes, call, note=None):
        elapsed = ctypes.c_double()
        times = []
        for _ in range(repeat):
            call(ctypes.byref(elapsed))
            times.append(elapsed.value)
        seconds = reduce_max(statistics.median(times[1:]))
        gbs = nbytes / seconds / 1e9
        result[name] = {"ms": seconds * 1e3, "gbs_per_gpu": gbs, "gbs_total": gbs * world,
                        "frac_of_measured_peak": gbs / peak, "frac_of_nominal_8TBs": gbs / NOMINAL_PEAK_GBS,
                        "bytes_per_gpu": nbytes}
        if note:
            result[name]["note"] = note

    # STREAM at 2^30 float64 elements per array (BASELINE.json configs[1])
    n = TRIAD_N
    buffers = [capi.DeviceBuffer(8 * n) for _ in range(3)]
    ptrs = [b.ptr for b in buffers]
    lib.sb200_stream_op(capi.STREAM_INIT, capi.F64, *ptrs, n, 3.0, 0, None, None)
    for op, name, factor in ((capi.STREAM_COPY, "copy", 2), (capi.STREAM_SCALE, "scale", 2),
                             (capi.STREAM_ADD, "add", 3), (capi.STREAM_TRIAD, "triad", 3)):
        # scalar 1e-3 keeps the values finite over repeated triads
        record(f"stream_{name}_2^30_f64", factor * 8 * n,
               lambda t, op=op: lib.sb200_stream_op(op, capi.F64, *ptrs, n, 1e-3, 0, t, None))
    # ... and the size sweep of configs[1] (triad, this rank's GPU; small sizes live in the L2)
    sweep = {}
    elapsed = ctypes.c_double()
    for log2n in range(20, 30, 2):
        times = []
        for _ in range(repeat):
            lib.sb200_stream_op(capi.STREAM_TRIAD, capi.F64, *ptrs, 1 << log2n, 1e-3, 0, ctypes.byref(elapsed), None)
            times.append(elapsed.value)
        sweep[f"2^{log2n}"] = round(3 * 8 * (1 << log2n) / statistics.median(times[1:]) / 1e9, 1)
    sweep["2^30"] = round(result["stream_triad_2^30_f64"]["gbs_per_gpu"], 1)
    result["stream_triad_size_sweep_f64_gbs"] = sweep
    del buffers

    # basic stencils 1024x1024x80 (configs[2]), float32 and float64
    domain, halo = (1024, 1024, 80), (3, 3, 3)
    for dtype in ("float32", "float64"):
        size = np.dtype(dtype).itemsize
        code = capi.dtype_code(dtype)
        sy, sz, total, interior = padded_geometry(domain, halo, size)
        inp, out = Field(total, interior, size, 0.25), Field(total, interior, size, 0.5)
        nbytes = 2 * int(np.prod(domain)) * size
        tag = "f32" if size == 4 else "f64"
        for name, kind, axis, mask in (("copy", capi.BASIC_COPY, 0, 0),
                                       ("onesided_avg_i", capi.BASIC_ONESIDED_AVG, 0, 0),
                                       ("symmetric_avg_i", capi.BASIC_SYMMETRIC_AVG, 0, 0),
                                       ("laplacian_ij", capi.BASIC_LAPLACIAN, 0, 3)):
            record(f"basic_{name}_1024x1024x80_{tag}", nbytes,
                   lambda t, kind=kind, axis=axis, mask=mask: lib.sb200_basic(
                       kind, code, vp(inp.interior), vp(out.interior), *domain, 1, sy, sz, axis, mask, 0, t, None))
        del inp, out

    # vertical advection 1024x1024x160 float64 (configs[4]): u alone, and u/v/w in one sweep
    domain = (1024, 1024, 160)
    size, code = 8, capi.F64
    sy, sz, total, interior = padded_geometry(domain, halo, size)
    points = int(np.prod(domain))
    components = [[Field(total, interior, size, 0.1 + 0.1 * c + 0.02 * f) for f in range(4)] for c in range(3)]
    wcon = Field(total, interior, size, 0.7)

    def table(f, count):
        return (vp * count)(*[vp(components[c][f].interior) for c in range(count)])

    three = ctypes.c_int * 3
    if primary != "vadv":
        one = ctypes.c_int * 1
        record("vadv_1024x1024x160_f64", 6 * points * size,
               lambda t: lib.sb200_vadv_components(code, 1, table(0, 1), table(1, 1), table(2, 1), table(3, 1),
                                                   one(1), one(0), vp(wcon.interior), None, None, *domain, 1, sy, sz,
                                                   capi.VADV_AUTO, 0, t, None),
               note="in-out field re-swept in place: the Thomas solve is applied repeatedly, values stay finite")
    record("vadv_uvw_one_sweep_1024x1024x160_f64", 16 * points * size,
           lambda t: lib.sb200_vadv_components(code, 3, table(0, 3), table(1, 3), table(2, 3), table(3, 3),
                                               three(1, 0, 0), three(0, 1, 0), vp(wcon.interior), None, None,
                                               *domain, 1, sy, sz, capi.VADV_AUTO, 0, t, None),
           note="all_components=True: 13 reads + 3 writes per point, wcon shared through L2")
    del components, wcon

    if primary != "hdiff":
        domain = (2048, 2048, 80)
        sy, sz, total, interior = padded_geometry(domain, halo, size)
        fields = [Field(total, interior, size, v) for v in (0.2, 0.4, 0.6)]
        record("hdiff_2048x2048x80_f64", algorithmic_bytes("hdiff", domain),
               lambda t: lib.sb200_hdiff(code, *[vp(f.interior) for f in fields], *domain, 1, sy, sz, 0, t, None))
        del fields
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
    parser.add_argument("--iterate", action="store_true",
                        help="time loop: inp and out swap every step (hdiff); at N > 1 the sweeps of "
                             "neighbouring GPUs are ordered by step flags in peer memory")
    parser.add_argument("--e2e-steps", type=int, default=3)
    parser.add_argument("--e2e-chunks", type=int, default=8)
    parser.add_argument("--no-extras", action="store_true")
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--optional-until-s", type=float, default=300.0,
                        help="wall-clock seconds after which the optional comparisons of the N=1 line "
                             "(cpu_baseline.also, reference_gpu) are skipped")
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
