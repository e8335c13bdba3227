"""Loader for oracle/_ref: the reference's own OpenMP CPU kernels, built by oracle/build_ref.py.

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by tests/ (to validate the oracle
restatement against the real reference code) and by bench.py's ``cpu_baseline``
and ``--impl reference`` legs (kind "reference").  Never imported by the product.
"""

import ctypes
import json
import os
import pathlib
import time

import numpy as np

HERE = pathlib.Path(__file__).parent.resolve()
REF = HERE / "_ref"


def available():
    return (REF / "manifest.json").exists()


def manifest():
    return json.loads((REF / "manifest.json").read_text())


def cpu_flags():
    try:
        for line in pathlib.Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("flags"):
                return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def best_isa():
    flags = cpu_flags()
    if {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags:
        return "x86-64-v4"
    if {"avx2", "fma", "bmi2"} <= flags:
        return "x86-64-v3"
    raise RuntimeError("host CPU supports neither x86-64-v3 nor x86-64-v4")


def alloc_field(entry):
    """Uninitialised field with exactly the strides the kernel was rendered for."""
    dtype = np.dtype(entry["dtype"])
    shape = tuple(d + 2 * h for d, h in zip(entry["domain"], entry["halo"]))
    strides = tuple(s * dtype.itemsize for s in entry["strides"])
    extent = sum((n - 1) * s for n, s in zip(shape, strides)) + dtype.itemsize
    alignment = max(int(entry["alignment"]), dtype.itemsize)
    raw = np.empty(extent + alignment, dtype=np.uint8)
    interior = sum(s * h for s, h in zip(strides, entry["halo"]))
    offset = -(raw.ctypes.data + interior) % alignment
    return np.ndarray(shape=shape, dtype=dtype, buffer=raw, offset=offset, strides=strides)


def cpu_model():
    try:
        for line in pathlib.Path("/proc/cpuinfo").read_text().splitlines():
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def native_library(name, entry):
    """The kernel compiled ON THIS MACHINE with the reference's own flags, ``-march=native
    -mtune=native`` included (openmp/mixin.py:88-101), from the source the reference rendered
    (oracle/_ref/src, written by build_ref.py).  Cached per CPU model under oracle/_ref/native/.
    Returns None where that is not possible (no g++, no source): the prebuilt portable-ISA
    libraries are used then."""
    import hashlib
    import shutil
    import subprocess

    compiler = shutil.which("g++")
    source = REF / "src" / (name + ".cpp")
    if compiler is None or not source.exists():
        return None
    tag = hashlib.sha1((cpu_model() + "".join(sorted(cpu_flags()))).encode()).hexdigest()[:10]
    target = REF / "native" / f"{name}.{tag}.so"
    if not target.exists():
        target.parent.mkdir(parents=True, exist_ok=True)
        scratch = target.with_suffix(f".{os.getpid()}.tmp")
        command = [compiler, "-o", str(scratch), str(source)] + list(entry["compile_flags"]) + [
            "-march=native", "-mtune=native", "-shared", "-fPIC"]
        try:
            result = subprocess.run(command, capture_output=True, text=True, timeout=300)
        except (OSError, subprocess.SubprocessError):
            return None
        if result.returncode != 0:
            return None
        os.replace(scratch, target)
    return target


class Kernel:
    """One compiled reference kernel: ``kernel(double* time, long long* counter, T* fields...)``.

    ``isa``: "native" (default: compiled here with the reference's ``-march=native``, falling back
    to the best prebuilt level), or one of the prebuilt levels x86-64-v3 / x86-64-v4."""

    def __init__(self, name, isa="native"):
        self.name = name
        self.entry = manifest()[name]
        path = None
        if isa in (None, "native"):
            path = native_library(name, self.entry)
            isa = "native" if path is not None else best_isa()
        if path is None:
            path = REF / self.entry["libraries"][isa]
        self.isa = isa
        self.library = ctypes.CDLL(str(path))
        self.function = self.library.kernel
        self.function.restype = ctypes.c_int

    def fields(self, seed=0, fast=False):
        """Seeded U[0,1) fields in ``args`` order.

        ``fast`` draws one random plane per field and scales it per level (enough for
        timing multi-GB fields; values stay in (0, 1) and differ between levels)."""
        result = []
        for index, _ in enumerate(self.entry["args"]):
            field = alloc_field(self.entry)
            rng = np.random.default_rng([seed, index])
            levels = field.shape[2]
            plane = rng.random(field.shape[:2]) if fast else None
            for k in range(levels):
                if fast:
                    np.multiply(plane, 0.5 + 0.5 * (k + 1) / levels, out=field[:, :, k])
                else:
                    field[:, :, k] = rng.random(field.shape[:2])
            result.append(field)
        return result

    def __call__(self, fields):
        """One sweep; returns the time the kernel measured itself (seconds)."""
        elapsed = ctypes.c_double()
        counter = ctypes.c_longlong()
        pointers = [
            ctypes.c_void_p(f.ctypes.data + sum(s * h for s, h in zip(f.strides, self.entry["halo"])))
            for f in fields
        ]
        if self.function(ctypes.byref(elapsed), ctypes.byref(counter), *pointers) != 0:
            raise RuntimeError(f"reference kernel {self.name} failed")
        return elapsed.value


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_cores():
    """Make the OpenMP kernels use every core this process may run on.

    torchrun exports OMP_NUM_THREADS=1 for its workers; the reference arm is meant to use all
    host threads, so the setting is overridden in the environment (for a runtime that is not
    loaded yet) and through omp_set_num_threads (for one that is)."""
    n = usable_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except OSError:
        pass
    return n


def threads():
    try:
        return int(ctypes.CDLL("libgomp.so.1").omp_get_max_threads())
    except OSError:
        return int(os.environ.get("OMP_NUM_THREADS", usable_cores()))


def time_kernel(name, budget_s=10.0, warmup=1, seed=0):
    """Median sweep time of a reference kernel within a time budget."""
    kernel = Kernel(name)
    fields = kernel.fields(seed)
    for _ in range(warmup):
        kernel(fields)
    samples = []
    start = time.perf_counter()
    while not samples or (time.perf_counter() - start < budget_s and len(samples) < 50):
        samples.append(kernel(fields))
    samples.sort()
    return dict(name=name, isa=kernel.isa, threads=threads(), sweeps=len(samples),
                median_s=samples[len(samples) // 2], min_s=samples[0], entry=kernel.entry)
