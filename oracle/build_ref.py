"""Build the reference's own OpenMP CPU kernels for the hot path into oracle/_ref/.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/stencils.py).  Runs in the dev
container, where /root/reference exists; the GPU box only uses the files this
script leaves in oracle/_ref/ (git-ignored, shipped by gpurun).

The reference generates its CPU kernels at run time: a Jinja template
(stencil_benchmarks/benchmarks_collection/stencils/openmp/templates/*.j2) is
rendered with the domain, strides and dtype baked in as literals
(openmp/mixin.py:64-83, :103-122) and compiled with
``g++ -std=c++11 -Wall -DNDEBUG -fopenmp -Ofast -march=native -mtune=native``
(openmp/mixin.py:88-101).  This script lets the reference do exactly that --
its classes are instantiated unmodified from a scratch copy of the tree (the
two pybind11 helper modules need an in-place build, which /root/reference,
being read-only, cannot hold) -- but intercepts ``GnuLibrary`` so that instead
of a temporary .so (deleted on exit, compilation.py:139-153) the rendered
source and the compile command are captured.  Each kernel is then compiled
twice with the reference's flags, ``-march=native`` replaced by the portable
ISA levels x86-64-v3 (AVX2) and x86-64-v4 (AVX-512) because the GPU box's CPU
is not this container's CPU; the loader (oracle/ref_cpu.py) picks the best
level the host supports.  No reference source is copied into the repository:
rendered sources, objects and the manifest live in oracle/_ref/ only.

Exported ABI of every library (openmp/templates/base.j2:213-251):
``int kernel(double* time, long long* counter, T* field0, ..., T* fieldN)`` --
host pointers to the first interior element, fields in ``args`` order.
"""

import json
import pathlib
import shutil
import subprocess
import sys

HERE = pathlib.Path(__file__).parent.resolve()
OUT = HERE / "_ref"
REFERENCE = pathlib.Path("/root/reference")
SCRATCH = pathlib.Path("/tmp/sb200_reference_build")
ISA_LEVELS = ["x86-64-v3", "x86-64-v4"]


def prepare_reference():
    """Scratch copy of the reference with its two pybind11 modules built in place."""
    marker = SCRATCH / ".built"
    if not marker.exists():
        if SCRATCH.exists():
            shutil.rmtree(SCRATCH)
        shutil.copytree(REFERENCE, SCRATCH)
        subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=SCRATCH,
                       check=True, capture_output=True)
        marker.touch()
    sys.path.insert(0, str(SCRATCH))


def install_package():
    """The reference's Python package, with its two pybind11 helpers built, as oracle/_ref/pkg:
    TEST infrastructure that travels to the GPU box (which has no /root/reference), so that the
    `-m gpu` drop-in tests can run the B200 classes under the reference's own ``Stencil.run()`` /
    ``verify_stencil`` / CLI.  Build output only -- oracle/_ref/ is git-ignored."""
    target = OUT / "pkg" / "stencil_benchmarks"
    if target.exists():
        shutil.rmtree(target)
    shutil.copytree(SCRATCH / "stencil_benchmarks", target,
                    ignore=shutil.ignore_patterns("__pycache__", "*.cpp", "test"))
    return target


def configurations():
    from stencil_benchmarks.benchmarks_collection.stencils.openmp import (
        basic,
        horizontal_diffusion as hdiff,
        vertical_advection as vadv,
    )

    f64 = dict(dtype="float64", verify=False, platform_preset="native")
    tuned = dict(alignment=64, **f64)
    return [
        # BASELINE.json configs[0]: the reference's own CPU-runnable case, defaults
        ("hdiff_otf_128x128x80_f64", hdiff.OnTheFly, dict(domain=(128, 128, 80), **f64)),
        # tuned variants in the spirit of scripts/sbench_rome_collection.py:191-223
        ("hdiff_otfvec_128x128x80_f64", hdiff.OnTheFlyVec,
         dict(domain=(128, 128, 80), vector_size=8, streaming_stores=True, block_size=(128, 16, 1), **tuned)),
        ("hdiff_otf_2048x2048x80_f64", hdiff.OnTheFly, dict(domain=(2048, 2048, 80), **tuned)),
        ("hdiff_otfvec_2048x2048x80_f64", hdiff.OnTheFlyVec,
         dict(domain=(2048, 2048, 80), vector_size=8, streaming_stores=True, block_size=(1024, 16, 1), **tuned)),
        ("hdiff_minimummem_2048x2048x80_f64", hdiff.MinimumMem,
         dict(domain=(2048, 2048, 80), vector_size=8, streaming_stores=True, block_size=(1024, 16, 1), **tuned)),
        ("hdiff_rolling_2048x2048x80_f64", hdiff.Rolling,
         dict(domain=(2048, 2048, 80), vector_size=8, streaming_stores=True, block_size=(1024, 16, 1), **tuned)),
        ("hdiff_rolling_128x128x80_f64", hdiff.Rolling,
         dict(domain=(128, 128, 80), vector_size=4, block_size=(128, 16, 1), **tuned)),
        ("vadv_kinnermost_128x128x80_f64", vadv.KInnermost, dict(domain=(128, 128, 80), **f64)),
        # scripts/sbench_rome_collection.py:244-256
        ("vadv_kmiddlevec_1024x1024x160_f64", vadv.KMiddleVec,
         dict(domain=(1024, 1024, 160), vector_size=8, block_size=(128, 1), streaming_stores=True, **tuned)),
        ("vadv_kinnermostvec_1024x1024x160_f64", vadv.KInnermostVec,
         dict(domain=(1024, 1024, 160), vector_size=8, block_size=(64, 1), **tuned)),
        ("copy_1dvec_1024x1024x80_f64", basic.Copy,
         dict(domain=(1024, 1024, 80), loop="1D-vec", vector_size=8, streaming_stores=True, **tuned)),
        ("laplacian_3dvec_1024x1024x80_f64", basic.Laplacian,
         dict(domain=(1024, 1024, 80), loop="3D-blocked-vec", vector_size=8, streaming_stores=True,
              block_size=(1024, 16, 1), **tuned)),
    ]


def cuda_configurations():
    """The reference's own GPU kernels rendered for the BASELINE.json sizes and compiled for sm_100
    ("the kernel to beat").  Seed: the block sizes of scripts/sbench_h100_collection.py:113-152
    (names without a size suffix); around it a grid over `block_size` (and `unroll_factor` for the
    vertical advection) for the variants that lead on Hopper, so that the comparison is against
    the best B200 configuration of the reference's kernels, not against an H100 tuning."""
    from stencil_benchmarks.benchmarks_collection.stencils.cuda_hip import (
        basic,
        horizontal_diffusion as hdiff,
        vertical_advection as vadv,
    )

    common = dict(backend="cuda", compiler="nvcc", gpu_architecture="sm_100", verify=False,
                  dry_runs=1, alignment=128, dtype="float64")
    hd = dict(domain=(2048, 2048, 80), **common)
    va = dict(domain=(1024, 1024, 160), **common)
    ba = dict(domain=(1024, 1024, 80), halo=(1, 1, 1), loop="3D", block_size=(128, 2, 1), **common)
    configs = [
        ("cuda_hdiff_classic", hdiff.Classic, dict(block_size=(32, 12, 1), **hd)),
        ("cuda_hdiff_otf", hdiff.OnTheFly, dict(block_size=(32, 16, 1), loop="3D", **hd)),
        ("cuda_hdiff_otfincache", hdiff.OnTheFlyIncache, dict(block_size=(32, 8, 1), **hd)),
        ("cuda_hdiff_jscansharedmem", hdiff.JScanSharedMem, dict(block_size=(256, 32, 1), **hd)),
        ("cuda_hdiff_jscanotfincache", hdiff.JScanOtfIncache, dict(block_size=(128, 4, 1), **hd)),
        ("cuda_hdiff_jscanotf", hdiff.JScanOtf, dict(block_size=(128, 4, 1), **hd)),
        ("cuda_hdiff_jscanshuffleincache", hdiff.JScanShuffleIncache, dict(block_size=(28, 8, 2), **hd)),
        ("cuda_hdiff_jscanshuffle", hdiff.JScanShuffle, dict(block_size=(28, 8, 2), **hd)),
        ("cuda_hdiff_jscanshufflesystolic", hdiff.JScanShuffleSystolic, dict(block_size=(28, 4, 3), **hd)),
        ("cuda_vadv_classic", vadv.Classic, dict(block_size=(128, 1), unroll_factor=8, **va)),
        ("cuda_vadv_localmem", vadv.LocalMem, dict(block_size=(128, 1), unroll_factor=28, **va)),
        ("cuda_vadv_sharedmem", vadv.SharedMem, dict(block_size=(64, 1), unroll_factor=0, **va)),
        ("cuda_vadv_localmemmerged", vadv.LocalMemMerged, dict(block_size=(128, 1), unroll_factor=2, **va)),
        ("cuda_vadv_localmemmerged_uvw", vadv.LocalMemMerged,
         dict(block_size=(128, 1), unroll_factor=2, all_components=True, **va)),
        ("cuda_basic_copy", basic.Copy, ba),
        ("cuda_basic_avg_i", basic.OnesidedAverage, dict(axis=0, **ba)),
        ("cuda_basic_lap_ij", basic.Laplacian, ba),
    ]
    # --- B200 sweep ---
    for by in (4, 8, 16, 32):
        for bz in (1, 2, 4):
            if (by, bz) != (8, 2):
                configs.append((f"cuda_hdiff_jscanshuffle_28x{by}x{bz}", hdiff.JScanShuffle,
                                dict(block_size=(28, by, bz), **hd)))
    for bx in (64, 128, 256):
        for by in (4, 8, 16, 32):
            if (bx, by) != (128, 4):
                configs.append((f"cuda_hdiff_jscanotf_{bx}x{by}x1", hdiff.JScanOtf,
                                dict(block_size=(bx, by, 1), **hd)))
    for bx, by in ((32, 4), (32, 8), (32, 16), (64, 8), (64, 4)):
        configs.append((f"cuda_hdiff_classic_{bx}x{by}x1", hdiff.Classic, dict(block_size=(bx, by, 1), **hd)))
    for by, bz in ((16, 2), (8, 4)):
        configs.append((f"cuda_hdiff_jscanshuffleincache_28x{by}x{bz}", hdiff.JScanShuffleIncache,
                        dict(block_size=(28, by, bz), **hd)))
    for bx in (32, 64, 256):
        for unroll in (8, 28):
            configs.append((f"cuda_vadv_localmem_{bx}_u{unroll}", vadv.LocalMem,
                            dict(block_size=(bx, 1), unroll_factor=unroll, **va)))
    configs.append(("cuda_vadv_localmem_128_u8", vadv.LocalMem, dict(block_size=(128, 1), unroll_factor=8, **va)))
    for bx in (32, 64, 256):
        configs.append((f"cuda_vadv_classic_{bx}_u8", vadv.Classic, dict(block_size=(bx, 1), unroll_factor=8, **va)))
    for bx in (32, 64):
        configs.append((f"cuda_vadv_localmemmerged_uvw_{bx}", vadv.LocalMemMerged,
                        dict(block_size=(bx, 1), unroll_factor=2, all_components=True, **va)))
    return configs


def build_cuda(captured, manifest):
    """Render + cross-compile the reference's CUDA kernels (sm_100 SASS; runs on the GPU box).
    Rendering goes through the reference's classes one by one; the nvcc runs are independent and
    run a few at a time."""
    import concurrent.futures
    import os

    jobs = []
    for name, cls, kwargs in cuda_configurations():
        try:
            bench = cls(**kwargs)
        except Exception as error:  # e.g. a variant that cannot be rendered for this size
            print(f"skipped {name}: {error}")
            continue
        source = OUT / "src" / (name + ".cu")
        source.write_text(captured["code"])
        target = OUT / f"{name}.sm_100.so"
        flags = [f for f in captured["command"][1:] if f not in ("-x", "cu")]
        command = [captured["command"][0], "-o", str(target), "-x", "cu", str(source)] + flags + [
            "-Xcompiler", "-shared", "-Xcompiler", "-fPIC"]
        entry = dict(
            reference_class=f"{cls.__module__}.{cls.__name__}",
            kwargs={k: (list(v) if isinstance(v, tuple) else v) for k, v in kwargs.items()},
            args=list(bench.args), domain=list(bench.domain), halo=list(bench.halo),
            strides=[int(s) for s in bench.strides], alignment=int(bench.alignment), dtype=bench.dtype,
            data_size=int(bench.data_size), compile_flags=flags, libraries={"sm_100": target.name},
        )
        jobs.append((name, command, entry))
        del bench

    def compile_one(job):
        name, command, entry = job
        return name, entry, subprocess.run(command, capture_output=True, text=True)

    with concurrent.futures.ThreadPoolExecutor(max_workers=max(2, (os.cpu_count() or 2))) as pool:
        for name, entry, result in pool.map(compile_one, jobs):
            if result.returncode != 0:
                print(f"skipped {name}: does not compile for this configuration: "
                      + result.stderr.strip().splitlines()[-1])
                continue
            manifest[name] = entry
            print(f"built {name}: strides {entry['strides']}")


def main():
    if not REFERENCE.exists():
        print("no /root/reference here: keeping the prebuilt oracle/_ref as it is")
        return 0
    prepare_reference()
    OUT.mkdir(parents=True, exist_ok=True)
    install_package()
    from stencil_benchmarks.benchmarks_collection.stencils import base
    from stencil_benchmarks.tools import compilation

    captured = {}

    class CaptureLibrary:
        def __init__(self, code, compile_command=None, extension=None):
            captured.update(code=code, command=list(compile_command), extension=extension or ".cpp")

    compilation.GnuLibrary = CaptureLibrary
    # the fields are only needed for their strides: skip the random fill of multi-GB arrays
    base.Stencil.random_field = base.Stencil.empty_field

    (OUT / "src").mkdir(parents=True, exist_ok=True)
    manifest = {}
    for name, cls, kwargs in configurations():
        bench = cls(**kwargs)
        source = OUT / "src" / (name + captured["extension"])
        source.write_text(captured["code"])
        flags = [f for f in captured["command"][1:] if not f.startswith(("-march", "-mtune"))]
        libraries = {}
        for isa in ISA_LEVELS:
            target = OUT / f"{name}.{isa}.so"
            command = [captured["command"][0], "-o", str(target), str(source)] + flags + [
                f"-march={isa}", "-shared", "-fPIC"]
            result = subprocess.run(command, capture_output=True, text=True)
            if result.returncode != 0:
                raise RuntimeError(result.stderr)
            libraries[isa] = target.name
        manifest[name] = dict(
            reference_class=f"{cls.__module__}.{cls.__name__}",
            kwargs={k: (list(v) if isinstance(v, tuple) else v) for k, v in kwargs.items()},
            args=list(bench.args),
            domain=list(bench.domain),
            halo=list(bench.halo),
            strides=[int(s) for s in bench.strides],
            alignment=int(bench.alignment),
            dtype=bench.dtype,
            data_size=int(bench.data_size),
            compile_flags=flags,
            libraries=libraries,
        )
        print(f"built {name}: strides {manifest[name]['strides']}")
        del bench
    build_cuda(captured, manifest)
    (OUT / "manifest.json").write_text(json.dumps(manifest, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
