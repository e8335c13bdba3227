"""Host side of the B200 backend without a GPU: registration names, parameter validation,
field layout, byte accounting, C-ABI symbol table."""

import ctypes
import subprocess

import numpy as np
import pytest

from stencil_benchmarks_b200 import benchmark, capi
from stencil_benchmarks_b200 import benchmarks_collection  # noqa: F401  (registers the classes)
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import (
    basic,
    horizontal_diffusion,
    vertical_advection,
)
from stencil_benchmarks_b200.benchmarks_collection.stream import b200 as stream
from stencil_benchmarks_b200.tools import cabi, fields

CPU = dict(pinned=False, verify=False)  # no device needed to construct


def cli_path(cls):
    """Command path stencil_benchmarks.cli derives for a class (cli.py:47-50, :230)."""
    def kebab(name):
        out = ""
        for ch in name:
            out += ("-" + ch.lower()) if ch.isupper() and out else ch.lower()
        return out
    return [part.replace("_", "-") for part in cls.__module__.split(".")[2:]] + [kebab(cls.__name__)]


def test_registered_command_names():
    names = {" ".join(cli_path(c)) for c in benchmark.REGISTRY
             if c.__module__.startswith("stencil_benchmarks_b200.")}
    assert names == {
        "stencils b200 basic empty", "stencils b200 basic copy",
        "stencils b200 basic onesided-average", "stencils b200 basic symmetric-average",
        "stencils b200 basic laplacian", "stencils b200 horizontal-diffusion fused",
        "stencils b200 horizontal-diffusion partitioned",
        "stencils b200 vertical-advection thomas", "stream b200 native",
        "stencils b200 basic partitioned-copy", "stencils b200 basic partitioned-onesided-average",
        "stencils b200 basic partitioned-symmetric-average", "stencils b200 basic partitioned-laplacian",
    }


def test_library_exports_every_declared_symbol():
    declared = capi.declared_symbols()
    assert set(declared) == set(capi.PROTOTYPES), "include/sbench_b200.h and capi.PROTOTYPES differ"
    lib = capi.library()
    for name in declared:
        assert isinstance(getattr(lib.raw, name), ctypes._CFuncPtr)
    dynamic = subprocess.run(["nm", "-D", "--defined-only", str(capi.LIBRARY_PATH)],
                             capture_output=True, text=True).stdout
    for name in declared:
        assert f" T {name}" in dynamic
    assert lib.raw.sb200_version() >= 100


def test_library_contains_sm100a_code_with_tma():
    """The shipped kernels are sm_100a SASS and the hdiff/vadv fast paths really use TMA."""
    out = subprocess.run(["cuobjdump", "-lelf", str(capi.LIBRARY_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", str(capi.LIBRARY_PATH)], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass          # cp.async.bulk.tensor
    assert "LDG.E.128" in sass and "STG.E" in sass
    assert "DFMA" in sass


def test_failed_call_raises_with_stderr_message():
    lib = capi.library()
    with pytest.raises(cabi.ExecutionError, match="domain must be positive"):
        lib.sb200_hdiff(capi.F64, None, None, None, 0, 4, 4, 1, 8, 64, 0, None, None)
    with pytest.raises(cabi.ExecutionError, match="layout"):
        lib.sb200_basic(capi.BASIC_COPY, capi.F64, None, None, 4, 4, 4, 2, 8, 64, 0, 0, 0, None, None)
    with pytest.raises(cabi.ExecutionError, match="unsupported dtype"):
        lib.sb200_stream_run(7, 16, 2, 0)


def test_kernels_compile_at_setup_like_the_reference_backends(tmp_path, monkeypatch):
    """`compiler=...`: source string -> compile -> ctypes handle at setup() (cuda_hip/mixin.py:60-84);
    compilation problems are ParameterErrors; calls carry the header's prototypes."""
    monkeypatch.chdir(tmp_path)  # GnuLibrary keeps its sources under ./benchmarks_source_code
    bench = vertical_advection.Thomas(domain=(16, 8, 4), compiler="nvcc", **CPU)
    assert isinstance(bench._kernels, capi.TypedLibrary) and bench._kernels is not bench._lib
    with pytest.raises((cabi.ExecutionError, RuntimeError), match="null component table"):
        bench._kernels.sb200_vadv_components(capi.F64, 1, None, None, None, None, None, None, None, None, None,
                                             4, 4, 1, 1, 8, 64, 0, 0, None, None)
    again = vertical_advection.Thomas(domain=(8, 8, 4), compiler="nvcc", **CPU)
    assert again._kernels is bench._kernels  # one compilation per process
    with pytest.raises(benchmark.ParameterError, match="error"):
        vertical_advection.Thomas(domain=(8, 8, 4), compiler="nvcc", compiler_flags="-Dint=", **CPU)
    with pytest.raises(benchmark.ParameterError, match="not found"):
        vertical_advection.Thomas(domain=(8, 8, 4), compiler="/no/such/nvcc", **CPU)


def test_parameters_of_cuda_hip_scripts_are_accepted():
    """Keyword sets the reference's collection scripts pass to its cuda_hip classes
    (scripts/sbench_h100_collection.py:72-152) construct the B200 classes unchanged."""
    common = dict(backend="cuda", verify=False, dry_runs=1, alignment=128, dtype="float32", pinned=False)
    assert basic.Copy(loop="3D", block_size=(128, 2, 1), halo=(1, 1, 1), domain=(32, 32, 8), **common).loop == "3D"
    basic.Laplacian(loop="1D", block_size=(1024, 1, 1), threads_per_block=(0, 0, 0), domain=(32, 32, 8), **common)
    horizontal_diffusion.Fused(block_size=(28, 8, 2), index_type="int", domain=(32, 32, 8), **common)
    vertical_advection.Thomas(block_size=(128, 1), unroll_factor=28, domain=(32, 32, 8), **common)
    native = stream.Native(array_size=1000, block_size=256, vector_size=4, unroll_factor=2, axis="x",
                           explicit_vectorization=True, launch_bounds=True, index_type="std::size_t",
                           streaming_stores=True, streaming_loads=True, dtype="float64")
    assert native.array_size % 16 == 0
    with pytest.raises(benchmark.ParameterError, match="vector_size"):
        stream.Native(array_size=1000, vector_size=3)
    with pytest.raises(benchmark.ParameterError, match="unroll_factor"):
        stream.Native(array_size=1000, unroll_factor=3)
    with pytest.raises(benchmark.ParameterError, match="block_size"):
        stream.Native(array_size=1000, block_size=100)
    with pytest.raises(benchmark.ParameterError, match="together"):
        stream.Native(array_size=1000, streaming_loads=False)
    with pytest.raises(benchmark.ParameterError):
        stream.Native(array_size=1000, load_cache_modifier="cg")


def test_no_device_means_error_not_fallback():
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    bench = horizontal_diffusion.Fused(domain=(8, 8, 4), **CPU)
    with pytest.raises(benchmark.ExecutionError, match="no CUDA device"):
        bench.run()
    with pytest.raises(benchmark.ParameterError, match="no CUDA device"):
        horizontal_diffusion.Fused(domain=(8, 8, 4), verify=False)  # pinned host memory needs the device
    with pytest.raises(benchmark.ExecutionError):
        stream.Native(array_size=1024).run()


def test_parameter_errors():
    with pytest.raises(benchmark.ParameterError, match="layout"):
        horizontal_diffusion.Fused(layout=(0, 1, 2), **CPU)
    with pytest.raises(benchmark.ParameterError, match="at least 2"):
        horizontal_diffusion.Fused(halo=(1, 2, 0), **CPU)
    with pytest.raises(benchmark.ParameterError, match="positive halo"):
        vertical_advection.Thomas(halo=(0, 3, 3), **CPU)
    with pytest.raises(benchmark.ParameterError, match="positive halo"):
        vertical_advection.Thomas(halo=(1, 0, 1), all_components=True, **CPU)
    with pytest.raises(benchmark.ParameterError, match="positive halo"):
        basic.OnesidedAverage(axis=1, halo=(1, 0, 1), **CPU)
    with pytest.raises(benchmark.ParameterError, match="halo"):
        basic.Laplacian(halo=(0, 1, 1), **CPU)
    with pytest.raises(benchmark.ParameterError, match="at least one axis"):
        basic.Laplacian(along_x=False, along_y=False, **CPU)
    with pytest.raises(benchmark.ParameterError, match="dtype"):
        basic.Copy(dtype="int32", **CPU)
    with pytest.raises(benchmark.ParameterError):
        basic.Copy(gpu_architecture="sm_90", **CPU)
    with pytest.raises(benchmark.ParameterError, match="not divisible"):
        basic.Copy(alignment=12, **CPU)
    if not benchmark.HAVE_REFERENCE:
        assert basic.Copy(pinned=False).verify is False  # stand-alone default: no oracle in the product
        with pytest.raises(benchmark.ParameterError, match="verify"):
            basic.Copy(pinned=False, verify=True)  # needs the reference's oracle


@pytest.mark.parametrize("dtype,alignment", [("float64", 128), ("float32", 128), ("float64", 0)])
def test_field_layout_matches_reference_rules(dtype, alignment):
    bench = horizontal_diffusion.Fused(domain=(128, 128, 80), dtype=dtype, alignment=alignment, **CPU)
    data = bench.data()
    assert data._fields == ("inp", "coeff", "out")
    size = np.dtype(dtype).itemsize
    sx, sy, sz = bench.strides
    assert sx == 1 and sz == sy * 134
    if alignment:
        # 134-long rows pad to 144 (f64) / 160 (f32); first interior element aligned
        assert sy == {8: 144, 4: 160}[size]
        for field in data:
            interior = field.ctypes.data + sum(s * h for s, h in zip(field.strides, bench.halo))
            assert interior % alignment == 0
    else:
        assert sy == 134
    assert all(f.shape == (134, 134, 86) for f in data)
    assert all(0 <= f.min() and f.max() < 1 for f in data)
    assert bench.geometry() == (128, 128, 80, 1, sy, sz)


def test_seeded_fields_reproduce():
    a = basic.Copy(domain=(9, 7, 5), seed=3, **CPU).data()
    b = basic.Copy(domain=(9, 7, 5), seed=3, **CPU).data()
    c = basic.Copy(domain=(9, 7, 5), seed=4, **CPU).data()
    assert np.array_equal(a.inp, b.inp) and not np.array_equal(a.inp, c.inp)
    assert not np.array_equal(a.inp, a.out)


def test_byte_accounting_at_baseline_sizes():
    """sbench figures and algorithmic minima."""
    hd = horizontal_diffusion.Fused(domain=(128, 128, 80), **CPU)
    assert hd.data_size == 32680448
    # pure arithmetic for the large configs (constructing them would allocate tens of GB)
    class Shape:
        dtype = "float64"
        all_components = False
    big = Shape()
    big.domain = (2048, 2048, 80)
    assert horizontal_diffusion.HorizontalDiffusionMixin.algorithmic_bytes.fget(big) == 8063559680
    big.domain = (1024, 1024, 160)
    assert vertical_advection.VerticalAdvectionMixin.algorithmic_bytes.fget(big) == 8053063680
    va = vertical_advection.Thomas(domain=(16, 16, 8), **CPU)
    assert va.data_size == 10 * 16 * 16 * 8 * 8 and va.algorithmic_bytes == 6 * 16 * 16 * 8 * 8
    assert va.data()._fields == ("ustage", "upos", "utens", "utensstage", "wcon", "ccol", "dcol", "datacol")
    va3 = vertical_advection.Thomas(domain=(16, 16, 8), all_components=True, **CPU)
    assert len(va3.args) == 16 and va3.data_size == 20 * 16 * 16 * 8 * 8


def test_stream_array_size_is_padded_with_a_warning():
    with pytest.warns(UserWarning, match="adapting array size"):
        s = stream.Native(array_size=1000001)
    assert s.array_size == 1000016
    assert stream.Native(array_size=1 << 20).array_size == 1 << 20
    with pytest.raises(benchmark.ParameterError):
        stream.Native(ntimes=1)


def test_fields_alloc_and_nbytes():
    x = fields.alloc_array((2, 3), "int32", (1, 0))
    assert x.strides == (4, 8)
    y = fields.alloc_array((2, 3), "int32", (0, 1), alignment=64)
    assert y.strides == (64, 4) and y.ctypes.data % 64 == 0 and fields.nbytes(y) == 76
    z = fields.alloc_array((4, 5, 6), "float64", (2, 1, 0), 128, index_to_align=(1, 1, 1))
    assert (z.ctypes.data + 8 + z.strides[1] + z.strides[2]) % 128 == 0
    with pytest.raises(ValueError):
        fields.alloc_array((2, 3), "int32", (0, 0))
    assert cabi.data_ptr(z, (1, 1, 1)).value == z.ctypes.data + 8 + z.strides[1] + z.strides[2]
    assert cabi.dtype_cname("float64") == "double" and cabi.dtype_cname("int16") == "std::int16_t"


def test_missing_library_is_a_parameter_error(monkeypatch, tmp_path):
    """No library, no benchmark: like a failed compilation in the reference (cuda_hip/mixin.py:79-84)."""
    capi.library.cache_clear()
    monkeypatch.setattr(capi, "LIBRARY_PATH", tmp_path / "libsbench_b200.so")
    try:
        with pytest.raises(benchmark.ParameterError, match="does not exist"):
            horizontal_diffusion.Fused(domain=(8, 8, 4), **CPU)
        with pytest.raises(benchmark.ParameterError, match="does not exist"):
            stream.Native(array_size=1024)
    finally:
        monkeypatch.undo()
        capi.library.cache_clear()
        capi.library()


def test_bench_accounting_and_config():
    import importlib.util
    import pathlib

    spec = importlib.util.spec_from_file_location("bench", pathlib.Path(__file__).parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.algorithmic_bytes("hdiff", (2048, 2048, 80)) == 8063559680
    assert bench.algorithmic_bytes("vadv", (1024, 1024, 160)) == 8053063680
    assert bench.algorithmic_bytes("triad", None) == 3 * 8 * (1 << 30)
    cfg = bench.workload_config("hdiff", 8, "peer")
    assert cfg["global_domain"] == [2048, 16384, 80] and "peer memory" in cfg["halo_exchange"]
    assert bench.workload_config("hdiff", 1, "peer")["halo_exchange"] == "none"
    assert "NCCL" in bench.workload_config("hdiff", 2, "nccl")["halo_exchange"]
    peak, source = bench.measured_peak()
    assert peak > 1000 and ("measured" in source or "fallback" in source)
    # strong scaling: the BASELINE domain is split into J slabs, the job's bytes are the slabs' bytes
    assert bench.local_domain("hdiff", 8, 3, "strong") == (2048, 256, 80)
    assert bench.local_domain("hdiff", 3, 0, "strong") == (2048, 683, 80)
    assert bench.local_domain("hdiff", 3, 2, "strong") == (2048, 682, 80)
    assert bench.local_domain("hdiff", 8, 3, "weak") == (2048, 2048, 80)
    assert bench.local_domain("hdiff", 1, 0, "strong") == (2048, 2048, 80)
    strong = bench.workload_config("hdiff", 8, "peer", "strong")
    assert strong["global_domain"] == [2048, 2048, 80] and strong["per_gpu_domain"] == [2048, 256, 80]
    assert strong["bytes_per_step_per_gpu"] == bench.algorithmic_bytes("hdiff", (2048, 256, 80))
    assert sum(bench.local_domain("vadv", 4, r, "strong")[1] for r in range(4)) == 1024


def _hdiff_tiles(dtype, domain):
    """Every CTA's (i tile, level, first row, last row) decoded the way hdiff_tma_kernel does."""
    import ctypes

    lib = capi.library()
    xtiles, segments, jt, ctas = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
    lib.sb200_hdiff_tiling(capi.dtype_code(dtype), *domain, ctypes.byref(xtiles), ctypes.byref(segments),
                           ctypes.byref(jt), ctypes.byref(ctas))
    b = np.arange(ctas.value)
    xt = b % xtiles.value
    seg = (b // xtiles.value) % segments.value
    k = b // (xtiles.value * segments.value)
    j0 = seg * jt.value
    j1 = np.minimum(j0 + jt.value, domain[1])
    return xtiles.value, xt, k, j0, j1


@pytest.mark.parametrize("cfg", [None, "0,64", "0,6", "0,2048"])
@pytest.mark.parametrize("dtype,domain", [("float64", (2048, 2048, 80)), ("float64", (300, 67, 5)),
                                          ("float32", (1000, 33, 7)), ("float64", (64, 1, 1)),
                                          ("float32", (513, 4100, 2))])
def test_hdiff_tiling_covers_every_row_once(monkeypatch, dtype, domain, cfg):
    """The work decomposition of the TMA kernel is a partition: every (i tile, level, row) belongs
    to exactly one CTA (include/sbench_b200.h: sb200_hdiff_tiling)."""
    if cfg is not None:
        monkeypatch.setenv("SB200_HDIFF_CFG", cfg)
    nx, ny, nz = domain
    xtiles, xt, k, j0, j1 = _hdiff_tiles(dtype, domain)
    tile_width = 128 * (2 if dtype == "float64" else 4)
    assert xtiles == -(-nx // tile_width)
    assert (j1 > j0).all() and k.min() == 0 and k.max() == nz - 1 and xt.max() == xtiles - 1
    rows = np.zeros((xtiles, nz, ny), dtype=np.int32)
    for a, b_, c, d in zip(xt, k, j0, j1):
        rows[a, b_, c:d] += 1
    assert (rows == 1).all()
    if cfg is None:
        assert (j1 - j0).max() <= 32  # default segments (profiles/hdiff_segments_r01.log)


def test_call_convention_of_a_loaded_library(tmp_path):
    """The GnuLibrary convention (tools/compilation.py:155-196 of the reference, its doctest at
    :104-116): stdout of the C function is the call's value, a non-zero status raises with the
    stderr text, stderr text of a successful call is a warning, `argtypes=` sets the prototype;
    the descriptors are back in place afterwards."""
    import ctypes
    import os
    import warnings

    source = tmp_path / "convention.c"
    source.write_text(
        '#include <stdio.h>\n'
        'int hello(void) { printf("Hello world!"); fflush(stdout); return 0; }\n'
        'int fails(int code) { fprintf(stderr, "reason %d", code); fflush(stderr); return code; }\n'
        'int chatty(void) { fprintf(stderr, "note"); fflush(stderr); printf("data"); fflush(stdout); return 0; }\n'
        'int scaled(double x) { printf("%.1f", 2 * x); fflush(stdout); return 0; }\n')
    before = (os.fstat(1).st_ino, os.fstat(2).st_ino)
    lib = cabi.compile_library([source], tmp_path / "convention.so", ["gcc", "-O1"])
    assert lib.hello() == "Hello world!"
    assert lib.hello.__name__ == "hello"
    with pytest.raises(cabi.ExecutionError, match="reason 3"):
        lib.fails(ctypes.c_int(3))
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        assert lib.chatty() == "data"
    assert any("chatty" in str(w.message) and "note" in str(w.message) for w in caught)
    assert lib.scaled(1.25, argtypes=[ctypes.c_double]) == "2.5"
    with pytest.raises(AttributeError):
        lib.no_such_function
    assert (os.fstat(1).st_ino, os.fstat(2).st_ino) == before
    with pytest.raises(cabi.CompilationError):
        bad = tmp_path / "bad.c"
        bad.write_text("int broken( { return 0; }\n")
        cabi.compile_library([bad], tmp_path / "bad.so", ["gcc"])


def test_ctypes_prototypes_match_the_header_parameter_by_parameter():
    """A ctypes prototype that drifts from include/sbench_b200.h (a missing argument, an int where
    the header says int64_t) corrupts the call on the GPU box only: compare them here."""
    import re

    text = capi.HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    declarations = re.findall(r"\b(int|uint64_t)\s+(sb200_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text)
    assert {name for _, name, _ in declarations} == set(capi.PROTOTYPES)
    scalars = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64,
               "size_t": ctypes.c_size_t, "double": ctypes.c_double, "uint32_t": ctypes.c_uint32,
               "unsigned": ctypes.c_uint32, "unsigned int": ctypes.c_uint32}
    for restype, name, parameters in declarations:
        expected_restype, argtypes = capi.PROTOTYPES[name]
        assert scalars[restype] is expected_restype, name
        declared = [p.strip() for p in parameters.replace("\n", " ").split(",")
                    if p.strip() and p.strip() != "void"]
        assert len(declared) == len(argtypes), f"{name}: {len(declared)} parameters declared, {len(argtypes)} bound"
        for position, (parameter, bound) in enumerate(zip(declared, argtypes)):
            where = f"{name} argument {position} ({parameter})"
            if "*" in parameter:
                assert bound in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(bound, ctypes._Pointer), where
                continue
            ctype = re.sub(r"\bconst\b", "", parameter).strip().rsplit(" ", 1)[0].strip()
            assert scalars[ctype] is bound, where


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C99 and a C program must link
    against the library and call it (entry points that need no GPU: version, launch counter, the
    host-side tiling query, an argument error with its stderr reason)."""
    header_dir = capi.HEADER_PATH.parent
    for compiler, language, standard in (("gcc", "c", "-std=c99"), ("g++", "c++", "-std=c++11")):
        check = subprocess.run([compiler, standard, "-Wall", "-Wextra", "-pedantic", "-fsyntax-only", "-x", language,
                                str(capi.HEADER_PATH)], capture_output=True, text=True)
        assert check.returncode == 0 and not check.stderr, check.stderr
    source = tmp_path / "client.c"
    source.write_text(
        '#include <stdio.h>\n#include "sbench_b200.h"\n'
        "int main(void) {\n"
        "  int rows = 0, segments = 0, xtiles = 0; int64_t ctas = 0;\n"
        "  if (sb200_version() < 100) return 2;\n"
        "  if (sb200_launch_count() != 0) return 3;\n"
        "  if (sb200_hdiff_tiling(SB200_F64, 2048, 2048, 80, &xtiles, &segments, &rows, &ctas) != 0) return 4;\n"
        '  printf("%d %d %d %lld\\n", rows, segments, xtiles, (long long)ctas);\n'
        "  return sb200_stream_configure(33, 0, 0, -1) != 0 ? 0 : 5;  /* rejected: not a multiple of 32 */\n"
        "}\n")
    binary = tmp_path / "client"
    build = subprocess.run(["gcc", "-std=c99", "-Wall", "-I", str(header_dir), str(source), "-o", str(binary),
                            str(capi.LIBRARY_PATH), f"-Wl,-rpath,{capi.LIBRARY_PATH.parent}"],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(binary)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, (run.returncode, run.stderr)
    rows, segments, xtiles, ctas = (int(v) for v in run.stdout.split())
    assert rows == 32 and segments == 64 and xtiles == 8 and ctas == 8 * 64 * 80
    assert "multiple of 32" in run.stderr


def test_stream_run_parses_the_mccalpin_table(monkeypatch):
    """`stream b200 native`.run(): the library prints the reference's table (stream/cuda_hip.j2:
    250-275), the plugin turns it into the reference's result dicts (stream/cuda_hip.py:127-144):
    bandwidth in MB/s from the minimum time, avg / min / max time in seconds."""
    table = ("Function    Best Rate MB/s  Avg time     Min time     Max time\n"
             "Copy:         6921034.2     0.002490     0.002482     0.002511\n"
             "Scale:        6913201.9     0.002491     0.002485     0.002500\n"
             "Add:          7113400.0     0.003630     0.003623     0.003650\n"
             "Triad:        7133020.5     0.003620     0.003613     0.003641\n")
    calls = []

    class Fake:
        def sb200_set_device(self, device):
            calls.append(("device", device))

        def sb200_stream_configure(self, *args):
            calls.append(("configure", args))

        def sb200_stream_run(self, *args):
            calls.append(("run", args))
            return table

    native = stream.Native(array_size=1 << 20, ntimes=7, dtype="float32", vector_size=8, unroll_factor=2,
                           block_size=256, device=3)
    monkeypatch.setattr(capi, "require_device", lambda: None)
    native._lib = native._kernels = Fake()
    results = native.run()
    assert [r["name"] for r in results] == ["copy", "scale", "add", "triad"]
    assert results[3] == {"name": "triad", "bandwidth": 7133020.5, "avg-time": 0.003620, "time": 0.003613,
                          "max-time": 0.003641}
    # vector_size is in elements (the reference's meaning): 8 floats = 32-byte vectors
    assert calls == [("device", 3), ("configure", (256, 2, 32, 1)), ("run", (capi.F32, 1 << 20, 7, 1))]
