"""Drop-in check against the REAL reference package (dev container only; skipped elsewhere):
with `stencil_benchmarks` importable, the B200 classes must subclass the reference's Benchmark,
land in its REGISTRY, inherit its NumPy oracle and appear in its CLI tree."""

import os
import pathlib
import subprocess
import sys

import pytest

REFERENCE = pathlib.Path("/root/reference")
SCRATCH = pathlib.Path("/tmp/sb200_reference_build")
ROOT = pathlib.Path(__file__).parent.parent.resolve()

pytestmark = pytest.mark.skipif(not REFERENCE.exists(), reason="reference tree not available")


@pytest.fixture(scope="module")
def reference_path():
    marker = SCRATCH / ".built"
    if not marker.exists():
        # the same scratch build oracle/build_ref.py uses (two pybind11 helpers, ~20 s)
        sys.path.insert(0, str(ROOT / "oracle"))
        import build_ref

        build_ref.prepare_reference()
    return str(SCRATCH)


def run(code, reference_path, *argv):
    env = dict(os.environ, PYTHONPATH=f"{reference_path}:{ROOT}")
    return subprocess.run([sys.executable, "-c", code, *argv], capture_output=True, text=True, env=env,
                          timeout=300, cwd="/tmp")


def test_classes_register_in_the_reference(reference_path):
    code = """
import stencil_benchmarks.benchmark as ref
import stencil_benchmarks.benchmarks_collection.stencils.base as ref_base
import stencil_benchmarks_b200.benchmark as ours
import stencil_benchmarks_b200.benchmarks_collection
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion, vertical_advection
from stencil_benchmarks.cli import _cli_command
assert ours.HAVE_REFERENCE and ours.REGISTRY is ref.REGISTRY and ours.Benchmark is ref.Benchmark
assert issubclass(horizontal_diffusion.Fused, ref_base.HorizontalDiffusionStencil)
assert horizontal_diffusion.Fused.verify_stencil is ref_base.HorizontalDiffusionStencil.verify_stencil
assert horizontal_diffusion.Fused in ref.REGISTRY and vertical_advection.Thomas in ref.REGISTRY
assert horizontal_diffusion.Fused.parameters["alignment"].default == 128
print(" ".join(_cli_command(horizontal_diffusion.Fused)))
b = horizontal_diffusion.Fused(domain=(10, 12, 5), pinned=False)   # verify=True is allowed here
print(b.strides, b.verify)
"""
    result = run(code, reference_path)
    assert result.returncode == 0, result.stderr
    lines = result.stdout.strip().splitlines()
    # _cli_command keeps module underscores; the click groups turn them into dashes (cli.py:230)
    assert lines[0] == "stencils b200 horizontal_diffusion fused"
    assert lines[1] == "(1, 16, 288) True"


def test_sbench_cli_lists_the_backend(reference_path):
    code = "import sys; from stencil_benchmarks_b200.scripts.sbench_b200 import main; sys.argv[0] = 'sbench'; main()"
    top = run(code, reference_path, "stencils", "b200", "--help")
    assert top.returncode == 0, top.stderr
    for name in ("basic", "horizontal-diffusion", "vertical-advection"):
        assert name in top.stdout
    fused = run(code, reference_path, "stencils", "b200", "horizontal-diffusion", "fused", "--help")
    assert fused.returncode == 0, fused.stderr
    for option in ("--domain", "--halo", "--dtype", "--alignment", "--dry-runs", "--verify", "--seed"):
        assert option in fused.stdout
    stream = run(code, reference_path, "stream", "b200", "native", "--help")
    assert stream.returncode == 0 and "--array-size" in stream.stdout


def test_jit_goes_through_the_reference_gnulibrary(reference_path, tmp_path):
    """With the reference importable, `compiler=nvcc` compiles the kernels through ITS
    tools.compilation.GnuLibrary (source kept under ./benchmarks_source_code, compilation.py:130-137)."""
    code = """
import pathlib
import stencil_benchmarks.tools.compilation as compilation
from stencil_benchmarks_b200 import capi
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion
bench = horizontal_diffusion.Fused(domain=(16, 12, 4), pinned=False, verify=False, compiler="nvcc")
assert isinstance(bench._kernels.library, compilation.GnuLibrary)
sources = list(pathlib.Path("benchmarks_source_code").glob("*.cu"))
assert len(sources) == 1 and "hdiff.cu" in sources[0].read_text()
try:
    bench._kernels.sb200_hdiff(1, None, None, None, 0, 4, 4, 1, 8, 64, 0, None, None)
except compilation.ExecutionError as error:
    print("reference ExecutionError:", str(error).strip())
"""
    env = dict(os.environ, PYTHONPATH=f"{reference_path}:{ROOT}")
    result = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                            timeout=300, cwd=str(tmp_path))
    assert result.returncode == 0, result.stderr
    assert "reference ExecutionError: sb200_hdiff: domain must be positive" in result.stdout


def test_integration_stub_registers_in_the_reference(reference_path):
    """The reference-side binding printed in INTEGRATION.md §2 is real code: executed as a module
    of the reference's benchmark collection it registers in the CLI tree, loads the library at
    setup(), resolves every prototype it sets, and maps a missing library to ParameterError."""
    text = (ROOT / "INTEGRATION.md").read_text()
    start = text.index("```python\n# b200/mixin.py") + len("```python\n")
    stub = text[start:text.index("```", start)]
    library = ROOT / "stencil_benchmarks_b200" / "csrc" / "libsbench_b200.so"
    code = f"""
import sys, types
import stencil_benchmarks.benchmark as ref
from stencil_benchmarks.cli import _cli_command
name = "stencil_benchmarks.benchmarks_collection.stencils.b200.horizontal_diffusion"
module = types.ModuleType(name)
sys.modules[name] = module
exec(compile({stub!r}, "INTEGRATION.md", "exec"), module.__dict__)
assert module.Fused in ref.REGISTRY and module.StencilMixin not in ref.REGISTRY
print(" ".join(_cli_command(module.Fused)))
bench = module.Fused(domain=(16, 12, 5), library={str(library)!r}, verify=False)
print(bench.strides, bench.lib.sb200_hdiff.argtypes is not None, bench.lib.sb200_version())
try:
    module.Fused(domain=(16, 12, 5), library="/nonexistent/libsbench_b200.so")
except ref.ParameterError as error:
    print("ParameterError")
"""
    result = run(code, reference_path)
    assert result.returncode == 0, result.stderr
    lines = result.stdout.strip().splitlines()
    assert lines[0] == "stencils b200 horizontal_diffusion fused"
    # alignment 128: rows of 16 + 2*3 doubles padded to 32; the library answers through the stub's handle
    assert lines[1] == f"(1, 32, {32 * 18}) True 100"
    assert lines[2] == "ParameterError"


def test_committed_golden_vectors_are_what_the_reference_produces(reference_path):
    """Provenance of tests/golden/*.npz: re-derive every case from the reference's own
    verify_stencil (the committed generator, unmodified reference) and compare bit for bit."""
    code = f"""
import sys
import numpy as np
sys.path.insert(0, {str(ROOT / "tests" / "golden")!r})
import make_golden as g
checked = 0
for seed, (name, stencil_class, kwargs, outputs) in enumerate(g.CASES):
    bench = g.probe(stencil_class)(**kwargs)
    data = g.seeded_fill(bench, 1000 + seed)
    expected = g.capture_expected(bench, data)
    stored = np.load(g.OUT / (name + ".npz"))
    for field_name, field in zip(bench.args, data):
        assert np.array_equal(stored["in_" + field_name], field), (name, field_name)
    for output in outputs:
        assert stored["expected_" + output].dtype == expected[output].dtype
        assert np.array_equal(stored["expected_" + output], expected[output]), (name, output)
    checked += 1
print(checked)
"""
    result = run(code, reference_path)
    assert result.returncode == 0, result.stderr
    assert int(result.stdout.strip().splitlines()[-1]) == len(list((ROOT / "tests" / "golden").glob("*.npz"))) == 40


def test_oracle_equals_the_live_reference_on_random_cases(reference_path):
    """Beyond the committed vectors: 48 random (stencil, domain, halo, dtype) cases, the expected
    arrays captured from the reference's verify_stencil live, both oracle restatements bit for bit."""
    code = f"""
import sys
import numpy as np
sys.path.insert(0, {str(ROOT / "tests" / "golden")!r})
sys.path.insert(0, {str(ROOT / "tests")!r})
import make_golden as g
import test_oracle as t
from stencil_benchmarks.benchmarks_collection.stencils import base

rng = np.random.default_rng(20261017)
checked = 0
for n in range(48):
    kind = ["copy", "onesided", "symmetric", "laplacian", "hdiff", "vadv", "vadv_all"][n % 7]
    dtype = ["float64", "float32"][(n // 7) % 2]
    domain = tuple(int(v) for v in rng.integers(3, 22, 3))
    halo = [int(v) for v in rng.integers(0, 4, 3)]
    kwargs, outputs, name = dict(dtype=dtype, verify=True), ["out"], kind
    if kind == "copy":
        cls = base.CopyStencil
    elif kind in ("onesided", "symmetric"):
        axis = int(rng.integers(0, 3))
        halo[axis] = max(halo[axis], 1)
        cls = base.OnesidedAverageStencil if kind == "onesided" else base.SymmetricAverageStencil
        kwargs["axis"], name = axis, f"{{kind}}_ax{{axis}}_x"
    elif kind == "laplacian":
        mask = int(rng.integers(1, 8))
        halo = [max(h, 1) if mask >> a & 1 else h for a, h in enumerate(halo)]
        cls = base.LaplacianStencil
        kwargs.update(along_x=bool(mask & 1), along_y=bool(mask & 2), along_z=bool(mask & 4))
        name = f"laplacian_m{{mask}}_x"
    elif kind == "hdiff":
        halo = [max(halo[0], 2), max(halo[1], 2), halo[2]]
        cls = base.HorizontalDiffusionStencil
    else:
        cls = base.VerticalAdvectionStencil
        domain = domain[:2] + (max(domain[2], 2),)
        if kind == "vadv_all":
            halo = [max(h, 1) for h in halo]
            kwargs["all_components"] = True
            outputs, name = ["utensstage", "vtensstage", "wtensstage"], "vadv_all_x"
        else:
            halo[0] = max(halo[0], 1)
            outputs, name = ["utensstage"], "vadv_x"
    bench = g.probe(cls)(domain=domain, halo=tuple(halo), **kwargs)
    data = g.seeded_fill(bench, 5000 + n)
    expected = g.capture_expected(bench, data)
    case = {{"in_" + f: np.ascontiguousarray(v) for f, v in zip(bench.args, data)}}
    case["halo"], case["domain"] = np.array(bench.halo), np.array(bench.domain)
    for restatement in (t.expected_numpy, t.expected_c):
        got = restatement(name, dict(case))
        for output in outputs:
            assert np.array_equal(got[output], expected[output]), (n, kind, dtype, domain, halo, restatement.__name__)
    checked += 1
print(checked)
"""
    result = run(code, reference_path)
    assert result.returncode == 0, result.stderr[-3000:]
    assert int(result.stdout.strip().splitlines()[-1]) == 48


def test_reference_unit_tests_pass_against_the_stand_alone_modules(tmp_path):
    """Where the reference package is not importable (the GPU box) `benchmark.py`, `tools/fields.py`
    and `tools/cabi.py` stand in for its `benchmark`, `tools.array` and `tools.compilation`.  The
    reference's OWN unit tests of those modules -- unmodified, loaded from the reference tree --
    must pass against the stand-ins (27 tests: plugin API, alloc_array, GnuLibrary and its helpers)."""
    code = f"""
import importlib.util, sys, types, unittest
import stencil_benchmarks_b200.benchmark as benchmark
assert not benchmark.HAVE_REFERENCE            # the reference is NOT on the path of this process
from stencil_benchmarks_b200.tools import cabi, fields
package = types.ModuleType("stencil_benchmarks"); package.__path__ = []
tools = types.ModuleType("stencil_benchmarks.tools"); tools.__path__ = []
package.benchmark, package.tools, tools.array, tools.compilation = benchmark, tools, fields, cabi
sys.modules.update({{"stencil_benchmarks": package, "stencil_benchmarks.benchmark": benchmark,
                    "stencil_benchmarks.tools": tools, "stencil_benchmarks.tools.array": fields,
                    "stencil_benchmarks.tools.compilation": cabi}})
suite = unittest.TestSuite()
for index, relative in enumerate(["test_benchmark.py", "tools/test_array.py", "tools/test_compilation.py"]):
    spec = importlib.util.spec_from_file_location(f"reference_test_{{index}}",
                                                  {str(REFERENCE / "stencil_benchmarks" / "test")!r} + "/" + relative)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    suite.addTests(unittest.defaultTestLoader.loadTestsFromModule(module))
result = unittest.TextTestRunner(stream=sys.stderr, verbosity=0).run(suite)
print(result.testsRun, len(result.failures), len(result.errors), len(result.skipped))
"""
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    result = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300,
                            cwd=tmp_path)
    assert result.returncode == 0, result.stderr[-3000:]
    assert result.stdout.strip().splitlines()[-1] == "27 0 0 0", result.stderr[-3000:]


def test_stand_alone_stencil_definitions_equal_the_reference(reference_path, tmp_path):
    """The stand-alone stencil base classes (used where the reference is absent) against the
    reference's classes, live, on 60 random parameter sets: field names, data_size, strides, field
    shapes, alignment of the first interior element, inner_slice -- and the same ParameterError
    verdicts for halos the stencil cannot live with."""
    code = f"""
import sys
import numpy as np
import stencil_benchmarks_b200.benchmark as ours_benchmark
assert not ours_benchmark.HAVE_REFERENCE
import stencil_benchmarks_b200.benchmarks_collection.stencils.base as ours
sys.path.insert(0, {reference_path!r})
import stencil_benchmarks.benchmark as ref_benchmark
import stencil_benchmarks.benchmarks_collection.stencils.base as ref

def probe(cls):
    class Probe(cls):
        def run_stencil(self, data):
            return dict(time=1.0)
    return Probe

def build(module, errors, name, kwargs):
    try:
        return probe(getattr(module, name))(**kwargs), None
    except errors as error:
        return None, type(error).__name__

rng = np.random.default_rng(7)
names = ["CopyStencil", "OnesidedAverageStencil", "SymmetricAverageStencil", "LaplacianStencil",
         "HorizontalDiffusionStencil", "VerticalAdvectionStencil"]
compared = refused = 0
for n in range(60):
    name = names[n % len(names)]
    kwargs = dict(domain=tuple(int(v) for v in rng.integers(1, 40, 3)),
                  halo=tuple(int(v) for v in rng.integers(0, 4, 3)),
                  dtype=["float64", "float32"][int(rng.integers(0, 2))],
                  alignment=int(rng.choice([0, 8, 64, 128, 12])), verify=False)
    if "Average" in name:
        kwargs["axis"] = int(rng.integers(0, 3))
    if name == "LaplacianStencil":
        kwargs.update(along_x=bool(rng.integers(0, 2)), along_y=bool(rng.integers(0, 2)), along_z=bool(rng.integers(0, 2)))
    if name == "VerticalAdvectionStencil":
        kwargs["all_components"] = bool(rng.integers(0, 2))
    a, a_error = build(ours, (ours_benchmark.ParameterError,), name, kwargs)
    b, b_error = build(ref, (ref_benchmark.ParameterError,), name, kwargs)
    assert a_error == b_error, (name, kwargs, a_error, b_error)
    if a is None:
        refused += 1
        continue
    assert tuple(a.args) == tuple(b.args), (name, kwargs)
    assert int(a.data_size) == int(b.data_size), (name, kwargs, a.data_size, b.data_size)
    assert tuple(int(s) for s in a.strides) == tuple(int(s) for s in b.strides), (name, kwargs, a.strides, b.strides)
    assert tuple(a.domain_with_halo) == tuple(b.domain_with_halo)
    fa, fb = a._data[0][0], b._data[0][0]
    assert fa.shape == fb.shape and fa.strides == fb.strides and fa.dtype == fb.dtype
    if kwargs["alignment"]:
        first = fa[tuple(slice(h, None) for h in kwargs["halo"])]
        assert first.ctypes.data % kwargs["alignment"] == 0
    assert a.inner_slice() == b.inner_slice() and a.inner_slice(shift=(1, 0, -1)) == b.inner_slice(shift=(1, 0, -1))
    assert a.inner_slice(expand=(1, 1, 0)) == b.inner_slice(expand=(1, 1, 0))
    compared += 1
print(compared, refused)
"""
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    result = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300,
                            cwd=tmp_path)
    assert result.returncode == 0, result.stderr[-3000:]
    compared, refused = (int(v) for v in result.stdout.split())
    assert compared + refused == 60 and compared >= 25 and refused >= 5


def test_reference_run_and_oracle_judge_the_backend_classes_on_an_emulated_device(reference_path):
    """The plugin contract under the REAL reference, without a GPU: with `stencil_benchmarks`
    importable the backend classes are driven by the reference's own `Stencil.run()` and judged by its
    `verify_stencil` (base.py:151-166) -- results written back into the host fields before returning,
    inputs bit-unchanged, a positive `time` and no `bandwidth` key.  The device is emulated
    (tests/test_host_datapath.py), so this checks the wiring, not the kernels; the same test runs
    with real kernels on the GPU box (tests/test_gpu_dropin.py)."""
    code = f"""
import sys
sys.path.insert(0, {str(ROOT / "tests")!r})
import numpy as np
import stencil_benchmarks.benchmark as ref
from stencil_benchmarks.benchmarks_collection.stencils import base as ref_base
import test_host_datapath as emulated
from stencil_benchmarks_b200 import capi
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import basic, horizontal_diffusion, vertical_advection

fake = emulated.FakePartitionDevice()
capi.require_device = lambda: None
capi.device_count = lambda: 8
capi.DeviceBuffer = emulated.FakeBuffer
capi.synchronize = lambda stream=None: None
import ctypes
capi.memcpy_h2d = lambda d, h, n, stream=None, sync=True: ctypes.memmove(d, h, n)
capi.memcpy_d2h = lambda h, d, n, stream=None, sync=True: ctypes.memmove(h, d, n)

cases = [
    (horizontal_diffusion.Fused, dict(domain=(21, 14, 4))),
    (horizontal_diffusion.Fused, dict(domain=(21, 14, 4), chunks=3)),
    (horizontal_diffusion.Partitioned, dict(domain=(21, 14, 4), gpus=2)),
    (vertical_advection.Thomas, dict(domain=(12, 9, 7))),
    (vertical_advection.Thomas, dict(domain=(12, 9, 7), all_components=True, chunks=2)),
    (basic.Laplacian, dict(domain=(15, 8, 5), along_z=True)),
    (basic.OnesidedAverage, dict(domain=(15, 8, 5), axis=1, chunks=2)),
    (basic.PartitionedLaplacian, dict(domain=(15, 8, 5), gpus=3)),
    (basic.PartitionedSymmetricAverage, dict(domain=(15, 8, 5), axis=1, gpus=2)),
]
for cls, kwargs in cases:
    assert issubclass(cls, ref.Benchmark) and cls.run is ref_base.Stencil.run
    bench = cls(pinned=False, verify=True, **kwargs)        # verify=True: the reference's default
    bench._lib = bench._kernels = fake
    result = bench.run()                                     # raises if verify_stencil finds a difference
    assert result["time"] > 0 and result["bandwidth"] > 0
# negative control: a sweep that is not executed must be rejected by the reference's validation
class Idle(emulated.FakePartitionDevice):
    def sb200_hdiff(self, *args):
        return self._finish(args[-2])   # reports a time, computes nothing
bench = horizontal_diffusion.Fused(pinned=False, verify=True, domain=(21, 14, 4))
bench._lib = bench._kernels = Idle()
import stencil_benchmarks.tools.validation as validation
try:
    bench.run()
except validation.ValidationError as error:
    print("rejected:", str(error).splitlines()[0])
print(len(cases))
"""
    result = run(code, reference_path)
    assert result.returncode == 0, result.stderr[-3000:]
    lines = result.stdout.strip().splitlines()
    assert lines[-1] == "9" and lines[-2].startswith("rejected: validation of field out failed at 1176 points")


def test_reference_cli_drives_the_backend_on_an_emulated_device(reference_path, tmp_path):
    """`sbench -e 2 -o out.csv stencils b200 ...` through the reference's click tree (cli.py:47-73; the
    reference's own CLI test does the same with its numpy backend, test/test_cli.py:38-54), range
    syntax and CSV writer included, verification on (the default) -- on the emulated device."""
    import pandas as pd

    code = f"""
import ctypes, sys
sys.path.insert(0, {str(ROOT / "tests")!r})
import stencil_benchmarks.benchmarks_collection
import stencil_benchmarks_b200.benchmarks_collection
import test_host_datapath as emulated
from stencil_benchmarks_b200 import capi
from stencil_benchmarks import cli
fake = emulated.FakePartitionDevice()
capi.require_device = lambda: None
capi.device_count = lambda: 8
capi.library = lambda: fake
capi.DeviceBuffer = emulated.FakeBuffer
capi.synchronize = lambda stream=None: None
capi.memcpy_h2d = lambda d, h, n, stream=None, sync=True: ctypes.memmove(d, h, n)
capi.memcpy_d2h = lambda h, d, n, stream=None, sync=True: ctypes.memmove(h, d, n)
cli.main(args=sys.argv[1:], standalone_mode=False)
"""
    out = tmp_path / "hdiff.csv"
    result = run(code, reference_path, "--executions", "2", "--output", str(out), "stencils", "b200",
                 "horizontal-diffusion", "fused", "--domain", "10", "10", "10", "--no-pinned", "--chunks", "[1,2]")
    assert result.returncode == 0, result.stdout + result.stderr[-3000:]
    table = pd.read_csv(out)
    assert len(table) == 4 and sorted(set(table["chunks"])) == [1, 2]
    assert (table["bandwidth"] > 0).all() and table["verify"].all()
    assert {"time", "bandwidth-algorithmic", "alignment", "sbench-version"} <= set(table.columns)

    out = tmp_path / "partitioned.csv"
    result = run(code, reference_path, "--executions", "1", "--output", str(out), "stencils", "b200", "basic",
                 "partitioned-laplacian", "--domain", "12", "9", "4", "--no-pinned", "--gpus", "[1,3]", "--along-z")
    assert result.returncode == 0, result.stdout + result.stderr[-3000:]
    assert list(pd.read_csv(out)["gpus"]) == [1, 3]
