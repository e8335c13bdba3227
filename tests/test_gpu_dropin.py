"""The B200 backend under the REAL reference package, on the GPU.

oracle/_ref/pkg holds the reference's Python package (unmodified, its two pybind11 helpers built;
put there by oracle/build_ref.py -- test infrastructure, git-ignored, shipped with the repo
snapshot).  With it on ``sys.path`` the B200 classes subclass the reference's ``Benchmark`` /
abstract stencils (stencil_benchmarks_b200/benchmark.py), so here the reference's own code
drives and judges the GPU kernels:

* ``Stencil.run()`` -> our ``run_stencil`` -> the reference's ``verify_stencil`` + ``check_equality``
  (stencil_benchmarks/benchmarks_collection/stencils/base.py:151-166, tools/validation.py:104-116);
* ``cli.main([...])`` as in the reference's own CLI test (test/test_cli.py:38-54), CSV output
  included (cli.py:86-116);
* the collection script through the reference's ``tools.multirun`` (multirun.py:75-92).

Every case runs in a fresh interpreter (this pytest process has already imported the stand-alone
mirror of the plugin API) in a scratch directory.
"""

import json
import os
import pathlib
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).parent.parent.resolve()
PACKAGE = ROOT / "oracle" / "_ref" / "pkg"

pytestmark = [
    pytest.mark.gpu,
    pytest.mark.skipif(not (PACKAGE / "stencil_benchmarks").exists(),
                       reason="oracle/_ref/pkg is not built (oracle/build_ref.py, dev container)"),
]


def run(code, tmp_path, *argv, timeout=600):
    env = dict(os.environ, PYTHONPATH=f"{PACKAGE}:{ROOT}")
    return subprocess.run([sys.executable, "-c", code, *argv], capture_output=True, text=True, env=env,
                          timeout=timeout, cwd=str(tmp_path))


VERIFY = """
import json, sys
import stencil_benchmarks.benchmark as ref
import stencil_benchmarks.benchmarks_collection.stencils.base as ref_base
import stencil_benchmarks.tools.validation as validation
import stencil_benchmarks_b200.benchmark as ours
from stencil_benchmarks_b200 import capi
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import (
    basic, horizontal_diffusion, vertical_advection)

assert ours.HAVE_REFERENCE and ours.Benchmark is ref.Benchmark
checked = []
original = validation.check_equality
def counting(name, result, expected):
    checked.append(name)
    return original(name, result, expected)
validation.check_equality = counting

cases = json.loads(sys.argv[1])
classes = dict(copy=basic.Copy, onesided=basic.OnesidedAverage, symmetric=basic.SymmetricAverage,
               laplacian=basic.Laplacian, hdiff=horizontal_diffusion.Fused, vadv=vertical_advection.Thomas)
for name, kwargs in cases:
    cls = classes[name]
    kwargs = {k: tuple(v) if isinstance(v, list) else v for k, v in kwargs.items()}
    bench = cls(verify=True, **kwargs)
    assert cls.run is ref_base.Stencil.run, "the reference's run() must drive the sweep"
    before = len(checked)
    launches = capi.launch_count()
    result = bench.run()
    # (a library compiled at setup() counts its launches itself, not in the prebuilt one)
    assert "compiler" in kwargs or capi.launch_count() > launches, "no sb200 kernel was launched"
    assert len(checked) > before, "the reference's check_equality was not reached"
    print(name, kwargs, f"{result['time'] * 1e6:.1f} us", f"{result['bandwidth']:.1f} GB/s", flush=True)
print("verified", len(checked), "fields")
"""


def test_reference_run_and_oracle_judge_the_gpu_kernels(tmp_path):
    """Fused / Thomas / Laplacian / Copy / averages with verify=True: the reference's NumPy oracle
    validates what the B200 kernels wrote, at its own tolerances."""
    cases = [
        ("copy", dict(domain=(64, 48, 20))),
        ("copy", dict(domain=(100, 31, 7), dtype="float32", halo=(0, 0, 0))),
        ("onesided", dict(domain=(70, 33, 9), axis=1)),
        ("symmetric", dict(domain=(70, 33, 9), axis=2, dtype="float32")),
        ("laplacian", dict(domain=(128, 128, 80))),
        ("laplacian", dict(domain=(65, 47, 11), along_z=True, dtype="float32")),
        ("hdiff", dict(domain=(128, 128, 80))),                      # BASELINE.json configs[0] size
        ("hdiff", dict(domain=(1030, 75, 6), dtype="float32")),     # TMA path, ragged tile
        ("hdiff", dict(domain=(40, 40, 3), alignment=0)),           # unaligned: generic kernel
        ("vadv", dict(domain=(128, 128, 80))),
        ("vadv", dict(domain=(200, 17, 160))),                      # BASELINE level count
        ("vadv", dict(domain=(96, 24, 40), all_components=True)),
        ("vadv", dict(domain=(33, 9, 12), coefficients="global")),
        # kernels compiled at setup() through the reference's tools.compilation.GnuLibrary
        ("hdiff", dict(domain=(300, 40, 5), compiler="nvcc")),
        ("vadv", dict(domain=(130, 20, 30), compiler="nvcc")),
    ]
    result = run(VERIFY, tmp_path, json.dumps(cases))
    assert result.returncode == 0, result.stdout + result.stderr
    assert "verified" in result.stdout


def test_reference_oracle_rejects_a_wrong_sweep(tmp_path):
    """Negative control: if the kernel is not run, the reference's validation must fail -- the
    check above has teeth."""
    code = """
import stencil_benchmarks.tools.validation as validation
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion
bench = horizontal_diffusion.Fused(domain=(64, 32, 4), verify=True)
import ctypes
def fake(pointers, dry_runs, time_ptr, stream, domain=None):
    ctypes.cast(time_ptr, ctypes.POINTER(ctypes.c_double))[0] = 1e-6
bench.launch = fake
try:
    bench.run()
except validation.ValidationError as error:
    print("rejected:", str(error).splitlines()[0])
else:
    raise SystemExit("a sweep that was never run passed validation")
"""
    result = run(code, tmp_path)
    assert result.returncode == 0, result.stdout + result.stderr
    assert "rejected: validation of field out failed" in result.stdout


CLI = """
import sys
import stencil_benchmarks.benchmarks_collection
import stencil_benchmarks_b200.benchmarks_collection
from stencil_benchmarks import cli
cli.main(args=sys.argv[1:], standalone_mode=False)
"""


def test_reference_cli_runs_the_backend(tmp_path):
    """`sbench ... stencils b200 horizontal-diffusion fused` and `stream b200 native` through the
    reference's click tree, with its range syntax and CSV writer."""
    import pandas as pd

    out = tmp_path / "hdiff.csv"
    result = run(CLI, tmp_path, "--executions", "3", "--output", str(out), "stencils", "b200",
                 "horizontal-diffusion", "fused", "--domain", "128", "128", "80", "--dtype",
                 "[float32,float64]")
    assert result.returncode == 0, result.stdout + result.stderr
    table = pd.read_csv(out)
    assert len(table) == 6 and set(table["dtype"]) == {"float32", "float64"}
    assert (table["bandwidth"] > 0).all() and (table["verify"]).all()
    assert {"time", "bandwidth-algorithmic", "alignment", "sbench-version"} <= set(table.columns)

    out = tmp_path / "vadv.csv"
    result = run(CLI, tmp_path, "--executions", "2", "--output", str(out), "stencils", "b200",
                 "vertical-advection", "thomas", "--domain", "128", "128", "80")
    assert result.returncode == 0, result.stdout + result.stderr
    assert len(pd.read_csv(out)) == 2

    out = tmp_path / "basic.csv"
    result = run(CLI, tmp_path, "--executions", "2", "--output", str(out), "stencils", "b200", "basic",
                 "laplacian", "--domain", "10", "10", "10", "--along-z")
    assert result.returncode == 0, result.stdout + result.stderr

    out = tmp_path / "partitioned.csv"
    result = run(CLI, tmp_path, "--executions", "2", "--output", str(out), "stencils", "b200",
                 "horizontal-diffusion", "partitioned", "--domain", "300", "64", "5", "--gpus", "1")
    assert result.returncode == 0, result.stdout + result.stderr
    assert list(pd.read_csv(out)["gpus"]) == [1, 1]

    out = tmp_path / "stream.csv"
    result = run(CLI, tmp_path, "--executions", "2", "--output", str(out), "stream", "b200", "native",
                 "--array-size", "4194304")
    assert result.returncode == 0, result.stdout + result.stderr
    table = pd.read_csv(out)
    assert sorted(set(table["name"])) == ["add", "copy", "scale", "triad"]
    assert (table["bandwidth"] > 1e5).all()  # MB/s, as the reference reports STREAM


def test_cli_exit_codes_follow_the_reference(tmp_path):
    """Invalid parameters are skippable ParameterErrors (cli.py:53-61): with `-s` the int32 member
    of the range is skipped and the float64 one runs."""
    import pandas as pd

    out = tmp_path / "skip.csv"
    result = run(CLI, tmp_path, "-s", "--executions", "1", "--output", str(out), "stencils", "b200",
                 "horizontal-diffusion", "fused", "--domain", "16", "16", "4", "--dtype", "[int32,float64]")
    assert result.returncode == 0, result.stdout + result.stderr
    assert list(pd.read_csv(out)["dtype"]) == ["float64"]


def test_collection_script_writes_the_reference_csv(tmp_path):
    """One family of sbench_b200_collection.py through the reference's multirun tool."""
    import pandas as pd

    out = tmp_path / "collection.csv"
    code = ("import sys; from stencil_benchmarks_b200.scripts.sbench_b200_collection import main; "
            "main(args=sys.argv[1:], standalone_mode=False)")
    result = run(code, tmp_path, "horizontal-diffusion-bandwidth", str(out), "--executions", "3",
                 "--option", "dtype='float32'", timeout=900)
    assert result.returncode == 0, result.stdout + result.stderr
    table = pd.read_csv(out)
    assert len(table) == 3 * 7  # 3 executions x domains 32^2 ... 2048^2 (x 80)
    assert (table["bandwidth"] > 0).all() and set(table["dtype"]) == {"float32"}
    assert set(table["name"]) == {"fused"}


def test_partitioned_class_is_verified_by_the_reference_oracle(tmp_path):
    """All GPUs of the box (at least 2): the gathered global field passes the reference's
    verify_stencil."""
    import sys as _sys
    _sys.path.insert(0, str(ROOT))
    from stencil_benchmarks_b200 import capi

    gpus = min(capi.device_count(), 8)
    if gpus < 2:
        pytest.skip("needs two GPUs")
    code = """
import sys
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion
for dtype in ("float64", "float32"):
    bench = horizontal_diffusion.Partitioned(domain=(520, 264, 6), dtype=dtype, gpus=int(sys.argv[1]), verify=True)
    print(dtype, bench.run()["bandwidth"])
print("verified")
"""
    result = run(code, tmp_path, str(gpus))
    assert result.returncode == 0, result.stdout + result.stderr
    assert "verified" in result.stdout
