"""Generate the golden vectors in tests/golden/ from the reference's own NumPy oracle.

Run in the dev container (the reference is not available on the GPU box):

    cp -r /root/reference /tmp/ref && (cd /tmp/ref && python setup.py build_ext --inplace)
    PYTHONPATH=/tmp/ref python tests/golden/make_golden.py

For each case a tiny probe class derives from the reference's abstract stencil
(stencil_benchmarks/benchmarks_collection/stencils/base.py) -- so the reference's
``verify_stencil`` runs unmodified -- seeded inputs replace its random fields, and
``validation.check_equality`` is intercepted to capture the ``expected`` array
the reference compares every backend against.  Inputs + expected outputs are
stored as ``<case>.npz``.  Nothing of this repository's oracle or kernels is
involved in producing the files.
"""

import pathlib
import sys

import numpy as np

from stencil_benchmarks.benchmarks_collection.stencils import base
from stencil_benchmarks.tools import validation

OUT = pathlib.Path(__file__).parent.resolve()


def seeded_fill(bench, seed):
    """Overwrite every (padded) field of the benchmark with reproducible U[0,1) values."""
    data = bench._data[0]
    for index, field in enumerate(data):
        rng = np.random.default_rng([seed, index])
        field[...] = rng.random(field.shape, dtype=np.float64).astype(field.dtype)
    return data


def capture_expected(bench, data):
    """Run the reference's verify_stencil; return {field name: expected interior array}."""
    captured = {}
    original = validation.check_equality

    def recorder(name, result, expected):
        captured[name] = np.array(expected, copy=True)

    validation.check_equality = recorder
    try:
        before = type(data)(*(np.array(f, copy=True) for f in data))
        after = type(data)(*(np.array(f, copy=True) for f in data))
        bench.verify_stencil(before, after)
    finally:
        validation.check_equality = original
    return captured


def probe(stencil_class):
    class Probe(stencil_class):
        def run_stencil(self, data):
            return dict(time=1.0)

    return Probe


CASES = []
for dtype in ("float64", "float32"):
    tag = "f64" if dtype == "float64" else "f32"
    common = dict(dtype=dtype, verify=True)
    CASES += [
        (f"copy_{tag}", base.CopyStencil, dict(domain=(11, 7, 5), halo=(3, 3, 3), **common), ["out"]),
        (f"copy_h0_{tag}", base.CopyStencil, dict(domain=(9, 6, 4), halo=(0, 0, 0), **common), ["out"]),
    ]
    for axis in range(3):
        CASES += [
            (f"onesided_ax{axis}_{tag}", base.OnesidedAverageStencil,
             dict(domain=(11, 7, 5), halo=(1, 2, 3), axis=axis, **common), ["out"]),
            (f"symmetric_ax{axis}_{tag}", base.SymmetricAverageStencil,
             dict(domain=(10, 9, 6), halo=(3, 1, 2), axis=axis, **common), ["out"]),
        ]
    for mask in range(1, 8):
        CASES.append(
            (f"laplacian_m{mask}_{tag}", base.LaplacianStencil,
             dict(domain=(12, 8, 6), halo=(2, 1, 1), along_x=bool(mask & 1), along_y=bool(mask & 2),
                  along_z=bool(mask & 4), **common), ["out"]))
    CASES += [
        (f"hdiff_{tag}", base.HorizontalDiffusionStencil,
         dict(domain=(13, 10, 6), halo=(3, 3, 3), **common), ["out"]),
        (f"hdiff_h2_{tag}", base.HorizontalDiffusionStencil,
         dict(domain=(17, 9, 3), halo=(2, 2, 0), alignment=64, **common), ["out"]),
        (f"vadv_{tag}", base.VerticalAdvectionStencil,
         dict(domain=(9, 7, 12), halo=(3, 3, 3), **common), ["utensstage"]),
        (f"vadv_h1_{tag}", base.VerticalAdvectionStencil,
         dict(domain=(8, 5, 2), halo=(1, 0, 1), **common), ["utensstage"]),
        (f"vadv_all_{tag}", base.VerticalAdvectionStencil,
         dict(domain=(7, 6, 9), halo=(1, 1, 1), all_components=True, **common),
         ["utensstage", "vtensstage", "wtensstage"]),
    ]


def main():
    for seed, (name, stencil_class, kwargs, outputs) in enumerate(CASES):
        bench = probe(stencil_class)(**kwargs)
        data = seeded_fill(bench, 1000 + seed)
        expected = capture_expected(bench, data)
        arrays = {"in_" + field_name: np.ascontiguousarray(field)
                  for field_name, field in zip(bench.args, data)}
        for output in outputs:
            arrays["expected_" + output] = np.ascontiguousarray(expected[output])
        arrays["halo"] = np.array(bench.halo)
        arrays["domain"] = np.array(bench.domain)
        np.savez_compressed(OUT / f"{name}.npz", **arrays)
        print(f"{name}: {', '.join(outputs)}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
