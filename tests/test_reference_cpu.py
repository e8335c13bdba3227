"""The reference's own compiled OpenMP kernels (oracle/_ref) agree with the oracle.

This validates the oracle restatement against real reference code and
pins BASELINE.json configs[0]: openmp horizontal-diffusion, 128x128x80 float64.
Skipped where oracle/_ref has not been built (it is built in the dev container only).
"""

import numpy as np
import pytest

from oracle import ref_cpu, stencils

pytestmark = pytest.mark.skipif(not ref_cpu.available(), reason="oracle/_ref not built")


def inner(entry):
    return tuple(slice(h, h + d) for d, h in zip(entry["domain"], entry["halo"]))


@pytest.mark.parametrize("name", ["hdiff_otf_128x128x80_f64", "hdiff_otfvec_128x128x80_f64",
                                  "hdiff_rolling_128x128x80_f64"])
def test_reference_openmp_hdiff_matches_oracle(name):
    kernel = ref_cpu.Kernel(name)
    inp, coeff, out = kernel.fields(seed=4)
    inp0, coeff0 = inp.copy(), coeff.copy()
    assert kernel([inp, coeff, out]) > 0
    expected = stencils.hdiff(inp0, coeff0)
    sl = inner(kernel.entry)
    np.testing.assert_allclose(out[sl], expected[sl], rtol=1e-5, atol=1e-8)
    assert np.array_equal(inp, inp0) and np.array_equal(coeff, coeff0)


def test_reference_openmp_vadv_matches_oracle():
    kernel = ref_cpu.Kernel("vadv_kinnermost_128x128x80_f64")
    fields = kernel.fields(seed=9)
    before = [f.copy() for f in fields]
    kernel(fields)
    names = kernel.entry["args"]
    ustage, upos, utens, utensstage, wcon = (before[names.index(n)] for n in
                                             ("ustage", "upos", "utens", "utensstage", "wcon"))
    expected = stencils.vadv(ustage, upos, utens, utensstage, wcon, tuple(kernel.entry["halo"]))
    sl = inner(kernel.entry)
    np.testing.assert_allclose(fields[names.index("utensstage")][sl], expected[sl], rtol=1e-5, atol=1e-8)


def test_manifest_describes_baseline_configs():
    entries = ref_cpu.manifest()
    assert entries["hdiff_otf_128x128x80_f64"]["domain"] == [128, 128, 80]
    assert entries["hdiff_otf_128x128x80_f64"]["data_size"] == 32680448
    assert "hdiff_otfvec_2048x2048x80_f64" in entries and "vadv_kmiddlevec_1024x1024x160_f64" in entries


def test_bench_times_the_other_configs_on_the_host(monkeypatch):
    """bench.py's `cpu_baseline.also`: best OpenMP variant per config with algorithmic bytes;
    the primary workload's own config is left to `cpu_baseline.value`, a missing kernel is skipped."""
    import bench

    monkeypatch.setattr(bench, "CPU_OTHER_CONFIGS", [
        bench.CPU_OTHER_CONFIGS[0],
        ("hdiff_2048x2048x80_f64", ["hdiff_rolling_2048x2048x80_f64"], "hdiff", (2048, 2048, 80)),
        ("not built", ["no_such_kernel"], "vadv", (8, 8, 8)),
    ])
    result = bench.cpu_other_configs("hdiff", budget_s=0.5)
    label = bench.CPU_OTHER_CONFIGS[0][0]
    assert set(result) == {label, "cores"}
    entry = result[label]
    assert entry["variants_tried"] == 3 and entry["sweeps"] >= 1 and entry["ms"] > 0
    nbytes = (2 * 128 * 128 * 80 + 132 * 132 * 80) * 8
    assert entry["gbs"] == pytest.approx(nbytes / (entry["ms"] * 1e-3) / 1e9)
