"""GPU parity: every B200 benchmark class, called through the plugin API and the C ABI,
against the oracle (oracle/) and the golden vectors captured from the reference.

Mirrors the reference's only numerical test, which instantiates every registered class
and runs it with verification (stencil_benchmarks/test/benchmarks_collection/
test_benchmarks_collection.py:39-55) -- here the verification is the oracle, inputs are
seeded, and odd sizes / both dtypes / aligned and unaligned layouts are covered.

Tolerances are the reference's (stencil_benchmarks/tools/validation.py:98-106):
float64 rtol=1e-5 atol=1e-8, float32 rtol=1e-4 atol=1e-5.
"""

import glob
import pathlib

import numpy as np
import pytest

from oracle import native, stencils
from stencil_benchmarks_b200 import capi
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import (
    basic,
    horizontal_diffusion,
    vertical_advection,
)

pytestmark = pytest.mark.gpu

GOLDEN = pathlib.Path(__file__).parent / "golden"
CASES = sorted(pathlib.Path(p).stem for p in glob.glob(str(GOLDEN / "*.npz")))


def snapshot(bench):
    return {name: np.array(field, copy=True) for name, field in zip(bench.args, bench.data())}


def close(result, expected, dtype):
    return np.allclose(result, expected, **stencils.tolerances(dtype))


# float32 vertical advection: the reference's own OpenMP backend fails the reference tolerance on
# 0.026-0.03 % of the points of U[0,1) inputs (344-702 of 1.3-2.6 M); the B200
# kernels may miss it on at most ten times that fraction (and never on more than a handful of
# points of a tiny domain)
VADV_F32_MISMATCH = 3e-3


def f32_vadv_ok(bad):
    return bad.sum() <= max(3, VADV_F32_MISMATCH * bad.size)


def check_inputs_untouched(bench, before, written):
    after = bench.data()
    for name, field in zip(bench.args, after):
        role = bench.field_roles.get(name, "inout")
        if name not in written and role != "scratch":
            assert np.array_equal(before[name], field), f"input field {name} was modified"


# --------------------------------------------------------------------------------------
# golden vectors
# --------------------------------------------------------------------------------------
def class_for(case):
    kind = case.split("_")[0]
    if kind == "copy":
        return basic.Copy, {}
    if kind == "onesided":
        return basic.OnesidedAverage, dict(axis=int(case.split("_ax")[1][0]))
    if kind == "symmetric":
        return basic.SymmetricAverage, dict(axis=int(case.split("_ax")[1][0]))
    if kind == "laplacian":
        mask = int(case.split("_m")[1][0])
        return basic.Laplacian, dict(along_x=bool(mask & 1), along_y=bool(mask & 2),
                                     along_z=bool(mask & 4))
    if kind == "hdiff":
        return horizontal_diffusion.Fused, {}
    return vertical_advection.Thomas, dict(all_components="_all_" in case)


@pytest.mark.parametrize("alignment", [128, 0])
@pytest.mark.parametrize("case", CASES)
def test_golden(case, alignment):
    with np.load(GOLDEN / f"{case}.npz") as data:
        g = {key: data[key] for key in data.files}
    dtype = "float32" if case.endswith("f32") else "float64"
    if alignment % np.dtype(dtype).itemsize:
        pytest.skip("alignment not a multiple of the item size")
    cls, extra = class_for(case)
    bench = cls(domain=tuple(int(d) for d in g["domain"]), halo=tuple(int(h) for h in g["halo"]),
                dtype=dtype, alignment=alignment, verify=False, **extra)
    for name, field in zip(bench.args, bench.data()):
        field[...] = g["in_" + name]
    before = snapshot(bench)
    bench.run()
    after = bench.data()
    outputs = [key[len("expected_"):] for key in g if key.startswith("expected_")]
    for name in outputs:
        result = getattr(after, name)[bench.inner_slice()]
        expected = g["expected_" + name]
        if case.startswith("vadv") and dtype == "float32":
            # float32 vadv is ill-conditioned on U[0,1) inputs; the reference's own
            # OpenMP backend misses its tolerance on about 0.03 % of the points.  Policy: at most ten
            # times that fraction (VADV_F32_MISMATCH).
            bad = ~np.isclose(result, expected, **stencils.tolerances(dtype))
            assert f32_vadv_ok(bad), f"{case}:{name}: {bad.sum()} of {bad.size} points differ"
        else:
            assert close(result, expected, dtype), (
                f"{case}:{name}: max abs err {np.abs(result - expected).max()}")
    check_inputs_untouched(bench, before, outputs)


# --------------------------------------------------------------------------------------
# seeded random fields vs the NumPy oracle, odd sizes
# --------------------------------------------------------------------------------------
DOMAINS = [((10, 10, 10), (3, 3, 3)), ((33, 17, 5), (2, 2, 1)), ((128, 128, 80), (3, 3, 3)),
           ((257, 70, 3), (2, 3, 0))]


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("alignment", [128, 0])
@pytest.mark.parametrize("domain,halo", DOMAINS)
def test_hdiff_random(domain, halo, dtype, alignment):
    bench = horizontal_diffusion.Fused(domain=domain, halo=halo, dtype=dtype, alignment=alignment,
                                       verify=False, seed=11)
    before = snapshot(bench)
    result = bench.run()
    assert result["time"] > 0 and result["bandwidth"] > 0
    expected = stencils.hdiff(before["inp"], before["coeff"])
    inner = bench.inner_slice()
    out = bench.data().out[inner]
    assert close(out, expected[inner], dtype)
    if dtype == "float64":
        # same operation order as the oracle: float64 agrees to the last bits
        np.testing.assert_allclose(out, expected[inner], rtol=1e-14, atol=1e-15)
    check_inputs_untouched(bench, before, ["out"])
    # only interior points of out are written
    mask = np.ones(before["out"].shape, bool)
    mask[inner] = False
    assert np.array_equal(bench.data().out[mask], before["out"][mask])


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("alignment", [128, 0])
@pytest.mark.parametrize("domain,halo", DOMAINS)
def test_basic_random(domain, halo, dtype, alignment):
    cases = [(basic.Copy, {}, lambda f: stencils.copy(f, halo))]
    for axis in range(3):
        if halo[axis] >= 1:
            cases.append((basic.OnesidedAverage, dict(axis=axis),
                          lambda f, a=axis: stencils.onesided_average(f, halo, a)))
            cases.append((basic.SymmetricAverage, dict(axis=axis),
                          lambda f, a=axis: stencils.symmetric_average(f, halo, a)))
    for mask in range(1, 8):
        along = (bool(mask & 1), bool(mask & 2), bool(mask & 4))
        if all(h >= 1 for h, a in zip(halo, along) if a):
            cases.append((basic.Laplacian, dict(along_x=along[0], along_y=along[1], along_z=along[2]),
                          lambda f, al=along: stencils.laplacian(f, halo, al)))
    for cls, extra, oracle in cases:
        bench = cls(domain=domain, halo=halo, dtype=dtype, alignment=alignment, verify=False,
                    seed=5, **extra)
        before = snapshot(bench)
        bench.run()
        inner = bench.inner_slice()
        expected = oracle(before["inp"])[inner]
        out = bench.data().out[inner]
        assert close(out, expected, dtype), f"{cls.__name__} {extra}"
        check_inputs_untouched(bench, before, ["out"])


def test_empty_runs():
    bench = basic.Empty(domain=(64, 64, 8), verify=False)
    assert bench.run()["time"] > 0


@pytest.mark.parametrize("all_components", [False, True])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("domain,halo", [((10, 10, 10), (3, 3, 3)), ((33, 17, 2), (1, 1, 1)),
                                         ((128, 128, 80), (3, 3, 3)), ((65, 9, 160), (1, 1, 1))])
def test_vadv_random(domain, halo, dtype, all_components):
    bench = vertical_advection.Thomas(domain=domain, halo=halo, dtype=dtype, verify=False, seed=3,
                                      all_components=all_components)
    before = snapshot(bench)
    bench.run()
    inner = bench.inner_slice()
    components = "uvw" if all_components else "u"
    shifts = {"u": (1, 0), "v": (0, 1), "w": (0, 0)}
    for c in components:
        expected = stencils._vadv_component(
            before[c + "stage"], before[c + "pos"], before[c + "tens"], before[c + "tensstage"],
            before["wcon"], halo, *shifts[c])[inner]
        out = getattr(bench.data(), c + "tensstage")[inner]
        bad = ~np.isclose(out, expected, **stencils.tolerances(dtype))
        if dtype == "float64":
            assert not bad.any(), f"{c}: {bad.sum()} of {bad.size} points differ"
        else:  # float32 policy, SURVEY.md §8c: ill-conditioned columns, tiny mismatch fraction allowed
            assert f32_vadv_ok(bad), f"{c}: {bad.sum()} of {bad.size} points differ"
    check_inputs_untouched(bench, before, [c + "tensstage" for c in components])


# --------------------------------------------------------------------------------------
# BASELINE.json sizes: C oracle on the host cores (NumPy would need ~30 GB of temporaries)
# --------------------------------------------------------------------------------------
def test_hdiff_full_size():
    bench = horizontal_diffusion.Fused(domain=(2048, 2048, 80), dtype="float64", verify=False, seed=1)
    data = bench.data()
    result = bench.run()
    expected = bench.empty_field()  # same strides as the benchmark's fields
    native.hdiff(data.inp, data.coeff, expected, bench.halo)
    inner = bench.inner_slice()
    for k in range(0, 80, 8):  # compare plane by plane to keep temporaries small
        sl = (inner[0], inner[1], slice(inner[2].start + k, inner[2].start + k + 8))
        np.testing.assert_allclose(data.out[sl], expected[sl], rtol=1e-13, atol=1e-14)
    assert result["bandwidth-algorithmic"] > 100  # GB/s: it did run on the GPU


def test_vadv_full_size():
    bench = vertical_advection.Thomas(domain=(1024, 1024, 160), dtype="float64", verify=False, seed=2)
    data = bench.data()
    expected, scratch_c, scratch_d = (bench.empty_field() for _ in range(3))  # same strides
    expected[...] = data.utensstage
    native.vadv(data.ustage, data.upos, data.utens, expected, data.wcon, scratch_c, scratch_d,
                bench.halo)
    bench.run()
    inner = bench.inner_slice()
    bad = 0
    for k in range(0, 160, 16):
        sl = (inner[0], inner[1], slice(inner[2].start + k, inner[2].start + k + 16))
        bad += np.count_nonzero(~np.isclose(data.utensstage[sl], expected[sl], rtol=1e-5, atol=1e-8))
    assert bad == 0


# --------------------------------------------------------------------------------------
# pipelined execution (chunks > 1): slab-wise upload / sweep / download on three streams
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("chunks", [2, 5])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_pipelined_hdiff(chunks, dtype):
    bench = horizontal_diffusion.Fused(domain=(300, 67, 6), halo=(3, 3, 2), dtype=dtype, verify=False,
                                       seed=17, chunks=chunks)
    before = snapshot(bench)
    result = bench.run()
    assert result["time"] > 0 and result["time-end-to-end"] >= result["time"]
    inner = bench.inner_slice()
    expected = stencils.hdiff(before["inp"], before["coeff"])[inner]
    assert close(bench.data().out[inner], expected, dtype)
    check_inputs_untouched(bench, before, ["out"])
    mask = np.ones(before["out"].shape, bool)
    mask[inner] = False
    assert np.array_equal(bench.data().out[mask], before["out"][mask])
    # a second run on the same instance reuses streams, events and device mirrors
    bench.run()
    assert close(bench.data().out[inner], expected, dtype)


@pytest.mark.parametrize("all_components", [False, True])
def test_pipelined_vadv(all_components):
    halo = (1, 1, 1)
    bench = vertical_advection.Thomas(domain=(200, 37, 24), halo=halo, verify=False, seed=8, chunks=4,
                                      all_components=all_components)
    before = snapshot(bench)
    bench.run()
    inner = bench.inner_slice()
    shifts = {"u": (1, 0), "v": (0, 1), "w": (0, 0)}
    for c in ("uvw" if all_components else "u"):
        expected = stencils._vadv_component(
            before[c + "stage"], before[c + "pos"], before[c + "tens"], before[c + "tensstage"],
            before["wcon"], halo, *shifts[c])[inner]
        assert close(getattr(bench.data(), c + "tensstage")[inner], expected, "float64"), c
    check_inputs_untouched(bench, before, [c + "tensstage" for c in "uvw"])


def test_pipelined_basic():
    halo = (1, 1, 1)
    for extra, oracle in [
        (dict(along_x=True, along_y=True, along_z=True), lambda f: stencils.laplacian(f, halo, (1, 1, 1))),
        (dict(along_x=False, along_y=True, along_z=False), lambda f: stencils.laplacian(f, halo, (0, 1, 0))),
    ]:
        bench = basic.Laplacian(domain=(130, 50, 9), halo=halo, verify=False, seed=2, chunks=3, **extra)
        before = snapshot(bench)
        bench.run()
        inner = bench.inner_slice()
        assert close(bench.data().out[inner], oracle(before["inp"])[inner], "float64")
    bench = basic.OnesidedAverage(domain=(64, 33, 5), halo=halo, axis=1, verify=False, chunks=4)
    before = snapshot(bench)
    bench.run()
    inner = bench.inner_slice()
    assert close(bench.data().out[inner], stencils.onesided_average(before["inp"], halo, 1)[inner], "float64")


def test_resident_mode_and_wall_timer():
    """resident=True uploads once and leaves results on the device; repeated in-place sweeps then
    compose (vadv's utensstage is in/out), and the wall-clock timer reports a positive time."""
    halo = (1, 1, 1)
    bench = vertical_advection.Thomas(domain=(96, 20, 12), halo=halo, verify=False, seed=4, resident=True,
                                      timers="wall")
    before = snapshot(bench)
    assert bench.run()["time"] > 0
    bench.run()
    mirrors = bench._device_fields(bench.data())
    bench.download(bench.data(), mirrors)
    once = stencils.vadv(before["ustage"], before["upos"], before["utens"], before["utensstage"],
                         before["wcon"], halo)
    twice = stencils.vadv(before["ustage"], before["upos"], before["utens"], once, before["wcon"], halo)
    inner = bench.inner_slice()
    assert close(bench.data().utensstage[inner], twice[inner], "float64")


def test_data_sets_cycle():
    bench = basic.Copy(domain=(40, 12, 6), verify=False, data_sets=2, seed=9)
    first, second = bench.data(0), bench.data(1)
    assert not np.array_equal(first.inp, second.inp)
    bench.run()
    bench.run()
    inner = bench.inner_slice()
    assert np.array_equal(first.out[inner], first.inp[inner])
    assert np.array_equal(second.out[inner], second.inp[inner])


# --------------------------------------------------------------------------------------
# kernel selection: every variant must give the same answer, and fall back where it cannot run
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("nz", [33, 200])  # 200: part of the store spills from TMEM to smem
@pytest.mark.parametrize("coefficients", ["auto", "global", "onchip"])
def test_vadv_variants_agree(coefficients, nz):
    halo = (2, 2, 2)
    bench = vertical_advection.Thomas(domain=(200, 11, nz), halo=halo, verify=False, seed=6,
                                      coefficients=coefficients)
    before = snapshot(bench)
    bench.run()
    expected = stencils.vadv(before["ustage"], before["upos"], before["utens"], before["utensstage"],
                             before["wcon"], halo)
    inner = bench.inner_slice()
    assert close(bench.data().utensstage[inner], expected[inner], "float64")


def test_vadv_tall_columns_fall_back_to_global_kernel():
    """nz = 300 float64 exceeds the on-chip store (2 x 299 TMEM columns > 512)."""
    halo = (1, 1, 1)
    bench = vertical_advection.Thomas(domain=(128, 6, 300), halo=halo, verify=False, seed=12)
    before = snapshot(bench)
    bench.run()
    expected = stencils.vadv(before["ustage"], before["upos"], before["utens"], before["utensstage"],
                             before["wcon"], halo)
    inner = bench.inner_slice()
    assert close(bench.data().utensstage[inner], expected[inner], "float64")
    forced = vertical_advection.Thomas(domain=(128, 6, 300), halo=halo, verify=False, coefficients="onchip")
    from stencil_benchmarks_b200 import benchmark

    with pytest.raises(benchmark.ExecutionError):
        forced.run()


def test_vadv_onchip_needs_aligned_fields():
    from stencil_benchmarks_b200 import benchmark

    bench = vertical_advection.Thomas(domain=(131, 7, 9), halo=(1, 1, 1), alignment=0, verify=False,
                                      coefficients="onchip")
    with pytest.raises(benchmark.ExecutionError):
        bench.run()


@pytest.mark.parametrize("nx", [60, 127, 128, 129, 255, 256, 257, 511, 513])
def test_hdiff_tile_edges(nx):
    """Widths around the TMA tile (256 doubles) and the TMA / generic switch (128 doubles)."""
    bench = horizontal_diffusion.Fused(domain=(nx, 21, 3), halo=(2, 2, 0), verify=False, seed=nx)
    before = snapshot(bench)
    bench.run()
    inner = bench.inner_slice()
    expected = stencils.hdiff(before["inp"], before["coeff"])[inner]
    np.testing.assert_allclose(bench.data().out[inner], expected, rtol=1e-14, atol=1e-15)


@pytest.mark.parametrize("ny", [1, 2, 3, 5, 127, 129])
def test_hdiff_short_marches(ny):
    """Row counts around the pipeline depth (4 rows per stage) and the march length."""
    for dtype in ("float64", "float32"):
        bench = horizontal_diffusion.Fused(domain=(260, ny, 2), halo=(2, 2, 1), dtype=dtype, verify=False,
                                           seed=ny)
        before = snapshot(bench)
        bench.run()
        inner = bench.inner_slice()
        expected = stencils.hdiff(before["inp"], before["coeff"])[inner]
        assert close(bench.data().out[inner], expected, dtype)


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("kind", ["copy", "onesided_i", "onesided_k", "symmetric_j", "laplacian_ij", "laplacian_ijk"])
def test_basic_full_size(kind, dtype):
    """BASELINE.json configs[2]: the basic stencils at 1024x1024x80, both dtypes, against the C oracle."""
    common = dict(domain=(1024, 1024, 80), halo=(1, 1, 1), dtype=dtype, verify=False, seed=3)
    expected_call = {
        "copy": (basic.Copy, {}, lambda i, o, h: native.copy(i, o, h)),
        "onesided_i": (basic.OnesidedAverage, dict(axis=0), lambda i, o, h: native.average(i, o, h, 0, False)),
        "onesided_k": (basic.OnesidedAverage, dict(axis=2), lambda i, o, h: native.average(i, o, h, 2, False)),
        "symmetric_j": (basic.SymmetricAverage, dict(axis=1), lambda i, o, h: native.average(i, o, h, 1, True)),
        "laplacian_ij": (basic.Laplacian, {}, lambda i, o, h: native.laplacian(i, o, h, (True, True, False))),
        "laplacian_ijk": (basic.Laplacian, dict(along_z=True),
                          lambda i, o, h: native.laplacian(i, o, h, (True, True, True))),
    }
    cls, kwargs, oracle = expected_call[kind]
    bench = cls(**common, **kwargs)
    data = bench.data()
    bench.run()
    expected = bench.empty_field()
    oracle(data.inp, expected, tuple(bench.halo))
    inner = bench.inner_slice()
    assert close(data.out[inner], expected[inner], dtype)


@pytest.mark.parametrize("all_components", [False, True])
def test_vadv_float32_wide_batches(all_components):
    """float32 runs 256-column batches (8 warps, split wcon box): several batches per row, the last
    one partial, the i+1 neighbour of a batch's last column coming from the 16-byte edge box."""
    halo = (1, 1, 1)
    bench = vertical_advection.Thomas(domain=(600, 5, 40), halo=halo, dtype="float32", verify=False, seed=31,
                                      all_components=all_components, coefficients="onchip")
    before = snapshot(bench)
    bench.run()
    inner = bench.inner_slice()
    shifts = {"u": (1, 0), "v": (0, 1), "w": (0, 0)}
    for c in ("uvw" if all_components else "u"):
        expected = stencils._vadv_component(
            before[c + "stage"], before[c + "pos"], before[c + "tens"], before[c + "tensstage"],
            before["wcon"], halo, *shifts[c])[inner]
        out = getattr(bench.data(), c + "tensstage")[inner]
        bad = ~np.isclose(out, expected, **stencils.tolerances("float32"))
        assert bad.mean() < 1e-2, f"{c}: {bad.sum()} of {bad.size} points differ"
        # the columns next to a batch border (i = 255, 256, 511, 512) must be as good as the rest
        for i in (255, 256, 511, 512):
            assert bad[i].mean() < 5e-2, f"{c}: column block i={i} differs"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,domain,steps", [("float64", (300, 70, 5), 9), ("float32", (520, 33, 3), 4)])
def test_time_loop_single_gpu(dtype, domain, steps):
    """TimeLoop without neighbours: inp/out swapped every sweep, halos fixed -- the oracle applied
    `steps` times."""
    from oracle import native
    from stencil_benchmarks_b200 import distributed

    bench = horizontal_diffusion.Fused(domain=domain, dtype=dtype, verify=False, seed=3)
    data = bench.data()
    data.coeff[...] *= 0.025  # a stable time step (with U[0,1) rounding differences grow ~4x per sweep)
    mirrors = bench._device_fields(data)
    bench.upload(data, mirrors)
    loop = distributed.TimeLoop(bench, mirrors)
    for _ in range(steps):
        loop.step()
    capi.synchronize()
    state = loop.download(bench.empty_field())
    x, y = bench.empty_field(), bench.empty_field()
    x[...] = data.inp
    y[...] = data.inp
    for _ in range(steps):
        native.hdiff(x, data.coeff, y, tuple(bench.halo))
        x, y = y, x
    inner = bench.inner_slice()
    tol = dict(rtol=1e-12, atol=1e-14) if dtype == "float64" else dict(rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(state[inner], x[inner], **tol)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_partitioned_class_on_one_gpu(dtype):
    """The multi-GPU benchmark class degenerates to one slab: scatter, sweep, gather."""
    bench = horizontal_diffusion.Partitioned(domain=(300, 41, 6), dtype=dtype, gpus=1, verify=False, seed=5)
    data = bench.data()
    before = [np.array(f, copy=True) for f in data]
    result = bench.run()
    assert result["gpus"] == 1 and result["time"] > 0 and result["bandwidth"] > 0
    expected = stencils.hdiff(before[0], before[1], halo=bench.halo)
    inner = bench.inner_slice()
    assert close(data.out[inner], expected[inner], dtype)
    assert np.array_equal(data.inp, before[0]) and np.array_equal(data.coeff, before[1])


@pytest.mark.gpu
def test_vadv_float32_mismatch_fraction_is_of_the_reference_backends_order():
    """256x256x160 float32 on U[0,1) inputs: the fraction of points outside the reference tolerance
    stays within VADV_F32_MISMATCH; the measured fraction is recorded for DESIGN.md."""
    import json

    bench = vertical_advection.Thomas(domain=(256, 256, 160), dtype="float32", verify=False, seed=17)
    before = snapshot(bench)
    bench.run()
    inner = bench.inner_slice()
    expected = stencils._vadv_component(before["ustage"], before["upos"], before["utens"], before["utensstage"],
                                        before["wcon"], tuple(bench.halo), 1, 0)[inner]
    bad = ~np.isclose(bench.data().utensstage[inner], expected, **stencils.tolerances("float32"))
    fraction = float(bad.mean())
    out = pathlib.Path(__file__).parent.parent / "gpurun_out"
    if out.is_dir():
        (out / "vadv_f32_mismatch.json").write_text(json.dumps(
            dict(domain=[256, 256, 160], points=int(bad.size), mismatching=int(bad.sum()), fraction=fraction,
                 bound=VADV_F32_MISMATCH, reference_openmp_fraction="0.026-0.03 % (SURVEY.md 8c)")))
    assert fraction <= VADV_F32_MISMATCH, f"{bad.sum()} of {bad.size} points differ"
