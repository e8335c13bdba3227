"""The oracle (NumPy + C restatements) must reproduce the reference's own expected values.

tests/golden/*.npz were captured from the reference's verify_stencil
(tests/golden/make_golden.py); float64 must match bit for bit, float32 too
(same operation order, same precision).
"""

import glob
import pathlib

import numpy as np
import pytest

from oracle import native, stencils

GOLDEN = pathlib.Path(__file__).parent / "golden"
CASES = sorted(pathlib.Path(p).stem for p in glob.glob(str(GOLDEN / "*.npz")))


def load(case):
    with np.load(GOLDEN / f"{case}.npz") as data:
        return {key: data[key] for key in data.files}


def fortran_like(array):
    """Copy with i as the unit-stride axis (layout (2,1,0)) for the C oracle."""
    return np.asfortranarray(array)


def expected_numpy(case, g):
    halo = tuple(int(h) for h in g["halo"])
    inner = stencils.interior(g[next(k for k in g if k.startswith("in_"))].shape, halo)
    kind = case.split("_")[0]
    if kind == "copy":
        return {"out": stencils.copy(g["in_inp"], halo)[inner]}
    if kind == "onesided":
        return {"out": stencils.onesided_average(g["in_inp"], halo, int(case.split("_ax")[1][0]))[inner]}
    if kind == "symmetric":
        return {"out": stencils.symmetric_average(g["in_inp"], halo, int(case.split("_ax")[1][0]))[inner]}
    if kind == "laplacian":
        mask = int(case.split("_m")[1][0])
        return {"out": stencils.laplacian(g["in_inp"], halo, (mask & 1, mask & 2, mask & 4))[inner]}
    if kind == "hdiff":
        return {"out": stencils.hdiff(g["in_inp"], g["in_coeff"], halo)[inner]}
    if kind == "vadv" and "_all_" not in case:
        out = stencils.vadv(g["in_ustage"], g["in_upos"], g["in_utens"], g["in_utensstage"],
                            g["in_wcon"], halo)
        return {"utensstage": out[inner]}
    if kind == "vadv":
        u, v, w = stencils.vadv_all(
            *[tuple(g[f"in_{c}{f}"] for f in ("stage", "pos", "tens", "tensstage")) for c in "uvw"],
            g["in_wcon"], halo)
        return {"utensstage": u[inner], "vtensstage": v[inner], "wtensstage": w[inner]}
    raise AssertionError(case)


def expected_c(case, g):
    halo = tuple(int(h) for h in g["halo"])
    f = {k[3:]: fortran_like(v) for k, v in g.items() if k.startswith("in_")}
    first = next(iter(f.values()))
    inner = stencils.interior(first.shape, halo)
    kind = case.split("_")[0]
    if kind in ("copy", "onesided", "symmetric", "laplacian"):
        out = fortran_like(np.zeros_like(first))
        if kind == "copy":
            native.copy(f["inp"], out, halo)
        elif kind == "laplacian":
            mask = int(case.split("_m")[1][0])
            native.laplacian(f["inp"], out, halo, (mask & 1, mask & 2, mask & 4))
        else:
            native.average(f["inp"], out, halo, int(case.split("_ax")[1][0]), kind == "symmetric")
        return {"out": out[inner]}
    if kind == "hdiff":
        out = fortran_like(np.zeros_like(first))
        native.hdiff(f["inp"], f["coeff"], out, halo)
        return {"out": out[inner]}
    components = [("u", 1, 0)] + ([("v", 0, 1), ("w", 0, 0)] if "_all_" in case else [])
    result = {}
    for c, ishift, jshift in components:
        native.vadv(f[c + "stage"], f[c + "pos"], f[c + "tens"], f[c + "tensstage"], f["wcon"],
                    f["ccol"], f["dcol"], halo, ishift, jshift)
        result[c + "tensstage"] = f[c + "tensstage"][inner]
    return result


def test_golden_files_present():
    assert len(CASES) >= 40


@pytest.mark.parametrize("case", CASES)
def test_numpy_oracle_matches_reference(case):
    g = load(case)
    for name, values in expected_numpy(case, g).items():
        np.testing.assert_array_equal(values, g["expected_" + name], err_msg=f"{case}:{name}")


@pytest.mark.parametrize("case", CASES)
def test_c_oracle_matches_reference(case):
    g = load(case)
    for name, values in expected_c(case, g).items():
        np.testing.assert_array_equal(values, g["expected_" + name], err_msg=f"{case}:{name}")


@pytest.mark.parametrize("dtype,ntimes", [("float64", 10), ("float32", 10), ("float64", 2)])
def test_stream_closed_form(dtype, ntimes):
    """The recurrence of cuda_hip.j2:329-345: after one round a=15, b=3, c=4; x15 per round."""
    a, b, c = stencils.stream_expected(ntimes, dtype)
    if dtype == "float64":  # powers of 15 stay exact in float64 for these ntimes
        assert (float(a), float(b), float(c)) == (15.0**ntimes, 3 * 15.0 ** (ntimes - 1),
                                                  4 * 15.0 ** (ntimes - 1))
    x = np.full(64, 1, dtype), np.full(64, 2, dtype), np.zeros(64, dtype)
    for _ in range(ntimes):
        x = stencils.stream_ops(*x)
    assert (x[0][0], x[1][0], x[2][0]) == (a, b, c)
    y = [np.full(64, 1, dtype), np.full(64, 2, dtype), np.zeros(64, dtype)]
    for _ in range(ntimes):
        native.stream_round(*y)
    assert (y[0][-1], y[1][-1], y[2][-1]) == (a, b, c)
