"""CPU oracle: NumPy restatement of the reference's algorithms for the GPU hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``stencil_benchmarks_b200/`` may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the
checker or the reported CPU baseline.

Every function restates one ``verify_stencil`` of the reference
(stencil_benchmarks/benchmarks_collection/stencils/base.py, cited per function
as ``base.py:<lines>``) and returns the expected output field(s) instead of
comparing.  Operation order is kept identical to the reference so float64
results are bit-identical to what the reference computes.

Parity is PINNED: ``tests/golden/*.npz`` hold inputs and the ``expected`` arrays
captured from the reference's own ``verify_stencil`` (by intercepting its
``validation.check_equality`` calls, see ``tests/golden/make_golden.py``), and
``tests/test_oracle.py`` requires this module to reproduce them bit for bit.

Conventions: fields are full (i, j, k) arrays including the halo; ``halo`` is
the (hi, hj, hk) triple; results are returned as full-size arrays whose interior
holds the expected values (the halo of an output is unspecified in the reference,
here it is a copy of the corresponding input/initial field or zero).
"""

import numpy as np

# tolerances of the reference's validation (stencil_benchmarks/tools/validation.py:98-106)
TOLERANCES = {
    np.dtype("float64"): dict(rtol=1e-5, atol=1e-8),
    np.dtype("float32"): dict(rtol=1e-4, atol=1e-5),
}


def tolerances(dtype):
    return TOLERANCES[np.dtype(dtype)]


def interior(shape, halo, shift=(0, 0, 0)):
    """Slices of the interior, optionally shifted (base.py:127-137 ``inner_slice``)."""
    return tuple(
        slice(h + s, n - h + s) for n, h, s in zip(shape, halo, shift)
    )


def _unit(axis, sign=1):
    shift = [0, 0, 0]
    shift[axis] = sign
    return shift


def copy(inp, halo):
    """base.py:181-188: out[interior] = inp[interior]."""
    out = np.zeros_like(inp)
    inner = interior(inp.shape, halo)
    out[inner] = inp[inner]
    return out


def onesided_average(inp, halo, axis):
    """base.py:194-205: (inp[+1 along axis] + inp) / 2."""
    out = np.zeros_like(inp)
    inner = interior(inp.shape, halo)
    out[inner] = (inp[interior(inp.shape, halo, _unit(axis))] + inp[inner]) / 2
    return out


def symmetric_average(inp, halo, axis):
    """base.py:211-222: (inp[+1 along axis] + inp[-1 along axis]) / 2."""
    out = np.zeros_like(inp)
    inner = interior(inp.shape, halo)
    out[inner] = (
        inp[interior(inp.shape, halo, _unit(axis))] + inp[interior(inp.shape, halo, _unit(axis, -1))]
    ) / 2
    return out


def laplacian(inp, halo, along=(True, True, False)):
    """base.py:238-254: sum over the selected axes of 2*inp - inp[+1] - inp[-1]."""
    out = np.zeros_like(inp)
    inner = interior(inp.shape, halo)
    acc = np.zeros(inp[inner].shape, dtype=inp.dtype)
    for axis, active in enumerate(along):
        if active:
            acc += (
                2 * inp[inner]
                - inp[interior(inp.shape, halo, _unit(axis))]
                - inp[interior(inp.shape, halo, _unit(axis, -1))]
            )
    out[inner] = acc
    return out


def hdiff(inp, coeff, halo=None):
    """base.py:284-307: Laplacian, limited fluxes in i and j, update.

    The reference evaluates every stage on the whole padded array (one layer is
    lost per stage), so the result is defined on ``[2:-2, 2:-2, :]``; the caller
    compares the interior only.  ``halo`` is accepted for symmetry and unused.
    """
    lap = np.zeros_like(inp)
    lap[1:-1, 1:-1, :] = 4 * inp[1:-1, 1:-1, :] - (
        inp[2:, 1:-1, :] + inp[:-2, 1:-1, :] + inp[1:-1, 2:, :] + inp[1:-1, :-2, :]
    )

    flx = np.zeros_like(inp)
    flx[:-1] = lap[1:] - lap[:-1]
    flx[:-1] = np.where(flx[:-1] * (inp[1:] - inp[:-1]) > 0, 0, flx[:-1])

    fly = np.zeros_like(inp)
    fly[:, :-1] = lap[:, 1:] - lap[:, :-1]
    fly[:, :-1] = np.where(fly[:, :-1] * (inp[:, 1:] - inp[:, :-1]) > 0, 0, fly[:, :-1])

    out = np.zeros_like(inp)
    out[1:-1, 1:-1, :] = inp[1:-1, 1:-1, :] - coeff[1:-1, 1:-1, :] * (
        flx[1:-1, 1:-1, :] - flx[:-2, 1:-1, :] + fly[1:-1, 1:-1, :] - fly[1:-1, :-2, :]
    )
    return out


def _vadv_component(stage, pos, tens, tensstage, wcon, halo, ishift, jshift):
    """One forward + backward Thomas sweep (base.py:415-473) on interior columns.

    Returns the new ``tensstage`` (full array; halo copied from the input).
    """
    hi, hj, hk = halo
    nx, ny, nz = (n - 2 * h for n, h in zip(stage.shape, halo))
    dtype = stage.dtype.type

    def level(field, k, di=0, dj=0):
        return field[hi + di : hi + nx + di, hj + dj : hj + ny + dj, hk + k]

    dtr_stage = 3 / 20
    beta_v = 0
    bet_m = 0.5 * (1 - beta_v)
    bet_p = 0.5 * (1 + beta_v)

    ccol = np.empty((nx, ny, nz), dtype=dtype)
    dcol = np.empty((nx, ny, nz), dtype=dtype)

    # forward sweep, first level (base.py:417-429)
    k = 0
    gcv = 0.25 * (level(wcon, k + 1, ishift, jshift) + level(wcon, k + 1))
    cs = gcv * bet_m
    ccol[:, :, k] = gcv * bet_p
    bcol = dtr_stage - ccol[:, :, k]
    correction = -cs * (level(stage, k + 1) - level(stage, k))
    dcol[:, :, k] = dtr_stage * level(pos, k) + level(tens, k) + level(tensstage, k) + correction
    ccol[:, :, k] /= bcol
    dcol[:, :, k] /= bcol

    # interior levels (base.py:431-449)
    for k in range(1, nz - 1):
        gav = -0.25 * (level(wcon, k, ishift, jshift) + level(wcon, k))
        gcv = 0.25 * (level(wcon, k + 1, ishift, jshift) + level(wcon, k + 1))
        as_ = gav * bet_m
        cs = gcv * bet_m
        acol = gav * bet_p
        ccol[:, :, k] = gcv * bet_p
        bcol = dtr_stage - acol - ccol[:, :, k]
        correction = -as_ * (level(stage, k - 1) - level(stage, k)) - cs * (
            level(stage, k + 1) - level(stage, k)
        )
        dcol[:, :, k] = (
            dtr_stage * level(pos, k) + level(tens, k) + level(tensstage, k) + correction
        )
        divided = 1.0 / (bcol - ccol[:, :, k - 1] * acol)
        ccol[:, :, k] *= divided
        dcol[:, :, k] = (dcol[:, :, k] - dcol[:, :, k - 1] * acol) * divided

    # last level (base.py:451-462)
    k = nz - 1
    gav = -0.25 * (level(wcon, k, ishift, jshift) + level(wcon, k))
    as_ = gav * bet_m
    acol = gav * bet_p
    bcol = dtr_stage - acol
    correction = -as_ * (level(stage, k - 1) - level(stage, k))
    dcol[:, :, k] = dtr_stage * level(pos, k) + level(tens, k) + level(tensstage, k) + correction
    dcol[:, :, k] = (dcol[:, :, k] - dcol[:, :, k - 1] * acol) / (bcol - ccol[:, :, k - 1] * acol)

    # backward sweep (base.py:464-473)
    out = np.array(tensstage, copy=True)
    datacol = dcol[:, :, nz - 1].copy()
    level(out, nz - 1)[...] = dtr_stage * (datacol - level(pos, nz - 1))
    for k in range(nz - 2, -1, -1):
        datacol = dcol[:, :, k] - ccol[:, :, k] * datacol
        level(out, k)[...] = dtr_stage * (datacol - level(pos, k))
    return out


def vadv(ustage, upos, utens, utensstage, wcon, halo):
    """base.py:475-476 with all_components=False: the u solve, wcon neighbour (i+1, j)."""
    return _vadv_component(ustage, upos, utens, utensstage, wcon, halo, 1, 0)


def vadv_all(u, v, w, wcon, halo):
    """base.py:475-483 with all_components=True.

    ``u``, ``v``, ``w`` are (stage, pos, tens, tensstage) tuples; returns the three
    new tensstage fields.  wcon neighbours: (i+1, j) for u, (i, j+1) for v, (i, j) for w.
    """
    return (
        _vadv_component(*u, wcon, halo, 1, 0),
        _vadv_component(*v, wcon, halo, 0, 1),
        _vadv_component(*w, wcon, halo, 0, 0),
    )


def stream_expected(ntimes, dtype="float64", scalar=3):
    """Closed-form STREAM values after ``ntimes`` rounds (stream/cuda_hip.j2:329-345).

    Returns (a, b, c) computed in ``dtype`` arithmetic like the reference's host replay.
    """
    t = np.dtype(dtype).type
    a, b, c, q = t(1), t(2), t(0), t(scalar)
    for _ in range(ntimes):
        c = a
        b = q * c
        c = a + b
        a = b + q * c
    return a, b, c


def stream_ops(a, b, c, scalar=3):
    """One round of copy / scale / add / triad on arrays (stream/cuda_hip.j2:132-173)."""
    q = a.dtype.type(scalar)
    c = a.copy()
    b = q * c
    c = a + b
    a = b + q * c
    return a, b, c
