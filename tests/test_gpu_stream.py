"""GPU parity of the STREAM kernels (csrc/stream.cu) through the plugin API and the C ABI."""

import ctypes

import numpy as np
import pytest

from oracle import stencils
from stencil_benchmarks_b200 import capi
from stencil_benchmarks_b200.benchmarks_collection.stream import b200 as stream

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_native_runs_and_verifies(dtype):
    """Like `sbench stream cuda-hip native`: four result rows, MB/s from the min time, and the
    library's own closed-form verification (cuda_hip.j2:290-345) must pass."""
    with pytest.warns(UserWarning, match="adapting array size"):
        bench = stream.Native(array_size=(1 << 26) + 1, ntimes=5, dtype=dtype)
    assert bench.array_size % 16 == 0 and bench.array_size >= (1 << 26) + 1
    results = bench.run()
    assert [r["name"] for r in results] == ["copy", "scale", "add", "triad"]
    for r in results:
        assert r["bandwidth"] > 0 and 0 < r["time"] <= r["avg-time"] <= r["max-time"]
        factor = 2 if r["name"] in ("copy", "scale") else 3
        expected = 1e-6 * factor * bench.array_size * np.dtype(dtype).itemsize / r["time"]
        # the table prints times with 1e-6 s resolution (cuda_hip.j2:270-272)
        assert abs(r["bandwidth"] - expected) / expected < 0.6e-6 / r["time"] + 1e-3


def device_array(values):
    buffer = capi.DeviceBuffer(values.nbytes + 64)
    capi.memcpy_h2d(buffer.ptr, values.ctypes.data, values.nbytes)
    return buffer


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("n", [1, 3, 4, 31, 1024, 4099, 1 << 20, (1 << 20) + 5])
def test_ops_match_oracle(dtype, n):
    """copy/scale/add/triad on ragged sizes (tail handling) against the oracle, bit for bit."""
    lib = capi.library()
    rng = np.random.default_rng(n)
    host = [rng.random(n).astype(dtype) for _ in range(3)]
    dev = [device_array(h) for h in host]
    a, b, c = (h.copy() for h in host)
    q = a.dtype.type(3)
    code = capi.dtype_code(dtype)
    t = ctypes.c_double()
    expected_ops = [
        (capi.STREAM_COPY, lambda: (a, b, a.copy())),
        (capi.STREAM_SCALE, lambda: (a, q * c, c)),
        (capi.STREAM_ADD, lambda: (a, b, a + b)),
        (capi.STREAM_TRIAD, lambda: (b + q * c, b, c)),
    ]
    for op, expected in expected_ops:
        lib.sb200_stream_op(op, code, dev[0].ptr, dev[1].ptr, dev[2].ptr, n, 3.0, 0,
                            ctypes.byref(t), None)
        assert t.value > 0
        a, b, c = expected()
        for values, buffer, name in zip((a, b, c), dev, "abc"):
            out = np.empty_like(values)
            capi.memcpy_d2h(out.ctypes.data, buffer.ptr, out.nbytes)
            if op != capi.STREAM_TRIAD:
                np.testing.assert_array_equal(out, values, err_msg=f"op {op} array {name}")
            else:  # triad: the GPU's FMA rounds once, NumPy's multiply-then-add twice (<= 1 ulp)
                np.testing.assert_allclose(out, values, rtol=2.3e-16 if dtype == "float64" else 1.2e-7)


def test_closed_form_after_rounds():
    lib = capi.library()
    n = 1 << 16
    code = capi.F64
    dev = [capi.DeviceBuffer(8 * n) for _ in range(3)]
    lib.sb200_stream_op(capi.STREAM_INIT, code, dev[0].ptr, dev[1].ptr, dev[2].ptr, n, 3.0, 0, None, None)
    for _ in range(4):
        for op in (capi.STREAM_COPY, capi.STREAM_SCALE, capi.STREAM_ADD, capi.STREAM_TRIAD):
            lib.sb200_stream_op(op, code, dev[0].ptr, dev[1].ptr, dev[2].ptr, n, 3.0, 0, None, None)
    capi.synchronize()
    expected = stencils.stream_expected(4)
    for value, buffer in zip(expected, dev):
        out = np.empty(n)
        capi.memcpy_d2h(out.ctypes.data, buffer.ptr, out.nbytes)
        assert np.all(out == value)


def test_launch_counter_counts():
    before = capi.launch_count()
    bench = stream.Native(array_size=1 << 16, ntimes=2)
    bench.run()
    assert capi.launch_count() - before >= 1 + 2 * 4


@pytest.mark.parametrize("shape", [dict(block_size=256, vector_size=4, unroll_factor=2),
                                   dict(block_size=1024, vector_size=2, unroll_factor=1),
                                   dict(block_size=128, vector_size=4, unroll_factor=8, streaming_loads=False,
                                        streaming_stores=False)])
def test_native_with_the_reference_tuning_parameters(shape):
    """block_size / vector_size / unroll_factor / streaming_* of `stream cuda-hip native`
    (cuda_hip.py:44-66) select the launch shape; every shape verifies."""
    bench = stream.Native(array_size=1 << 24, ntimes=3, dtype="float64", **shape)
    try:
        results = bench.run()  # raises ExecutionError if the closed-form verification fails
    finally:
        capi.library().sb200_stream_configure(0, 0, 0, -1)
    assert [r["name"] for r in results] == ["copy", "scale", "add", "triad"]
    assert all(r["bandwidth"] > 1e5 for r in results)
