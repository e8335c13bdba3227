"""ctypes access to oracle/liboracle.so (the C restatement in oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of oracle.c.  Arrays are NumPy fields
including the halo; the wrappers pass pointers to the first interior element and
element strides, exactly like the product's C ABI.
"""

import ctypes
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).parent.resolve()
LIBRARY = HERE / "liboracle.so"


def build(force=False):
    source = HERE / "oracle.c"
    if force or not LIBRARY.exists() or LIBRARY.stat().st_mtime < source.stat().st_mtime:
        subprocess.run(
            ["gcc", "-O2", "-fopenmp", "-fno-fast-math", "-ffp-contract=off", "-shared", "-fPIC",
             "-o", str(LIBRARY), str(source), "-lm"],
            check=True,
        )
    return LIBRARY


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
    return _lib


def _suffix(array):
    return {np.dtype("float64"): "f64", np.dtype("float32"): "f32"}[array.dtype]


def _ptr(array, halo):
    offset = sum(int(s) * int(h) for s, h in zip(array.strides, halo))
    return ctypes.c_void_p(array.ctypes.data + offset)


def _geometry(array, halo):
    nx, ny, nz = (n - 2 * h for n, h in zip(array.shape, halo))
    sx, sy, sz = (s // array.itemsize for s in array.strides)
    assert sx == 1, "the C oracle expects i to be the unit-stride axis"
    return [ctypes.c_int64(v) for v in (nx, ny, nz, sy, sz)]


def copy(inp, out, halo):
    getattr(lib(), "oracle_copy_" + _suffix(inp))(_ptr(inp, halo), _ptr(out, halo), *_geometry(inp, halo))


def average(inp, out, halo, axis, symmetric):
    getattr(lib(), "oracle_average_" + _suffix(inp))(
        _ptr(inp, halo), _ptr(out, halo), *_geometry(inp, halo), ctypes.c_int(axis),
        ctypes.c_int(int(symmetric)))


def laplacian(inp, out, halo, along):
    mask = int(bool(along[0])) | int(bool(along[1])) << 1 | int(bool(along[2])) << 2
    getattr(lib(), "oracle_laplacian_" + _suffix(inp))(
        _ptr(inp, halo), _ptr(out, halo), *_geometry(inp, halo), ctypes.c_int(mask))


def hdiff(inp, coeff, out, halo):
    getattr(lib(), "oracle_hdiff_" + _suffix(inp))(
        _ptr(inp, halo), _ptr(coeff, halo), _ptr(out, halo), *_geometry(inp, halo))


def vadv(stage, pos, tens, tensstage, wcon, ccol, dcol, halo, ishift=1, jshift=0):
    """In place on ``tensstage`` (like the kernels); ccol/dcol are scratch."""
    getattr(lib(), "oracle_vadv_" + _suffix(stage))(
        _ptr(stage, halo), _ptr(pos, halo), _ptr(tens, halo), _ptr(tensstage, halo),
        _ptr(wcon, halo), _ptr(ccol, halo), _ptr(dcol, halo), *_geometry(stage, halo),
        ctypes.c_int(ishift), ctypes.c_int(jshift))


def stream_round(a, b, c, scalar=3.0):
    ctype = ctypes.c_double if a.dtype == np.float64 else ctypes.c_float
    getattr(lib(), "oracle_stream_round_" + _suffix(a))(
        ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(b.ctypes.data),
        ctypes.c_void_p(c.ctypes.data), ctypes.c_uint64(a.size), ctype(scalar))


def stream_triad(a, b, c, scalar=3.0):
    ctype = ctypes.c_double if a.dtype == np.float64 else ctypes.c_float
    getattr(lib(), "oracle_stream_triad_" + _suffix(a))(
        ctypes.c_void_p(a.ctypes.data), ctypes.c_void_p(b.ctypes.data),
        ctypes.c_void_p(c.ctypes.data), ctypes.c_uint64(a.size), ctype(scalar))
