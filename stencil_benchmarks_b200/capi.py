"""ctypes binding of ``libsbench_b200.so`` (``include/sbench_b200.h``).

This is the only place where Python meets the CUDA code.  It plays the role of
the two ctypes surfaces of the reference's GPU backend: the JIT library handle
returned by ``compilation.GnuLibrary`` (cuda_hip/mixin.py:79-84, :171) and the
``libcudart`` handle of ``cuda_hip/api.py:39-104``.

There is deliberately no fallback: if the library has not been built, loading
raises, and every benchmark that needs it fails with that error.
"""

import ctypes
import functools
import os
import pathlib
import re
import weakref

import numpy as np

from .tools import cabi

ROOT = pathlib.Path(__file__).parent.resolve()
LIBRARY_PATH = ROOT / "csrc" / "libsbench_b200.so"
HEADER_PATH = ROOT.parent / "include" / "sbench_b200.h"

F32, F64 = 0, 1
BASIC_EMPTY, BASIC_COPY, BASIC_ONESIDED_AVG, BASIC_SYMMETRIC_AVG, BASIC_LAPLACIAN = range(5)
STREAM_COPY, STREAM_SCALE, STREAM_ADD, STREAM_TRIAD, STREAM_INIT = range(5)
VADV_AUTO, VADV_GLOBAL, VADV_ONCHIP = range(3)

_vp = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_sz = ctypes.c_size_t
_dp = ctypes.POINTER(ctypes.c_double)
_geom = [_i64] * 6  # nx, ny, nz, sx, sy, sz

# name -> (restype, argtypes); must list every function declared in the header
PROTOTYPES = {
    "sb200_version": (_i, []),
    "sb200_device_count": (_i, [ctypes.POINTER(_i)]),
    "sb200_set_device": (_i, [_i]),
    "sb200_get_device": (_i, [ctypes.POINTER(_i)]),
    "sb200_device_info": (_i, [ctypes.c_char_p, _i, ctypes.POINTER(_i), ctypes.POINTER(_u64),
                               ctypes.POINTER(_u64)]),
    "sb200_malloc": (_i, [ctypes.POINTER(_vp), _sz]),
    "sb200_free": (_i, [_vp]),
    "sb200_host_alloc": (_i, [ctypes.POINTER(_vp), _sz]),
    "sb200_host_free": (_i, [_vp]),
    "sb200_host_register": (_i, [_vp, _sz]),
    "sb200_host_unregister": (_i, [_vp]),
    "sb200_memcpy_h2d": (_i, [_vp, _vp, _sz, _vp, _i]),
    "sb200_memcpy_d2h": (_i, [_vp, _vp, _sz, _vp, _i]),
    "sb200_memcpy_d2d": (_i, [_vp, _vp, _sz, _vp, _i]),
    "sb200_memset": (_i, [_vp, _i, _sz, _vp, _i]),
    "sb200_memcpy2d_h2d": (_i, [_vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "sb200_memcpy2d_d2h": (_i, [_vp, _sz, _vp, _sz, _sz, _sz, _vp]),
    "sb200_stream_create": (_i, [ctypes.POINTER(_vp)]),
    "sb200_stream_destroy": (_i, [_vp]),
    "sb200_event_create": (_i, [ctypes.POINTER(_vp)]),
    "sb200_event_destroy": (_i, [_vp]),
    "sb200_event_record": (_i, [_vp, _vp]),
    "sb200_stream_wait_event": (_i, [_vp, _vp]),
    "sb200_event_elapsed": (_i, [_vp, _vp, _dp]),
    "sb200_synchronize": (_i, [_vp]),
    "sb200_flush_l2": (_i, [_vp]),
    "sb200_launch_count": (_u64, []),
    "sb200_stream_configure": (_i, [_i, _i, _i, _i]),
    "sb200_stream_run": (_i, [_i, _u64, _i, _i]),
    "sb200_stream_op": (_i, [_i, _i, _vp, _vp, _vp, _u64, ctypes.c_double, _i, _dp, _vp]),
    "sb200_basic": (_i, [_i, _i, _vp, _vp] + _geom + [_i, _i, _i, _dp, _vp]),
    "sb200_hdiff": (_i, [_i, _vp, _vp, _vp] + _geom + [_i, _dp, _vp]),
    "sb200_hdiff_tiling": (_i, [_i, _i64, _i64, _i64, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i),
                                ctypes.POINTER(_i64)]),
    "sb200_hdiff_peer": (_i, [_i, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64] + _geom + [_i, _dp, _vp]),
    "sb200_hdiff_step": (_i, [_i, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, ctypes.c_uint32]
                         + _geom + [_dp, _vp]),
    "sb200_enable_peer_access": (_i, [_i, _i]),
    "sb200_ipc_get_handle": (_i, [_vp, _vp]),
    "sb200_ipc_open_handle": (_i, [_vp, ctypes.POINTER(_vp)]),
    "sb200_ipc_close_handle": (_i, [_vp]),
    "sb200_vadv": (_i, [_i] + [_vp] * 8 + _geom + [_i, _i, _i, _i, _dp, _vp]),
    "sb200_vadv_components": (_i, [_i, _i] + [ctypes.POINTER(_vp)] * 4 + [ctypes.POINTER(_i)] * 2 + [_vp] * 3
                              + _geom + [_i, _i, _dp, _vp]),
    "sb200_pack_rows": (_i, [_i, _vp, _vp] + [_i64] * 7 + [_vp]),
    "sb200_unpack_rows": (_i, [_i, _vp, _vp] + [_i64] * 7 + [_vp]),
}


def declared_symbols(header=HEADER_PATH):
    """Function names declared in the C header."""
    text = pathlib.Path(header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", text)))


@functools.lru_cache(maxsize=1)
def library() -> cabi.Library:
    """Load the library and attach prototypes; raises if it is missing or incomplete."""
    lib = cabi.Library(LIBRARY_PATH)
    missing = []
    for name, (restype, argtypes) in PROTOTYPES.items():
        try:
            func = getattr(lib.raw, name)
        except AttributeError:
            missing.append(name)
            continue
        func.restype = restype
        func.argtypes = argtypes
    if missing:
        raise cabi.CompilationError(
            f"{LIBRARY_PATH} does not export: {', '.join(missing)} (stale build?)"
        )
    return lib


class TypedLibrary:
    """A JIT-compiled library (the reference's ``GnuLibrary`` or the stand-alone mirror) whose
    calls carry the prototypes of include/sbench_b200.h: both wrappers take ``argtypes=`` per call
    (compilation.py:155-160), so int64 / double / pointer arguments are converted correctly."""

    def __init__(self, library):
        self.library = library

    def __getattr__(self, name):
        function = getattr(self.library, name)
        argtypes = PROTOTYPES[name][1]
        return lambda *args: function(*args, argtypes=argtypes)


_JIT_CACHE = {}


def jit_library(compiler, compiler_flags, source):
    """Compile one kernel family at ``setup()`` the way the reference's backends do
    (cuda_hip/mixin.py:60-84): source string -> ``tools.compilation.GnuLibrary`` -> ctypes handle.

    The source is csrc/runtime.cu + csrc/<source> as one translation unit; the command is
    ``<compiler> -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo <flags>``.
    With the reference package importable its own ``GnuLibrary`` compiles and loads (source kept
    under ./benchmarks_source_code, nvcc flags appended: compilation.py:125-153); otherwise the
    mirror in tools/cabi.py does the same.  One compilation per (compiler, flags, source) and process.
    """
    import shlex

    key = (compiler, compiler_flags, source)
    if key in _JIT_CACHE:
        return _JIT_CACHE[key]
    code = "".join(f'#include "{ROOT / "csrc" / name}"\n' for name in ("runtime.cu", source))
    command = [compiler, "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"]
    command += shlex.split(compiler_flags or "")
    try:
        from stencil_benchmarks.tools import compilation
    except ImportError:
        compilation = None
    if compilation is not None:
        try:
            library = compilation.GnuLibrary(code, command, extension=".cu")
        except compilation.CompilationError as error:
            raise cabi.CompilationError(*error.args) from error
        except FileNotFoundError as error:
            raise cabi.CompilationError(f"compiler not found: {compiler}") from error
    else:
        try:
            library = cabi.GnuLibrary(code, command, extension=".cu")
        except FileNotFoundError as error:
            raise cabi.CompilationError(f"compiler not found: {compiler}") from error
    _JIT_CACHE[key] = TypedLibrary(library)
    return _JIT_CACHE[key]


def dtype_code(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return F32
    if dtype == np.float64:
        return F64
    raise ValueError(f"unsupported dtype {dtype}: the B200 kernels exist for float32 and float64")


def driver_present() -> bool:
    """True if the NVIDIA kernel driver exposes its device nodes.

    Checked before any CUDA runtime call: on a machine without the driver the
    runtime can block for minutes instead of reporting "no device".
    """
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/dxg")


def device_count() -> int:
    if not driver_present():
        return 0
    count = _i(0)
    library().sb200_device_count(ctypes.byref(count))
    return count.value


def require_device():
    """Raise ExecutionError unless a CUDA device is usable (no CPU fallback)."""
    if device_count() < 1:
        raise cabi.ExecutionError(
            "no CUDA device available: the B200 backend has no CPU fallback"
        )


def device_info():
    name = ctypes.create_string_buffer(256)
    sms = _i()
    mem = _u64()
    l2 = _u64()
    library().sb200_device_info(name, 256, ctypes.byref(sms), ctypes.byref(mem), ctypes.byref(l2))
    return dict(name=name.value.decode(), sm_count=sms.value, global_mem_bytes=mem.value,
                l2_bytes=l2.value)


class DeviceBuffer:
    """Owning handle of a ``sb200_malloc`` allocation (freed when collected).

    Same ownership model as ``Runtime.malloc`` in the reference
    (cuda_hip/api.py:54-78): Python owns device memory, a finalizer frees it.
    """

    def __init__(self, nbytes: int):
        ptr = _vp()
        library().sb200_malloc(ctypes.byref(ptr), nbytes)
        self.ptr = ptr.value
        self.nbytes = nbytes
        self._finalizer = weakref.finalize(self, library().raw.sb200_free, _vp(self.ptr))

    def free(self):
        self._finalizer()


class PinnedBuffer:
    """Page-locked host allocation usable as a NumPy buffer (``sb200_host_alloc``)."""

    def __init__(self, nbytes: int):
        ptr = _vp()
        library().sb200_host_alloc(ctypes.byref(ptr), max(int(nbytes), 1))
        self.ptr = ptr.value
        self.nbytes = int(nbytes)
        self.buffer = (ctypes.c_byte * max(int(nbytes), 1)).from_address(self.ptr)
        self._finalizer = weakref.finalize(self.buffer, library().raw.sb200_host_free, _vp(self.ptr))


def pinned_alloc(nbytes: int):
    """``alloc`` callable for ``tools.fields.alloc_array`` returning pinned memory."""
    return PinnedBuffer(nbytes).buffer


def memcpy_h2d(dptr, host_ptr, nbytes, stream=None, sync=True):
    library().sb200_memcpy_h2d(_vp(dptr), _vp(host_ptr), nbytes, _vp(stream), int(sync))


def memcpy_d2h(host_ptr, dptr, nbytes, stream=None, sync=True):
    library().sb200_memcpy_d2h(_vp(host_ptr), _vp(dptr), nbytes, _vp(stream), int(sync))


def synchronize(stream=None):
    library().sb200_synchronize(_vp(stream))


def launch_count() -> int:
    return int(library().raw.sb200_launch_count())
