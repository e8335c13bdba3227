"""Loading and calling the C-ABI library, with the reference's call convention.

Mirrors ``stencil_benchmarks/tools/compilation.py`` of the reference (file kept under another name):

* ``Library`` wraps a shared library the way ``GnuLibrary`` (compilation.py:89-196)
  wraps its JIT-compiled one: every attribute is a callable around the C
  function, the C function returns ``int`` (0 = success), a non-zero return
  raises ``ExecutionError(captured stderr)``, stderr output on success becomes a
  Python warning and the captured stdout is the return value (that is how
  STREAM hands back its table, stream/cuda_hip.py:121-144).
* stdout/stderr are captured at file-descriptor level (compilation.py:54-86) so
  output written by C code is seen (``FdCapture``).
* ``GnuLibrary(code, compile_command, extension)`` compiles a source string and loads it as the
  reference's class of that name does (compilation.py:89-153); ``dtype_as_ctype`` / ``ctype_cname`` /
  ``dtype_cname`` / ``data_ptr`` (compilation.py:199-282) keep their meaning.  The reference's own unit
  tests of these (test/tools/test_compilation.py) run against this module in
  tests/test_sbench_dropin.py.

Unlike ``GnuLibrary`` the library is pre-built in-tree by ``__graft_entry__.build()``
(``nvcc -gencode arch=compute_100a,code=sm_100a``); ``compile_library`` offers the
reference-style "compile this source now" path for out-of-tree use.
"""

import contextlib
import ctypes
import os
import pathlib
import subprocess
import sys
import tempfile
import warnings
from typing import List, Optional, Tuple, Union

import numpy as np


class CompilationError(RuntimeError):
    pass


class ExecutionError(RuntimeError):
    pass


class FdCapture:
    """What C code writes to file descriptors 1 and 2 while the block runs.

    The descriptors are pointed at two anonymous files for the duration of the block and restored
    on every way out of it (Python-level redirection would not see ``printf`` / ``fprintf`` of a
    shared library); afterwards ``out`` and ``err`` hold the text.  Python's own buffered streams
    are flushed first so that nothing written earlier lands in the capture.
    """

    def __init__(self):
        self.out = self.err = ""
        self._slots = []

    def __enter__(self):
        for stream in (sys.stdout, sys.stderr):
            with contextlib.suppress(Exception):
                stream.flush()
        try:
            for fileno in (1, 2):
                scratch = tempfile.TemporaryFile()
                self._slots.append((fileno, os.dup(fileno), scratch))
                os.dup2(scratch.fileno(), fileno)
        except BaseException:
            self._restore()
            raise
        return self

    def _restore(self):
        texts = {}
        while self._slots:
            fileno, original, scratch = self._slots.pop()
            os.dup2(original, fileno)
            os.close(original)
            scratch.seek(0)
            texts[fileno] = scratch.read().decode(errors="replace")
            scratch.close()
        return texts

    def __exit__(self, *exc):
        texts = self._restore()
        self.out, self.err = texts.get(1, ""), texts.get(2, "")
        return False


class CFunction:
    """One ``extern "C" int f(...)`` of a library under the reference's call convention
    (compilation.py:155-196): the return value is a status, 0 = success; a failure raises
    ``ExecutionError`` carrying what the function wrote to stderr; stderr text of a successful
    call is a Python warning; the call evaluates to what the function wrote to stdout.  The
    optional keyword ``argtypes`` sets the ctypes prototype, as the reference's callers do."""

    def __init__(self, name: str, function):
        self.__name__ = name
        self._function = function

    def __call__(self, *args, argtypes=None) -> str:
        if argtypes is not None:
            self._function.argtypes = argtypes
        with FdCapture() as captured:
            status = self._function(*args)
        if status != 0:
            raise ExecutionError(captured.err)
        if captured.err:
            warnings.warn(f"unexpected output in call to {self.__name__}(…) to stderr:\n" + captured.err)
        return captured.out


class Library:
    """ctypes library whose functions follow the int-return / stderr convention."""

    def __init__(self, path: Union[str, os.PathLike]):
        self.path = pathlib.Path(path)
        if not self.path.exists():
            raise CompilationError(
                f"{self.path} does not exist: build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a)"
            )
        self._library = ctypes.CDLL(str(self.path))

    @property
    def raw(self) -> ctypes.CDLL:
        """The bare ctypes handle (no output capture, no exception mapping)."""
        return self._library

    def __getattr__(self, attr: str) -> CFunction:
        # only reached for names that are not instance attributes: every other name is a symbol
        if attr.startswith("_"):
            raise AttributeError(attr)
        return CFunction(attr, getattr(self._library, attr))


def compile_library(
    sources: List[Union[str, os.PathLike]],
    output: Union[str, os.PathLike],
    compile_command: Optional[List[str]] = None,
) -> Library:
    """Compile CUDA sources into a shared library with nvcc for sm_100a and load it.

    The flag handling follows ``GnuLibrary`` (compilation.py:125-128): for an nvcc
    command ``-Xcompiler -shared -Xcompiler -fPIC`` is appended.
    """
    if compile_command is None:
        compile_command = [
            "nvcc",
            "-std=c++17",
            "-O3",
            "-gencode",
            "arch=compute_100a,code=sm_100a",
            "-lineinfo",
        ]
    run_compiler(sources, output, list(compile_command))
    return Library(output)


def run_compiler(sources, output, command: List[str]) -> None:
    """``[compiler, -o, output, sources...] + flags`` plus the shared-library flags of the compiler
    family; raises ``CompilationError(stderr)``, warns about any other compiler output."""
    shared = ["-Xcompiler", "-shared", "-Xcompiler", "-fPIC"] if command[0].endswith("nvcc") else ["-shared", "-fPIC"]
    result = subprocess.run([command[0], "-o", str(output)] + [str(s) for s in sources] + command[1:] + shared,
                            capture_output=True)
    if result.returncode != 0:
        raise CompilationError(result.stderr.decode())
    if result.stdout or result.stderr:
        warnings.warn("unexpected compilation output: " + result.stdout.decode() + result.stderr.decode())


class GnuLibrary(Library):
    """``GnuLibrary(code, compile_command, extension)`` of the reference (compilation.py:89-153) for a
    machine without the reference package: the source string is written to
    ``./benchmarks_source_code/`` (kept, as there), compiled into a scratch ``.so`` with
    ``[compiler, -o, lib, src] + flags + -shared -fPIC`` (``-Xcompiler`` forms for nvcc) and
    loaded; defaults: extension ``.cpp``, compiler ``gcc`` for ``.c`` and ``g++`` otherwise.
    Compiler failure -> ``CompilationError(stderr)``, any compiler output -> a warning."""

    def __init__(self, code: str, compile_command: Optional[List[str]] = None,
                 extension: Optional[str] = None):
        extension = extension or ".cpp"
        if compile_command is None:
            compile_command = ["gcc" if extension.lower() == ".c" else "g++"]
        kept = pathlib.Path("benchmarks_source_code")
        kept.mkdir(exist_ok=True)
        handle, source = tempfile.mkstemp(suffix=extension, dir=kept)
        with os.fdopen(handle, "w") as stream:
            stream.write(code)
        with tempfile.TemporaryDirectory(prefix="sb200_lib_") as scratch:
            target = pathlib.Path(scratch) / "library.so"
            run_compiler([source], target, list(compile_command))
            super().__init__(target)  # stays mapped after the scratch directory is gone


_CTYPES = {"f4": ctypes.c_float, "f8": ctypes.c_double}
_CTYPES.update({f"{kind}{size}": getattr(ctypes, f"c_{'u' if kind == 'u' else ''}int{8 * size}")
                for kind in "iu" for size in (1, 2, 4, 8)})


def dtype_as_ctype(dtype):
    """ctypes scalar type of a NumPy dtype (compilation.py:199-225)."""
    dt = np.dtype(dtype)
    try:
        return _CTYPES[f"{dt.kind}{dt.itemsize}"]
    except KeyError:
        raise NotImplementedError(f"Conversion of type {dt} is not supported") from None


def ctype_cname(ctype) -> str:
    """C spelling of a ctypes scalar type (compilation.py:228-251)."""
    for key, candidate in _CTYPES.items():
        if candidate is ctype:
            return dtype_cname(np.dtype(key))
    raise NotImplementedError(f"Conversion of type {ctype} is not supported")


_C_NAMES = {"f2": "half", "f4": "float", "f8": "double"}


def dtype_cname(dtype) -> str:
    """C spelling of a NumPy dtype (same mapping as the reference's dtype_cname)."""
    dt = np.dtype(dtype)
    key = f"{dt.kind}{dt.itemsize}"
    if key in _C_NAMES:
        return _C_NAMES[key]
    if dt.kind in "iu":
        return f"std::{'u' if dt.kind == 'u' else ''}int{8 * dt.itemsize}_t"
    raise NotImplementedError(f"Conversion of type {dt} is not supported")


def data_ptr(array: np.ndarray, offset: Union[int, Tuple[int, ...], None] = None) -> ctypes.c_void_p:
    """Address of ``array[offset]`` as a ``void*``.

    ``offset`` is a flat element count or an index tuple (the stencils pass the
    halo, so the C side sees the first interior element -- the reference's
    convention, compilation.py:273-282).
    """
    address = array.ctypes.data
    if isinstance(offset, int):
        address += offset * array.itemsize
    elif offset is not None:
        address += sum(int(s) * int(o) for s, o in zip(array.strides, offset))
    return ctypes.c_void_p(address)
