"""STREAM on the B200 (``csrc/stream.cu``).

Counterpart of ``Native`` in stencil_benchmarks/benchmarks_collection/stream/cuda_hip.py:41-144:
same result format -- a list of four dicts (copy, scale, add, triad) with
``bandwidth`` in MB/s from the minimum time, ``avg-time``, ``time`` (min),
``max-time`` -- parsed from the McCalpin table the C side prints to stdout.
The reference's tuning parameters keep their names (cuda_hip.py:44-66) so that scripts written
for ``stream cuda-hip native`` run unchanged: ``block_size``, ``vector_size``, ``unroll_factor``
and ``streaming_loads/stores`` select the launch shape of the sm_100a kernels, with 0 / the
defaults meaning the measured best; parameters that only steer the reference's
code generator are accepted and have no effect.
"""

import re
import warnings

from ... import capi
from ...benchmark import Benchmark, ExecutionError, Parameter, ParameterError
from ...tools import cabi

_LINE = re.compile(r"(Copy|Scale|Add|Triad): +([0-9.]+) +([0-9.]+) +([0-9.]+) +([0-9.]+)")


class Native(Benchmark):
    array_size = Parameter("number of elements in arrays", 10000000)
    ntimes = Parameter("number of runs", 10)
    dtype = Parameter("data type in NumPy format, e.g. float32 or float64", "float64")
    verify = Parameter("verify results", True)
    compiler = Parameter(
        "compiler path: if given, the kernels are compiled at setup() through the reference's "
        "tools.compilation path (stream/cuda_hip.py:88-94); empty: the prebuilt library", "")
    compiler_flags = Parameter("additional compiler flags (with `compiler`)", "")
    device = Parameter("CUDA device ordinal", 0)
    block_size = Parameter("threads per block (0: the measured best, 512)", 0)
    vector_size = Parameter("elements per vector access (0: the measured best, 16 bytes; "
                            "16 or 32 bytes are possible)", 0)
    unroll_factor = Parameter("vectors per thread, 1/2/4/8 (0: the measured best, 4)", 0)
    streaming_stores = Parameter("use streaming (.cs) stores", True)
    streaming_loads = Parameter("use streaming (.cs) loads", True)
    # accepted for compatibility with scripts written for `stream cuda-hip native`; they steer the
    # reference's code generator and have no counterpart in the hand-written kernels
    axis = Parameter("compute grid dimension to use (no effect)", "x", choices=["x", "y", "z"])
    explicit_vectorization = Parameter("vector types instead of compiler vectorisation (no effect)", True)
    launch_bounds = Parameter("specify launch bounds (no effect)", True)
    index_type = Parameter("index data type (no effect)", "std::size_t")
    store_cache_modifier = Parameter("PTX cache modifier for stores (only the default)", "", choices=[""])
    load_cache_modifier = Parameter("PTX cache modifier for loads (only the default)", "", choices=[""])
    print_code = Parameter("print the CUDA source of the kernels", False)

    #: arrays are padded to whole 128-byte lines (the reference pads to
    #: block*vector*unroll elements, cuda_hip.py:81-86)
    granularity_bytes = 128

    def setup(self):
        super().setup()
        try:
            self._dtype_code = capi.dtype_code(self.dtype)
        except (ValueError, TypeError) as error:
            raise ParameterError(str(error)) from error
        if self.array_size <= 0:
            raise ParameterError("array size must be positive")
        itemsize = 4 if self._dtype_code == capi.F32 else 8
        if self.vector_size * itemsize not in (0, 16, 32):
            raise ParameterError("vector_size: the kernels move 16- or 32-byte vectors "
                                 f"({16 // itemsize} or {32 // itemsize} elements of {self.dtype})")
        if self.unroll_factor not in (0, 1, 2, 4, 8):
            raise ParameterError("unroll_factor must be 1, 2, 4 or 8")
        if self.block_size and (self.block_size % 32 or not 32 <= self.block_size <= 1024):
            raise ParameterError("block_size must be a multiple of 32 in [32, 1024]")
        if self.streaming_loads != self.streaming_stores:
            raise ParameterError("streaming loads and stores are switched together")
        if self.ntimes < 2:
            raise ParameterError("ntimes must be at least 2 (the first run is discarded)")
        elements = self.granularity_bytes // (4 if self._dtype_code == capi.F32 else 8)
        if self.array_size % elements:
            warnings.warn("adapting array size to match block and vector sizes")
        self.array_size = -(-self.array_size // elements) * elements
        try:
            self._lib = capi.library()
            self._kernels = (capi.jit_library(self.compiler, self.compiler_flags, "stream.cu")
                             if self.compiler else self._lib)
        except cabi.CompilationError as error:
            raise ParameterError(*error.args) from error
        if self.print_code:
            print((capi.ROOT / "csrc" / "stream.cu").read_text())

    def run(self):
        try:
            capi.require_device()
            self._lib.sb200_set_device(self.device)
            itemsize = 4 if self._dtype_code == capi.F32 else 8
            self._kernels.sb200_stream_configure(self.block_size, self.unroll_factor,
                                                 self.vector_size * itemsize, int(self.streaming_stores))
            output = self._kernels.sb200_stream_run(
                self._dtype_code, self.array_size, self.ntimes, int(self.verify)
            )
        except cabi.ExecutionError as error:
            raise ExecutionError(*error.args) from error
        return [
            {
                "name": match.group(1).lower(),
                "bandwidth": float(match.group(2)),
                "avg-time": float(match.group(3)),
                "time": float(match.group(4)),
                "max-time": float(match.group(5)),
            }
            for match in _LINE.finditer(output)
        ]
