"""STREAM on the B200 (``csrc/stream.cu``).

Counterpart of ``Native`` in stencil_benchmarks/benchmarks_collection/stream/cuda_hip.py:41-144:
same result format -- a list of four dicts (copy, scale, add, triad) with
``bandwidth`` in MB/s from the minimum time, ``avg-time``, ``time`` (min),
``max-time`` -- parsed from the McCalpin table the C side prints to stdout.
Vector width, unrolling and cache policy are fixed properties of the sm_100a
kernels; the tuning parameters of the reference are not offered.
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

    def run(self):
        try:
            capi.require_device()
            self._lib.sb200_set_device(self.device)
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
