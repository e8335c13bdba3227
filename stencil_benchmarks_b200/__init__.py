"""B200-native (sm_100a) backend for GridTools' stencil_benchmarks.

The package holds only the GPU hot path of the reference -- STREAM, the basic
stencils, horizontal diffusion and vertical advection -- behind the
reference's plugin API (``Benchmark`` / ``Parameter`` classes, registration by
module path, ctypes C-ABI loading).  All compute happens in hand-written CUDA
kernels in ``csrc/`` reached through ``libsbench_b200.so`` (``include/sbench_b200.h``).
There is no CPU fallback: without the library or without a GPU the benchmarks
raise.
"""

__version__ = "0.1.0"
