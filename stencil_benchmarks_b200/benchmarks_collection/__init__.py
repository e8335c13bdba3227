"""Benchmark definitions of the B200 backend.

Importing this package imports every benchmark module, which registers the
classes (``benchmark.REGISTRY``) -- the same mechanism as the reference's
``stencil_benchmarks/benchmarks_collection/__init__.py``.  Module paths are
chosen so that ``stencil_benchmarks.cli`` (cli.py:47-50) names the commands

    sbench stencils b200 basic {empty,copy,onesided-average,symmetric-average,laplacian}
    sbench stencils b200 horizontal-diffusion fused
    sbench stencils b200 vertical-advection thomas
    sbench stream b200 native
"""

from . import stencils, stream  # noqa: F401
