"""Horizontal diffusion on the B200: one fused kernel (``csrc/hdiff.cu``).

Counterpart of the nine reference variants in
stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/horizontal_diffusion.py:50-128.
"""

import numpy as np

from ....benchmark import Parameter
from .. import base
from .mixin import StencilMixin, _vp


class HorizontalDiffusionMixin(StencilMixin):
    field_roles = {"inp": "in", "coeff": "in", "out": "out"}
    kernel_source = "hdiff.cu"
    j_reach = 2

    @property
    def algorithmic_bytes(self):
        """inp on (nx+4)(ny+4)nz + coeff + out on the interior."""
        nx, ny, nz = self.domain
        return int((2 * nx * ny * nz + (nx + 4) * (ny + 4) * nz) * np.dtype(self.dtype).itemsize)

    def launch(self, pointers, dry_runs, time_ptr, stream, domain=None):
        self._lib.sb200_hdiff(
            self._dtype_code, pointers["inp"], pointers["coeff"], pointers["out"],
            *self.geometry(domain), dry_runs, time_ptr, _vp(stream),
        )


class Fused(HorizontalDiffusionMixin, base.HorizontalDiffusionStencil):
    alignment = Parameter("data alignment in bytes", 128)
