"""Vertical advection on the B200: per-column Thomas solve (``csrc/vadv.cu``).

Counterpart of the four reference variants in
stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/vertical_advection.py:60-73.
With ``all_components`` the u, v and w systems are solved in one sweep that shares
the wcon field (base.py:475-483; the reference's ``LocalMemMerged``).
"""

import ctypes

import numpy as np

from .... import capi
from ....benchmark import Parameter
from .. import base
from .mixin import StencilMixin, _vp


class VerticalAdvectionMixin(StencilMixin):
    block_size = Parameter("block size of the reference's templates (no effect)", (32, 1))
    unroll_factor = Parameter("unrolling of the vertical loop in the reference's templates (no effect)", -1)

    kernel_source = "vadv.cu"
    coefficients = Parameter(
        "where the eliminated Thomas coefficients live between the sweeps",
        "auto", choices=["auto", "global", "onchip"],
    )

    @property
    def field_roles(self):
        roles = {"wcon": "in", "ccol": "scratch", "dcol": "scratch", "datacol": "scratch"}
        for c in "uvw":
            roles.update({c + "stage": "in", c + "pos": "in", c + "tens": "in",
                          c + "tensstage": "inout"})
        return roles

    @property
    def algorithmic_bytes(self):
        """5 reads + 1 write per component, wcon shared: 6N / 16N elements."""
        fields_moved = 6 if not self.all_components else 16
        return int(fields_moved * np.prod(self.domain) * np.dtype(self.dtype).itemsize)

    @property
    def j_reach(self):
        return int(self.all_components)  # the v solve reads wcon(i, j+1)

    def launch(self, pointers, dry_runs, time_ptr, stream, domain=None, rows=None):
        variant = {"auto": capi.VADV_AUTO, "global": capi.VADV_GLOBAL,
                   "onchip": capi.VADV_ONCHIP}[self.coefficients]
        components = [("u", 1, 0)]
        if self.all_components:
            components += [("v", 0, 1), ("w", 0, 0)]
        n = len(components)

        def table(suffix):
            return (_vp * n)(*[pointers[c + suffix] for c, _, _ in components])

        # one sweep for all components: they share wcon, which is read from HBM once
        self._kernels.sb200_vadv_components(
            self._dtype_code, n, table("stage"), table("pos"), table("tens"), table("tensstage"),
            (ctypes.c_int * n)(*[i for _, i, _ in components]),
            (ctypes.c_int * n)(*[j for _, _, j in components]),
            pointers["wcon"], pointers["ccol"], pointers["dcol"], *self.geometry(domain), variant,
            dry_runs, time_ptr, _vp(stream),
        )


class Thomas(VerticalAdvectionMixin, base.VerticalAdvectionStencil):
    alignment = Parameter("data alignment in bytes", 128)
