"""Vertical advection on the B200: per-column Thomas solve (``csrc/vadv.cu``).

Counterpart of the four reference variants in
stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/vertical_advection.py:60-73.
With ``all_components`` the u, v and w solves run one after the other with the
wcon neighbour of each (base.py:475-483).
"""

import ctypes

import numpy as np

from .... import capi
from ....benchmark import Parameter
from .. import base
from .mixin import StencilMixin, _vp


class VerticalAdvectionMixin(StencilMixin):
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

    def launch(self, pointers, dry_runs, time_ptr, stream, domain=None):
        variant = {"auto": capi.VADV_AUTO, "global": capi.VADV_GLOBAL,
                   "onchip": capi.VADV_ONCHIP}[self.coefficients]
        components = [("u", 1, 0)]
        if self.all_components:
            components += [("v", 0, 1), ("w", 0, 0)]
        total = 0.0
        for index, (c, ishift, jshift) in enumerate(components):
            elapsed = ctypes.c_double()
            self._lib.sb200_vadv(
                self._dtype_code, pointers[c + "stage"], pointers[c + "pos"], pointers[c + "tens"],
                pointers[c + "tensstage"], pointers["wcon"], pointers["ccol"], pointers["dcol"],
                pointers["datacol"], *self.geometry(domain), ishift, jshift, variant,
                dry_runs, ctypes.byref(elapsed) if time_ptr is not None else None, _vp(stream),
            )
            total += elapsed.value
        if time_ptr is not None:
            ctypes.cast(time_ptr, ctypes.POINTER(ctypes.c_double))[0] = total


class Thomas(VerticalAdvectionMixin, base.VerticalAdvectionStencil):
    alignment = Parameter("data alignment in bytes", 128)
