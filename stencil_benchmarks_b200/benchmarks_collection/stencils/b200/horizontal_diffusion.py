"""Horizontal diffusion on the B200: one fused kernel (``csrc/hdiff.cu``).

Counterpart of the nine reference variants in
stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/horizontal_diffusion.py:50-128.
"""

import numpy as np

from .... import capi
from ....benchmark import Parameter
from ....tools import cabi
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

    #: ``distributed.PeerSlabs`` when this instance sweeps ONE J slab of a partitioned domain
    #: (``distributed.attach_neighbours``): the sweep then reads its j-halo rows from the
    #: neighbouring GPUs' slabs (``sb200_hdiff_peer``) and every run is ordered against the
    #: neighbours' uploads.  New with respect to the reference, which is single-GPU.
    peers = None

    def launch(self, pointers, dry_runs, time_ptr, stream, domain=None, rows=None):
        peers = self.peers
        if peers is None:
            self._lib.sb200_hdiff(
                self._dtype_code, pointers["inp"], pointers["coeff"], pointers["out"],
                *self.geometry(domain), dry_runs, time_ptr, _vp(stream),
            )
            return
        # a launch over rows [j0, j1) of the slab needs the lower neighbour only if it starts at
        # the slab's first row, the upper one only if it ends at its last row
        j0, j1 = rows if rows is not None else (0, int(self.domain[1]))
        lower = peers.lower if j0 == 0 else None
        upper = peers.upper if j1 == int(self.domain[1]) else None
        self._lib.sb200_hdiff_peer(
            self._dtype_code, pointers["inp"], pointers["coeff"], pointers["out"],
            _vp(lower), peers.ny_lower, peers.sz_lower, _vp(upper), peers.ny_upper, peers.sz_upper,
            *self.geometry(domain), dry_runs, time_ptr, _vp(stream),
        )

    def _before_sweeps(self, data, mirrors, stream):
        """Partitioned slab: the two edge rows on each side go up first and every rank waits
        until its neighbours' edge rows are in their HBM -- after that the slab-pipelined
        upload / sweep / download runs exactly as on a single GPU."""
        if self.peers is None:
            return
        host = data.inp
        size = host.itemsize
        sy, sz = int(self.strides[1]), int(self.strides[2])
        ny, nz = int(self.domain[1]), int(self.domain[2])
        hy, hk = int(self.halo[1]), int(self.halo[2])
        width = min(self.j_reach, ny)
        handle = stream if isinstance(stream, _vp) else _vp(stream)
        for first in (hy, hy + ny - width):
            offset = (hk * sz + first * sy) * size
            status = self._lib.raw.sb200_memcpy2d_h2d(
                _vp(mirrors["inp"][1] + offset), sz * size, _vp(host.ctypes.data + offset), sz * size,
                width * sy * size, nz, handle)
            if status != 0:
                raise cabi.ExecutionError("uploading the edge rows of a partitioned slab failed (see stderr)")
        capi.synchronize(handle.value)
        self.peers.barrier()

    def _after_sweeps(self):
        """No neighbour may overwrite its edge rows (the next run's upload) while this rank still
        sweeps: the run ends with a barrier among the ranks."""
        if self.peers is not None:
            capi.synchronize()
            self.peers.barrier()


class Fused(HorizontalDiffusionMixin, base.HorizontalDiffusionStencil):
    alignment = Parameter("data alignment in bytes", 128)
