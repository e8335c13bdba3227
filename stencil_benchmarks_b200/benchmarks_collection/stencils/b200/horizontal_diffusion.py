"""Horizontal diffusion on the B200: one fused kernel (``csrc/hdiff.cu``).

Counterpart of the nine reference variants in
stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/horizontal_diffusion.py:50-128.
"""

import ctypes
import weakref

import numpy as np

from .... import capi
from ....benchmark import ExecutionError, Parameter, ParameterError
from ....tools import cabi
from .. import base
from .mixin import SlabCopies, StencilMixin, _vp


class HorizontalDiffusionMixin(StencilMixin):
    block_size = Parameter("block size of the reference's templates (no effect)", (32, 8, 1))

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
            self._kernels.sb200_hdiff(
                self._dtype_code, pointers["inp"], pointers["coeff"], pointers["out"],
                *self.geometry(domain), dry_runs, time_ptr, _vp(stream),
            )
            return
        # a launch over rows [j0, j1) of the slab needs the lower neighbour only if it starts at
        # the slab's first row, the upper one only if it ends at its last row
        j0, j1 = rows if rows is not None else (0, int(self.domain[1]))
        lower = peers.lower if j0 == 0 else None
        upper = peers.upper if j1 == int(self.domain[1]) else None
        self._kernels.sb200_hdiff_peer(
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


class Partitioned(SlabCopies, HorizontalDiffusionMixin, base.HorizontalDiffusionStencil):
    """Horizontal diffusion of ONE domain partitioned over ``gpus`` GPUs, driven by this process
    (``sbench stencils b200 horizontal-diffusion partitioned --gpus 8``).

    New with respect to the reference, which is single-GPU.  The IJ plane is cut
    into contiguous J slabs (``distributed.split_rows``), one per device; every device holds its
    rows of the three fields (same row pitch as the host array, its own level pitch).  A sweep
    is one ``sb200_hdiff_peer`` launch per device: the two halo rows on either side of a slab are
    read by the kernel itself from the neighbouring device's slab over NVLink (peer access inside
    one process, no IPC), so the j-halo rows at internal slab boundaries are never uploaded.
    ``run()`` scatters the host fields, sweeps, gathers ``out`` -- and, with ``verify=True``
    under the reference package, the reference's NumPy oracle checks the gathered global field.
    ``time`` is the longest of the per-device sweep times (CUDA events), all devices started
    together; ``bandwidth`` therefore is the aggregate over the GPUs.
    """

    alignment = Parameter("data alignment in bytes", 128)
    gpus = Parameter("number of GPUs the domain is partitioned over (J slabs)", 2)

    def setup(self):
        super().setup()
        if self.gpus < 1:
            raise ParameterError("gpus must be at least 1")
        if int(self.domain[1]) < 2 * self.gpus:
            raise ParameterError("every J slab needs at least two rows")
        if self.chunks != 1 or self.resident:
            raise ParameterError("chunks / resident are not offered by the partitioned benchmark")
        if self.alignment % 16 or self.alignment == 0:
            raise ParameterError("the partitioned sweep reads its halos by TMA: alignment must be a multiple of 16")
        self._slabs = None

    # ---- device side -------------------------------------------------------------------
    def _partition(self, data):
        """Per device: rows, buffers and interior pointers of the three fields (allocated once)."""
        if self._slabs is not None:
            return self._slabs
        from .... import distributed

        capi.require_device()
        if capi.device_count() < self.device + self.gpus:
            raise cabi.ExecutionError(
                f"{self.gpus} GPUs requested from device {self.device} on, {capi.device_count()} present")
        lib = self._lib
        size = data.inp.itemsize
        sy = int(self.strides[1])
        hx, hy, hk = (int(h) for h in self.halo)
        nz = int(self.domain[2])
        slabs = []
        for index, (start, ny) in enumerate(distributed.split_rows(int(self.domain[1]), self.gpus)):
            device = self.device + index
            lib.sb200_set_device(device)
            sz = sy * (ny + 2 * hy)
            interior = hx + hy * sy + hk * sz
            slab = dict(device=device, start=start, ny=ny, sz=sz, buffers={}, first={}, interior={})
            for name in self.args:
                buffer = capi.DeviceBuffer(sz * (nz + 2 * hk) * size + 512)
                first = buffer.ptr + (-(buffer.ptr + interior * size) % 256)
                slab["buffers"][name] = buffer
                slab["first"][name] = first
                slab["interior"][name] = first + interior * size
                # rows that are never uploaded (internal j halos) must not hold stale numbers
                lib.sb200_memset(_vp(first), 0xFF, sz * (nz + 2 * hk) * size, None, 1)
            # out: halo and padding are mirrored once, so that the gather of whole padded rows
            # keeps them as the reference's whole-field copies do (cuda_hip/mixin.py:142-160)
            self._copy_rows(slab, data.out, "out", 0, ny + 2 * hy, True, (0, nz + 2 * hk))
            capi.synchronize()
            slab["events"] = []
            for _ in range(2):
                event = _vp()
                lib.sb200_event_create(ctypes.byref(event))
                slab["events"].append(event)
                weakref.finalize(self, lib.raw.sb200_event_destroy, event)
            slabs.append(slab)
        for a, b in zip(slabs, slabs[1:]):
            lib.sb200_enable_peer_access(a["device"], b["device"])
            lib.sb200_enable_peer_access(b["device"], a["device"])
        lib.sb200_set_device(self.device)
        self._slabs = slabs
        return slabs

    def run_stencil(self, data):
        try:
            return self._run_partitioned(data)
        except cabi.ExecutionError as error:
            raise ExecutionError(*error.args) from error
        finally:
            self._lib.raw.sb200_set_device(self.device)

    def _run_partitioned(self, data):
        lib = self._lib
        kernels = self._kernels
        slabs = self._partition(data)
        nx, _, nz = (int(d) for d in self.domain)
        hy, hk = int(self.halo[1]), int(self.halo[2])
        sy = int(self.strides[1])
        last = len(slabs) - 1
        # scatter: own rows plus the j halo where it is the GLOBAL boundary; out travels too so
        # that its halo rows survive the gather as they do in the reference (whole-field copies)
        for index, slab in enumerate(slabs):
            lib.sb200_set_device(slab["device"])
            first = 0 if index == 0 else hy
            stop = slab["ny"] + 2 * hy - (0 if index == last else hy)
            for name in ("inp", "coeff"):
                self._copy_rows(slab, getattr(data, name), name, first, stop - first, True, (hk, nz))
        self._sync_all(slabs)

        def sweep(slab, index, dry_runs):
            lower = slabs[index - 1] if index > 0 else None
            upper = slabs[index + 1] if index < last else None
            kernels.sb200_hdiff_peer(
                self._dtype_code, _vp(slab["interior"]["inp"]), _vp(slab["interior"]["coeff"]),
                _vp(slab["interior"]["out"]),
                _vp(lower["interior"]["inp"] if lower else None), lower["ny"] if lower else 0,
                lower["sz"] if lower else 0,
                _vp(upper["interior"]["inp"] if upper else None), upper["ny"] if upper else 0,
                upper["sz"] if upper else 0,
                nx, slab["ny"], nz, 1, sy, slab["sz"], dry_runs, None, None)

        if self.dry_runs:
            for index, slab in enumerate(slabs):
                lib.sb200_set_device(slab["device"])
                sweep(slab, index, self.dry_runs - 1)
            self._sync_all(slabs)
        for index, slab in enumerate(slabs):
            lib.sb200_set_device(slab["device"])
            lib.sb200_event_record(slab["events"][0], None)
            sweep(slab, index, 0)
            lib.sb200_event_record(slab["events"][1], None)
        self._sync_all(slabs)
        times = []
        for slab in slabs:
            lib.sb200_set_device(slab["device"])
            elapsed = ctypes.c_double()
            lib.sb200_event_elapsed(slab["events"][0], slab["events"][1], ctypes.byref(elapsed))
            times.append(elapsed.value)
        # gather the interior rows of out
        for slab in slabs:
            lib.sb200_set_device(slab["device"])
            self._copy_rows(slab, data.out, "out", hy, slab["ny"], False, (hk, nz))
        self._sync_all(slabs)
        longest = max(times)
        return {"time": longest, "time-per-gpu-min": min(times), "gpus": len(slabs),
                "bandwidth-algorithmic": self.algorithmic_bytes / longest / 1e9}
