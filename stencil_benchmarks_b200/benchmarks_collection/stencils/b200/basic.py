"""Basic stencils on the B200: empty, copy, one-sided / symmetric average, Laplacian -- and the
same four partitioned over several GPUs (``Partitioned*``).

Counterparts of the reference classes in
stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/basic.py:101-132; the
loop kind / block-size parameters of the reference (basic.py:44-48) are accepted
for script compatibility and have no effect on the fixed sm_100a kernels.
"""

import ctypes
import weakref

from .... import capi, distributed
from ....benchmark import ExecutionError, Parameter, ParameterError
from ....tools import cabi
from .. import base
from .mixin import SlabCopies, StencilMixin, _vp

_ALIGNMENT = Parameter("data alignment in bytes", 128)


class BasicStencilMixin(StencilMixin):
    loop = Parameter("loop kind (no effect)", "1D", choices=["1D", "3D"])
    block_size = Parameter("block size (no effect)", (1024, 1, 1))
    threads_per_block = Parameter("threads per block (no effect)", (0, 0, 0))

    field_roles = {"inp": "in", "out": "out"}
    kernel_source = "basic.cu"
    kind = capi.BASIC_EMPTY

    def axis_and_mask(self):
        return 0, 0

    def launch(self, pointers, dry_runs, time_ptr, stream, domain=None, rows=None):
        axis, mask = self.axis_and_mask()
        self._kernels.sb200_basic(
            self.kind, self._dtype_code, pointers["inp"], pointers["out"], *self.geometry(domain),
            axis, mask, dry_runs, time_ptr, _vp(stream),
        )


class Empty(BasicStencilMixin, base.EmptyStencil):
    alignment = _ALIGNMENT
    kind = capi.BASIC_EMPTY


class Copy(BasicStencilMixin, base.CopyStencil):
    alignment = _ALIGNMENT
    kind = capi.BASIC_COPY


class _AverageMixin(BasicStencilMixin):
    def setup(self):
        # the reference has no such check; reading beyond
        # the allocation must not happen here
        if self.halo[self.axis] < 1:
            raise ParameterError(
                f"positive halo size required along axis {self.axis} (given halo: {self.halo})"
            )
        super().setup()

    def axis_and_mask(self):
        return self.axis, 0

    @property
    def j_reach(self):
        return int(self.axis == 1)

    @property
    def k_reach(self):
        return int(self.axis == 2)


class OnesidedAverage(_AverageMixin, base.OnesidedAverageStencil):
    alignment = _ALIGNMENT
    kind = capi.BASIC_ONESIDED_AVG


class SymmetricAverage(_AverageMixin, base.SymmetricAverageStencil):
    alignment = _ALIGNMENT
    kind = capi.BASIC_SYMMETRIC_AVG


class Laplacian(BasicStencilMixin, base.LaplacianStencil):
    alignment = _ALIGNMENT
    kind = capi.BASIC_LAPLACIAN

    def setup(self):
        if not (self.along_x or self.along_y or self.along_z):
            raise ParameterError("Laplacian needs at least one axis")
        super().setup()

    def axis_and_mask(self):
        return 0, int(self.along_x) | int(self.along_y) << 1 | int(self.along_z) << 2

    @property
    def j_reach(self):
        return int(self.along_y)

    @property
    def k_reach(self):
        return int(self.along_z)


class _PartitionedMixin(SlabCopies, BasicStencilMixin):
    """A basic stencil on ONE domain partitioned over ``gpus`` GPUs driven by this process
    (``sbench stencils b200 basic partitioned-laplacian --gpus 8``; SURVEY.md §8e, "basic avg /
    Laplacian: one halo layer").

    New with respect to the reference, which is single-GPU.  The IJ plane is cut into contiguous J
    slabs (``distributed.split_rows``), one per device; every device holds its rows of ``inp`` and
    ``out`` with the host's row pitch.  The inputs of a sweep live on the host, so the halo rows of a
    slab -- rows of the neighbouring slabs -- arrive with the scatter (no device-to-device step
    exists for a single sweep of host data; the sweep-to-sweep exchange of a time loop is the
    horizontal-diffusion machinery, ``distributed.TimeLoop``).  One ``sb200_basic`` launch per
    device, all started together; ``time`` is the longest per-device sweep (CUDA events), so
    ``bandwidth`` is the aggregate over the GPUs; ``out`` is gathered, and with ``verify=True``
    under the reference package its NumPy oracle checks the gathered global field.
    """

    gpus = Parameter("number of GPUs the domain is partitioned over (J slabs)", 2)

    def setup(self):
        super().setup()
        if self.gpus < 1:
            raise ParameterError("gpus must be at least 1")
        if int(self.domain[1]) < self.gpus:
            raise ParameterError("every J slab needs at least one row")
        if self.chunks != 1 or self.resident:
            raise ParameterError("chunks / resident are not offered by the partitioned benchmarks")
        self._slabs = None

    def _partition(self, data):
        """Per device: rows and buffers of inp / out (allocated once)."""
        if self._slabs is not None:
            return self._slabs
        capi.require_device()
        if capi.device_count() < self.device + self.gpus:
            raise cabi.ExecutionError(
                f"{self.gpus} GPUs requested from device {self.device} on, {capi.device_count()} present")
        lib = self._lib
        size = data.inp.itemsize
        sy = int(self.strides[1])
        hx, hy, hk = (int(h) for h in self.halo)
        levels = int(self.domain[2]) + 2 * hk
        slabs = []
        for index, (start, ny) in enumerate(distributed.split_rows(int(self.domain[1]), self.gpus)):
            device = self.device + index
            lib.sb200_set_device(device)
            sz = sy * (ny + 2 * hy)
            interior = hx + hy * sy + hk * sz
            slab = dict(device=device, start=start, ny=ny, sz=sz, buffers={}, first={}, interior={}, events=[])
            for name in self.args:
                buffer = capi.DeviceBuffer(sz * levels * size + 512)
                first = buffer.ptr + (-(buffer.ptr + interior * size) % 256)
                slab["buffers"][name] = buffer
                slab["first"][name] = first
                slab["interior"][name] = first + interior * size
            # out: halo and padding are mirrored once, so that whole padded rows can be gathered
            # without disturbing them (the reference copies whole fields, cuda_hip/mixin.py:142-160)
            self._copy_rows(slab, data.out, "out", 0, ny + 2 * hy, True, (0, levels))
            capi.synchronize()
            for _ in range(2):
                event = _vp()
                lib.sb200_event_create(ctypes.byref(event))
                slab["events"].append(event)
                weakref.finalize(self, lib.raw.sb200_event_destroy, event)
            slabs.append(slab)
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
        slabs = self._partition(data)
        nx, _, nz = (int(d) for d in self.domain)
        hy, hk = int(self.halo[1]), int(self.halo[2])
        sy = int(self.strides[1])
        levels = nz + 2 * hk
        axis, mask = self.axis_and_mask()
        # scatter: every slab with its j halo -- the neighbours' rows, or the global boundary's
        for slab in slabs:
            lib.sb200_set_device(slab["device"])
            self._copy_rows(slab, data.inp, "inp", 0, slab["ny"] + 2 * hy, True, (0, levels))
        self._sync_all(slabs)

        def sweep(slab, dry_runs):
            self._kernels.sb200_basic(self.kind, self._dtype_code, _vp(slab["interior"]["inp"]),
                                      _vp(slab["interior"]["out"]), nx, slab["ny"], nz, 1, sy, slab["sz"],
                                      axis, mask, dry_runs, None, None)

        if self.dry_runs:
            for slab in slabs:
                lib.sb200_set_device(slab["device"])
                sweep(slab, self.dry_runs - 1)
            self._sync_all(slabs)
        for slab in slabs:
            lib.sb200_set_device(slab["device"])
            lib.sb200_event_record(slab["events"][0], None)
            sweep(slab, 0)
            lib.sb200_event_record(slab["events"][1], None)
        self._sync_all(slabs)
        times = []
        for slab in slabs:
            lib.sb200_set_device(slab["device"])
            elapsed = ctypes.c_double()
            lib.sb200_event_elapsed(slab["events"][0], slab["events"][1], ctypes.byref(elapsed))
            times.append(elapsed.value)
        for slab in slabs:  # gather the interior rows of out
            lib.sb200_set_device(slab["device"])
            self._copy_rows(slab, data.out, "out", hy, slab["ny"], False, (hk, nz))
        self._sync_all(slabs)
        longest = max(times)
        return {"time": longest, "time-per-gpu-min": min(times), "gpus": len(slabs),
                "bandwidth-algorithmic": self.algorithmic_bytes / longest / 1e9}


class PartitionedCopy(_PartitionedMixin, base.CopyStencil):
    alignment = _ALIGNMENT
    kind = capi.BASIC_COPY


class _PartitionedAverage(_PartitionedMixin):
    def setup(self):
        if self.halo[self.axis] < 1:
            raise ParameterError(
                f"positive halo size required along axis {self.axis} (given halo: {self.halo})"
            )
        super().setup()

    def axis_and_mask(self):
        return self.axis, 0


class PartitionedOnesidedAverage(_PartitionedAverage, base.OnesidedAverageStencil):
    alignment = _ALIGNMENT
    kind = capi.BASIC_ONESIDED_AVG


class PartitionedSymmetricAverage(_PartitionedAverage, base.SymmetricAverageStencil):
    alignment = _ALIGNMENT
    kind = capi.BASIC_SYMMETRIC_AVG


class PartitionedLaplacian(_PartitionedMixin, base.LaplacianStencil):
    alignment = _ALIGNMENT
    kind = capi.BASIC_LAPLACIAN

    def setup(self):
        if not (self.along_x or self.along_y or self.along_z):
            raise ParameterError("Laplacian needs at least one axis")
        super().setup()

    def axis_and_mask(self):
        return 0, int(self.along_x) | int(self.along_y) << 1 | int(self.along_z) << 2
