"""Basic stencils on the B200: empty, copy, one-sided / symmetric average, Laplacian.

Counterparts of the reference classes in
stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/basic.py:101-132; the
loop kind / block-size parameters of the reference (basic.py:44-48) are accepted
for script compatibility and have no effect on the fixed sm_100a kernels.
"""

from .... import capi
from ....benchmark import Parameter, ParameterError
from .. import base
from .mixin import StencilMixin, _vp

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
