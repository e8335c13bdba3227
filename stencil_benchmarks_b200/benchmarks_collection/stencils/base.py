"""Abstract stencil definitions: field lists, byte accounting, halo rules.

With the reference importable these ARE the reference's classes
(stencil_benchmarks/benchmarks_collection/stencils/base.py), including their
NumPy ``verify_stencil`` oracles, so ``verify=True`` validates every run of a
B200 kernel against the reference's own NumPy implementation exactly like its
other backends (base.py:151-166).

Stand-alone, the classes below restate the *definitions* only
(base.py:45-166 ``Stencil``; :169-254 basic stencils; :257-274 horizontal
diffusion; :314-347 vertical advection): parameters, ``args``, ``data_size``,
``inner_slice``, halo requirements and the ``run()`` protocol.  They contain
no CPU implementation of any stencil: ``verify=True`` needs the reference and
raises ``ParameterError`` without it (parity is then established by the test
suite against ``oracle/``).
"""

from ...benchmark import HAVE_REFERENCE

if HAVE_REFERENCE:
    from stencil_benchmarks.benchmarks_collection.stencils.base import (  # noqa: F401
        BasicStencil,
        CopyStencil,
        EmptyStencil,
        HorizontalDiffusionStencil,
        LaplacianStencil,
        OnesidedAverageStencil,
        Stencil,
        SymmetricAverageStencil,
        VerticalAdvectionStencil,
    )
else:
    import abc
    import collections
    import copy

    import numpy as np

    from ...benchmark import Benchmark, Parameter, ParameterError
    from ...tools import fields

    # pylint: disable=abstract-method

    class Stencil(Benchmark):
        domain = Parameter("domain size", (128, 128, 80))
        data_sets = Parameter(
            "number of data sets, if bigger than one, data sets are cycled before "
            "each execution to start with cold cache",
            1,
        )
        halo = Parameter("halo size", (3, 3, 3))
        dtype = Parameter("data type in NumPy format, e.g. float32 or float64", "float64")
        layout = Parameter("data layout, 2 means innermost dimension, 0 outermost", (2, 1, 0))
        alignment = Parameter("data alignment in bytes", 0)
        huge_pages = Parameter("use huge pages", "none", choices=["none", "transparent", "explicit"])
        offset_allocations = Parameter(
            "offset allocated data by some bytes to minimize cache conflicts", False
        )
        verify = Parameter("enable verification", True)

        def setup(self):
            super().setup()
            if any(h < 0 for h in self.halo):
                raise ParameterError(f"negative halo size given ({self.halo}")
            if tuple(sorted(self.layout)) != (0, 1, 2):
                raise ParameterError(f"invalid layout specification {self.layout}")
            if self.alignment < 0:
                raise ParameterError(f"negative alignment given ({self.alignment} bytes)")
            if self.alignment % self.dtype_size != 0:
                raise ParameterError(
                    f"alignment ({self.alignment} bytes) not divisible "
                    f"by dtype size ({self.dtype_size} bytes)"
                )
            if self.verify:
                raise ParameterError(
                    "verify=True validates against the NumPy oracle of the reference package "
                    "(stencil_benchmarks...stencils/base.py verify_stencil), which is not "
                    "importable here; pass verify=False"
                )
            record = collections.namedtuple("StencilData", self.args)
            self._data = [
                record._make(self.random_field() for _ in self.args)
                for _ in range(self.data_sets)
            ]
            self._run = 0

        def alloc_field(self, domain_with_halo, layout, index_to_align):
            return fields.alloc_array(
                domain_with_halo, self.dtype, layout, self.alignment, index_to_align=index_to_align
            )

        def empty_field(self):
            return self.alloc_field(self.domain_with_halo, self.layout, self.halo)

        def random_field(self):
            data = self.empty_field()
            data[...] = np.random.default_rng().random(data.shape, dtype=data.dtype)
            return data

        @property
        def dtype_size(self):
            return np.dtype(self.dtype).itemsize

        @property
        def domain_with_halo(self):
            return tuple(d + 2 * h for d, h in zip(self.domain, self.halo))

        @property
        def strides(self):
            return tuple(s // self.dtype_size for s in self._data[0][0].strides)

        @property
        def data_size(self):
            return len(self.args) * np.prod(self.domain) * self.dtype_size

        def inner_slice(self, shift=None, expand=None):
            ndim = len(self.domain)
            shift = [0] * ndim if shift is None else shift
            if expand is None:
                expand = [0] * ndim
            elif isinstance(expand, int):
                expand = [expand] * ndim
            return tuple(
                slice(h + s - e, h + d + s + e)
                for d, h, s, e in zip(self.domain, self.halo, shift, expand)
            )

        @abc.abstractmethod
        def run_stencil(self, data):
            pass

        def verify_stencil(self, data_before, data_after):
            raise ParameterError("no verification without the reference package")

        @abc.abstractproperty
        def args(self):
            pass

        def run(self):
            data = self._data[self._run % self.data_sets]
            before = copy.deepcopy(data) if self.verify else None
            result = self.run_stencil(data)
            if self.verify:
                self.verify_stencil(before, data)
            self._run += 1
            assert "time" in result and result["time"] > 0
            assert "bandwidth" not in result
            result["bandwidth"] = self.data_size / result["time"] / 1e9
            return result

    class BasicStencil(Stencil):
        @property
        def args(self):
            return "inp", "out"

    class EmptyStencil(BasicStencil):
        pass

    class CopyStencil(BasicStencil):
        pass

    class OnesidedAverageStencil(BasicStencil):
        axis = Parameter("axis along which to average", 0, choices=[0, 1, 2])

    class SymmetricAverageStencil(BasicStencil):
        axis = Parameter("axis along which to average", 0, choices=[0, 1, 2])

    class LaplacianStencil(BasicStencil):
        along_x = Parameter("include x-axis in Laplacian", True)
        along_y = Parameter("include y-axis in Laplacian", True)
        along_z = Parameter("include z-axis in Laplacian", False)

        def setup(self):
            super().setup()
            active = (self.along_x, self.along_y, self.along_z)
            if any(h < 1 for h, a in zip(self.halo, active) if a):
                raise ParameterError(
                    f"positive horizontal halo size required (given halo: {self.halo})"
                )

    class HorizontalDiffusionStencil(Stencil):
        def setup(self):
            super().setup()
            if any(h < 2 for h in self.halo[:2]):
                raise ParameterError(
                    f"horizontal halo size must be at least 2 (given halo: {self.halo})"
                )

        @property
        def args(self):
            return "inp", "coeff", "out"

        @property
        def data_size(self):
            # the reference's accounting (base.py:270-274), +4 on k included
            return (
                2 * np.prod(self.domain) + np.prod(np.array(self.domain) + 4)
            ) * self.dtype_size

    class VerticalAdvectionStencil(Stencil):
        all_components = Parameter(
            "advect all velocity components (like in the COSMO dycore) "
            "instead of the u component (like in the GridTools benchmark)",
            False,
        )

        def setup(self):
            super().setup()
            if self.halo[0] < 1 or (self.all_components and any(h < 1 for h in self.halo)):
                raise ParameterError(f"positive halo size required (given halo: {self.halo})")

        @property
        def args(self):
            u = ("ustage", "upos", "utens", "utensstage")
            v = ("vstage", "vpos", "vtens", "vtensstage")
            w = ("wstage", "wpos", "wtens", "wtensstage")
            common = ("wcon", "ccol", "dcol", "datacol")
            return u + common if not self.all_components else u + v + w + common

        @property
        def data_size(self):
            # the reference's accounting (base.py:339-347): ccol + dcol round trips counted
            reads, writes = (7, 3) if not self.all_components else (15, 5)
            return (reads + writes) * np.prod(self.domain) * self.dtype_size
