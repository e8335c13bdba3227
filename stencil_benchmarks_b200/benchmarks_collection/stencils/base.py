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

    _ELEMENT = np.dtype  # element type of a field, from its NumPy name

    class Stencil(Benchmark):
        """Fields on a regular (i, j, k) grid with a halo; a sweep reads some and writes others."""

        domain = Parameter("number of interior points along i, j, k", (128, 128, 80))
        data_sets = Parameter(
            "independent sets of fields; run() cycles through them so that every sweep "
            "starts from data that is not in cache", 1)
        halo = Parameter("halo width along i, j, k", (3, 3, 3))
        dtype = Parameter("element type, a NumPy name such as float32 or float64", "float64")
        layout = Parameter("stride order of the axes: 2 = unit stride, 0 = largest stride", (2, 1, 0))
        alignment = Parameter("byte alignment of the first interior element and of the rows", 0)
        huge_pages = Parameter("huge-page policy of the host allocator (accepted, unused here)", "none",
                               choices=["none", "transparent", "explicit"])
        offset_allocations = Parameter(
            "stagger allocations against cache-set conflicts (accepted, unused here)", False)
        # the oracle is the reference package's verify_stencil (base.py:151-166); this stand-alone
        # mirror has no CPU implementation of any stencil, so the default is off here and on
        # (the reference's default) whenever the reference package is importable
        verify = Parameter("check every sweep against the NumPy oracle of the reference package", False)

        def setup(self):
            super().setup()
            problems = [
                (min(self.halo) < 0, f"negative halo size given ({self.halo}"),
                (sorted(self.layout) != [0, 1, 2], f"invalid layout specification {self.layout}"),
                (self.alignment < 0, f"negative alignment given ({self.alignment} bytes)"),
                (self.alignment % self.dtype_size != 0,
                 f"alignment ({self.alignment} bytes) not divisible by dtype size "
                 f"({self.dtype_size} bytes)"),
                (self.verify,
                 "verify=True validates against the NumPy oracle of the reference package "
                 "(stencil_benchmarks...stencils/base.py verify_stencil), which is not "
                 "importable here; pass verify=False"),
            ]
            for failed, message in problems:
                if failed:
                    raise ParameterError(message)
            record = collections.namedtuple("StencilData", self.args)
            self._data = []
            for _ in range(self.data_sets):
                self._data.append(record(*[self.random_field() for _ in self.args]))
            self._run = 0

        # -- allocation hooks (backends override them) --
        def alloc_field(self, domain_with_halo, layout, index_to_align):
            return fields.alloc_array(domain_with_halo, self.dtype, layout, self.alignment,
                                      index_to_align=index_to_align)

        def empty_field(self):
            return self.alloc_field(self.domain_with_halo, self.layout, self.halo)

        def random_field(self):
            field = self.empty_field()
            field[...] = np.random.default_rng().random(field.shape).astype(field.dtype)
            return field

        # -- geometry --
        @property
        def dtype_size(self):
            return _ELEMENT(self.dtype).itemsize

        @property
        def domain_with_halo(self):
            return tuple(n + 2 * h for n, h in zip(self.domain, self.halo))

        @property
        def strides(self):
            """Element strides of the (identically allocated) fields."""
            first = self._data[0][0]
            return tuple(step // first.itemsize for step in first.strides)

        @property
        def data_size(self):
            """Bytes a sweep is charged with: every field once over the interior."""
            return len(self.args) * int(np.prod(self.domain)) * self.dtype_size

        def inner_slice(self, shift=None, expand=None):
            """Index of the interior, optionally moved by `shift` and grown by `expand` points."""
            rank = len(self.domain)
            shift = (0,) * rank if shift is None else tuple(shift)
            grow = (0,) * rank if expand is None else ((expand,) * rank if isinstance(expand, int)
                                                       else tuple(expand))
            index = []
            for axis in range(rank):
                begin = self.halo[axis] + shift[axis] - grow[axis]
                index.append(slice(begin, begin + self.domain[axis] + 2 * grow[axis]))
            return tuple(index)

        # -- protocol --
        @abc.abstractmethod
        def run_stencil(self, data):
            """One sweep on the host fields `data`; returns {"time": seconds}."""

        def verify_stencil(self, data_before, data_after):
            raise ParameterError("no verification without the reference package")

        @abc.abstractproperty
        def args(self):
            """Field names in the order the kernels take them."""

        def run(self):
            fields_now = self._data[self._run % self.data_sets]
            snapshot = copy.deepcopy(fields_now) if self.verify else None
            outcome = self.run_stencil(fields_now)
            if snapshot is not None:
                self.verify_stencil(snapshot, fields_now)
            self._run += 1
            if not outcome.get("time", 0) > 0 or "bandwidth" in outcome:
                raise AssertionError("run_stencil must return a positive time and no bandwidth")
            outcome["bandwidth"] = self.data_size / outcome["time"] / 1e9
            return outcome

    class BasicStencil(Stencil):
        """out = f(inp)."""

        @property
        def args(self):
            return ("inp", "out")

    class EmptyStencil(BasicStencil):
        pass

    class CopyStencil(BasicStencil):
        pass

    _AXIS = dict(description="axis of the neighbour(s): 0 = i, 1 = j, 2 = k", default=0, choices=[0, 1, 2])

    class OnesidedAverageStencil(BasicStencil):
        axis = Parameter(**_AXIS)

    class SymmetricAverageStencil(BasicStencil):
        axis = Parameter(**_AXIS)

    class LaplacianStencil(BasicStencil):
        along_x = Parameter("second difference along i", True)
        along_y = Parameter("second difference along j", True)
        along_z = Parameter("second difference along k", False)

        def setup(self):
            super().setup()
            for width, active in zip(self.halo, (self.along_x, self.along_y, self.along_z)):
                if active and width < 1:
                    raise ParameterError(
                        f"positive horizontal halo size required (given halo: {self.halo})")

    class HorizontalDiffusionStencil(Stencil):
        """out = inp - coeff * div(limited fluxes of the Laplacian of inp)."""

        def setup(self):
            super().setup()
            if min(self.halo[0], self.halo[1]) < 2:
                raise ParameterError(
                    f"horizontal halo size must be at least 2 (given halo: {self.halo})")

        @property
        def args(self):
            return ("inp", "coeff", "out")

        @property
        def data_size(self):
            # the reference's accounting (base.py:270-274): coeff + out on the interior, inp on the
            # interior grown by 2 -- on k as well, a quirk kept for comparable CSVs
            interior = int(np.prod(self.domain))
            grown = int(np.prod([n + 4 for n in self.domain]))
            return (2 * interior + grown) * self.dtype_size

    class VerticalAdvectionStencil(Stencil):
        """Implicit vertical advection: one tridiagonal solve per (i, j) column."""

        all_components = Parameter(
            "solve for u, v and w (COSMO dycore) instead of u only (GridTools benchmark)", False)

        def setup(self):
            super().setup()
            needed = self.halo if self.all_components else self.halo[:1]
            if min(needed) < 1:
                raise ParameterError(f"positive halo size required (given halo: {self.halo})")

        @property
        def args(self):
            names = []
            for component in ("u", "v", "w") if self.all_components else ("u",):
                names += [component + suffix for suffix in ("stage", "pos", "tens", "tensstage")]
            return tuple(names) + ("wcon", "ccol", "dcol", "datacol")

        @property
        def data_size(self):
            # the reference's accounting (base.py:339-347) charges the ccol / dcol round trips
            fields_charged = 20 if self.all_components else 10
            return fields_charged * int(np.prod(self.domain)) * self.dtype_size
