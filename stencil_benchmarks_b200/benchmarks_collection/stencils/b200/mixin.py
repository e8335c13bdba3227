"""Backend mixin of the B200 stencils: host fields, device mirrors, C-ABI calls.

Counterpart of the reference's ``cuda_hip.StencilMixin``
(stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/mixin.py:50-175):

=====================  =====================================================
reference               here
=====================  =====================================================
render Jinja + nvcc     pre-built ``libsbench_b200.so`` (sm_100a only), or --
at ``setup()``          with ``compiler`` given -- the same sources compiled at
(mixin.py:60-84)        ``setup()`` through ``tools.compilation.GnuLibrary``; a
                        missing library / failed compilation is a
                        ``ParameterError`` as it is there
``on_device``           ``_device_fields``: device mirrors with the host's
(mixin.py:125-160):     strides, allocated once per data set; per run H2D of
cudaMalloc + H2D of     the fields the stencil READS and D2H of the fields it
all fields, D2H of      WRITES (scratch fields never cross PCIe)
all fields, every run
``kernel(&time, ptrs)`` ``sb200_<stencil>(..., dry_runs, &time, stream)`` with
(mixin.py:162-175)      pointers to the first interior element
=====================  =====================================================

Error mapping is the reference's: library problems -> ``ParameterError`` (so
``sbench -s`` can skip), a failing C call -> ``benchmark.ExecutionError``.
"""

import ctypes
import time as _time

import numpy as np

from .... import capi
from ....benchmark import Benchmark, ExecutionError, Parameter, ParameterError
from ....tools import cabi, fields

_vp = ctypes.c_void_p


class StencilMixin(Benchmark):
    # names kept from the reference mixin (mixin.py:51-58) where they still mean something
    compiler = Parameter(
        "compiler path: if given, the kernels are compiled at setup() through the reference's "
        "tools.compilation path (as its cuda_hip backend does); empty: the prebuilt library", "")
    compiler_flags = Parameter("additional compiler flags (with `compiler`)", "")
    backend = Parameter("GPU programming model (CUDA only)", "cuda", choices=["cuda"])
    gpu_architecture = Parameter("GPU architecture (Blackwell B200 only)", "sm_100a",
                                 choices=["sm_100a"])
    print_code = Parameter("print the CUDA source of the kernel", False)
    dry_runs = Parameter("kernel dry-runs before the measurement", 0)
    timers = Parameter("timer type", default="gpu", choices=["gpu", "wall"])
    # tuning parameters of the reference's templates (cuda_hip/mixin.py:58 and the per-stencil
    # mixins): accepted so that scripts written for `stencils cuda-hip ...` run unchanged.  The
    # hand-written kernels have fixed, measured launch shapes; these have no effect.
    index_type = Parameter("index data type (no effect)", "std::ptrdiff_t")
    # the fast (128-bit) kernels need an aligned interior origin and aligned rows
    alignment = Parameter("data alignment in bytes", 128)
    # new
    device = Parameter("CUDA device ordinal", 0)
    pinned = Parameter("allocate host fields in page-locked memory", True)
    seed = Parameter("seed of the random input fields (negative: non-deterministic)", 42)
    chunks = Parameter(
        "j-slabs over which host<->device copies and the sweep are pipelined on three streams "
        "(1: upload everything, sweep once, download -- the reference's sequence)", 1)

    resident = Parameter(
        "keep the fields resident in HBM between runs: upload once, download only what "
        "verification needs (the reference copies every field both ways on every run)", False)

    #: role of each field: "in" (H2D every run), "out" (D2H every run),
    #: "inout" (both), "scratch" (device only)
    field_roles = {}
    #: CUDA source shown by --print-code
    kernel_source = None
    #: rows / levels of the INPUT fields a sweep reads beyond the rows / levels it writes
    j_reach = 0
    k_reach = 0

    def setup(self):
        if tuple(self.layout) != (2, 1, 0):
            raise ParameterError(
                f"layout {tuple(self.layout)} is not supported by the B200 kernels "
                "(i must be the unit-stride axis: layout (2, 1, 0))"
            )
        try:
            self._dtype_code = capi.dtype_code(self.dtype)
        except (ValueError, TypeError) as error:
            raise ParameterError(str(error)) from error
        try:
            self._lib = capi.library()
            # the C entry point of the stencil itself: from the prebuilt library, or compiled now
            self._kernels = (capi.jit_library(self.compiler, self.compiler_flags, self.kernel_source)
                             if self.compiler else self._lib)
        except cabi.CompilationError as error:
            raise ParameterError(*error.args) from error
        if self.chunks < 1:
            raise ParameterError("chunks must be at least 1")
        if self.pinned:
            try:
                capi.require_device()
                self._lib.sb200_set_device(self.device)
            except cabi.ExecutionError as error:
                raise ParameterError(*error.args) from error
        self._field_counter = 0
        self._device = {}

        super().setup()

        if self.print_code and self.kernel_source:
            print((capi.ROOT / "csrc" / self.kernel_source).read_text())

    # ---- host fields -------------------------------------------------------
    def alloc_field(self, domain_with_halo, layout, index_to_align):
        return fields.alloc_array(
            domain_with_halo,
            self.dtype,
            layout,
            self.alignment,
            index_to_align=index_to_align,
            alloc=capi.pinned_alloc if self.pinned else None,
        )

    def random_field(self):
        """U[0,1) over the whole padded field (base.py:106-109), reproducible per seed."""
        data = self.empty_field()
        seed = None if self.seed < 0 else [self.seed, self._field_counter]
        self._field_counter += 1
        rng = np.random.default_rng(seed)
        # plane by plane: keeps the temporary small for multi-GB fields
        for k in range(data.shape[2]):
            data[:, :, k] = rng.random(data.shape[:2], dtype=np.float64).astype(data.dtype)
        return data

    def data(self, index=None):
        """Host fields (namedtuple in ``args`` order) of a data set (default: the next one)."""
        if index is None:
            index = self._run % self.data_sets
        return self._data[index]

    # ---- device mirrors ------------------------------------------------------
    def _device_fields(self, data):
        """Device mirrors of one data set: dict name -> (DeviceBuffer, pointer to element 0)."""
        key = id(data)
        if key in self._device:
            return self._device[key]
        capi.require_device()
        self._lib.sb200_set_device(self.device)
        mirrors = {}
        align = max(self.alignment, 256)
        for name, host in zip(self.args, data):
            nbytes = fields.nbytes(host)
            # whole padded rows, so that slab copies may include the padding of the last row
            extent = host.strides[2] * host.shape[2]
            buffer = capi.DeviceBuffer(max(nbytes, extent) + align)
            # keep the host's alignment of the first interior element
            interior = sum(s * h for s, h in zip(host.strides, self.halo))
            first = buffer.ptr + (-(buffer.ptr + interior) % align)
            mirrors[name] = (buffer, first)
            if self.field_roles.get(name, "inout") == "out":
                # written fields: halo and padding are mirrored once so D2H keeps them
                capi.memcpy_h2d(first, host.ctypes.data, nbytes)
        self._device[key] = mirrors
        return mirrors

    def interior_ptr(self, first, host):
        return _vp(first + sum(s * h for s, h in zip(host.strides, self.halo)))

    def upload(self, data, mirrors, stream=None):
        for name, host in zip(self.args, data):
            if self.field_roles.get(name, "inout") in ("in", "inout"):
                capi.memcpy_h2d(mirrors[name][1], host.ctypes.data, fields.nbytes(host), stream,
                                sync=False)
        capi.synchronize(stream)

    def download(self, data, mirrors, stream=None):
        for name, host in zip(self.args, data):
            if self.field_roles.get(name, "inout") in ("out", "inout"):
                capi.memcpy_d2h(host.ctypes.data, mirrors[name][1], fields.nbytes(host), stream,
                                sync=False)
        capi.synchronize(stream)

    # ---- the run protocol -------------------------------------------------------
    def launch(self, pointers, dry_runs, time_ptr, stream, domain=None, rows=None):
        """Call the stencil's C entry point; ``pointers`` maps field name -> interior void*;
        ``domain`` overrides the swept domain (a j-slab of the fields, whose first row the pointers
        address); ``rows`` = (first, last + 1) row of that slab inside the instance's domain."""
        raise NotImplementedError

    # hooks of a J-partitioned instance (one slab of a global domain, see HorizontalDiffusionMixin)
    def _before_sweeps(self, data, mirrors, stream):
        """Called once per run before the first sweep is enqueued (``stream``: the upload stream)."""

    def _after_sweeps(self):
        """Called once per run after the last sweep has completed."""

    @property
    def algorithmic_bytes(self):
        """Minimum HBM traffic of one sweep; defaults to the sbench figure."""
        return int(self.data_size)

    def transfer_bytes(self):
        """(H2D, D2H) bytes one run() moves, as counted from the copies it issues."""
        data = self._data[0]
        roles = [self.field_roles.get(name, "inout") for name in self.args]
        if self.chunks == 1:
            h2d = sum(fields.nbytes(f) for f, r in zip(data, roles) if r in ("in", "inout"))
            d2h = sum(fields.nbytes(f) for f, r in zip(data, roles) if r in ("out", "inout"))
            return h2d, d2h
        ny, nz = int(self.domain[1]), int(self.domain[2])
        size = np.dtype(self.dtype).itemsize
        sy = int(self.strides[1])
        jr, kr = min(self.j_reach, self.halo[1]), min(self.k_reach, self.halo[2])
        up = (ny + 2 * jr) * sy * (nz + 2 * kr) * size
        down = ny * sy * nz * size
        return (up * sum(r in ("in", "inout") for r in roles),
                down * sum(r in ("out", "inout") for r in roles))

    # ---- pipelined execution ---------------------------------------------------------
    def _pipeline(self):
        """Three streams (upload, sweep, download) and per-slab events, created once."""
        if getattr(self, "_pipe", None) is None:
            def handle(create):
                out = _vp()
                create(ctypes.byref(out))
                return out

            n = self.chunks
            self._pipe = dict(
                streams=[handle(self._lib.sb200_stream_create) for _ in range(3)],
                uploaded=[handle(self._lib.sb200_event_create) for _ in range(n)],
                begin=[handle(self._lib.sb200_event_create) for _ in range(n)],
                end=[handle(self._lib.sb200_event_create) for _ in range(n)],
            )
        return self._pipe

    def _run_pipelined(self, data, mirrors):
        """Upload slab c+1 while slab c is swept and slab c-1 is downloaded.

        Slabs are ranges of j (rows stay contiguous; a slab of one level is one row of a
        2-D copy whose pitch is the k stride).  Only the rows / levels a sweep reads are
        uploaded, only interior rows / levels of the written fields are downloaded.
        Returns the summed device time of the slab sweeps."""
        raw = self._lib.raw

        def check(status):
            if status != 0:
                raise cabi.ExecutionError("a CUDA call of the pipelined run failed (see stderr)")

        pipe = self._pipeline()
        s_up, s_run, s_down = pipe["streams"]
        nx, ny, nz = (int(d) for d in self.domain)
        hx, hy, hk = (int(h) for h in self.halo)
        size = np.dtype(self.dtype).itemsize
        _, sy, sz = (int(s) for s in self.strides)
        jr, kr = min(self.j_reach, hy), min(self.k_reach, hk)
        bounds = [ny * c // self.chunks for c in range(self.chunks + 1)]
        plane0, planes_in = hk - kr, nz + 2 * kr
        roles = [self.field_roles.get(name, "inout") for name in self.args]
        frontier = hy - jr  # first array row not yet uploaded
        self._before_sweeps(data, mirrors, s_up)
        for c in range(self.chunks):
            j0, j1 = bounds[c], bounds[c + 1]
            if j1 == j0:
                continue
            upto = j1 + hy + jr
            upto = min(upto, ny + hy + jr)
            for name, host, role in zip(self.args, data, roles):
                if role in ("in", "inout") and upto > frontier:
                    offset = (plane0 * sz + frontier * sy) * size
                    check(raw.sb200_memcpy2d_h2d(
                        _vp(mirrors[name][1] + offset), sz * size, _vp(host.ctypes.data + offset),
                        sz * size, (upto - frontier) * sy * size, planes_in, s_up))
            frontier = max(frontier, upto)
            check(raw.sb200_event_record(pipe["uploaded"][c], s_up))
            check(raw.sb200_stream_wait_event(s_run, pipe["uploaded"][c]))
            pointers = {
                name: _vp(self.interior_ptr(mirrors[name][1], host).value + j0 * sy * size)
                for name, host in zip(self.args, data)
            }
            if self.dry_runs:
                self.launch(pointers, self.dry_runs - 1, None, s_run.value, domain=(nx, j1 - j0, nz),
                            rows=(j0, j1))
            check(raw.sb200_event_record(pipe["begin"][c], s_run))
            self.launch(pointers, 0, None, s_run.value, domain=(nx, j1 - j0, nz), rows=(j0, j1))
            check(raw.sb200_event_record(pipe["end"][c], s_run))
            check(raw.sb200_stream_wait_event(s_down, pipe["end"][c]))
            for name, host, role in zip(self.args, data, roles):
                if role in ("out", "inout"):
                    offset = (hk * sz + (j0 + hy) * sy) * size
                    check(raw.sb200_memcpy2d_d2h(
                        _vp(host.ctypes.data + offset), sz * size, _vp(mirrors[name][1] + offset),
                        sz * size, (j1 - j0) * sy * size, nz, s_down))
        check(raw.sb200_synchronize(s_down))
        check(raw.sb200_synchronize(s_up))
        self._after_sweeps()
        total = 0.0
        elapsed = ctypes.c_double()
        for c in range(self.chunks):
            if bounds[c + 1] > bounds[c]:
                check(raw.sb200_event_elapsed(pipe["begin"][c], pipe["end"][c], ctypes.byref(elapsed)))
                total += elapsed.value
        return total

    def run_stencil(self, data):
        if self.chunks > 1:
            try:
                mirrors = self._device_fields(data)
                t0 = _time.perf_counter()
                sweep = self._run_pipelined(data, mirrors)
                wall = _time.perf_counter() - t0
            except cabi.ExecutionError as error:
                raise ExecutionError(*error.args) from error
            return {"time": sweep, "time-end-to-end": wall,
                    "bandwidth-algorithmic": self.algorithmic_bytes / sweep / 1e9}
        try:
            fresh = id(data) not in self._device
            mirrors = self._device_fields(data)
            t0 = _time.perf_counter()
            if fresh or not self.resident:
                self.upload(data, mirrors)
            self._before_sweeps(data, mirrors, None)
            t1 = _time.perf_counter()
            pointers = {
                name: self.interior_ptr(mirrors[name][1], host)
                for name, host in zip(self.args, data)
            }
            elapsed = ctypes.c_double()
            if self.timers == "gpu":
                self.launch(pointers, self.dry_runs, ctypes.byref(elapsed), None)
            else:
                if self.dry_runs:
                    self.launch(pointers, self.dry_runs - 1, None, None)
                capi.synchronize()
                start = _time.perf_counter()
                self.launch(pointers, 0, None, None)
                capi.synchronize()
                elapsed.value = _time.perf_counter() - start
            self._after_sweeps()
            t2 = _time.perf_counter()
            if self.verify or not self.resident:
                self.download(data, mirrors)
            t3 = _time.perf_counter()
        except cabi.ExecutionError as error:
            raise ExecutionError(*error.args) from error
        return {
            "time": elapsed.value,
            "time-h2d": t1 - t0,
            "time-d2h": t3 - t2,
            "bandwidth-algorithmic": self.algorithmic_bytes / elapsed.value / 1e9,
        }

    def geometry(self, domain=None):
        """(nx, ny, nz, sx, sy, sz) as the C ABI expects them (element strides)."""
        domain = self.domain if domain is None else domain
        return tuple(int(d) for d in domain) + tuple(int(s) for s in self.strides)


class SlabCopies:
    """Copies between the host fields of a GLOBAL domain and the J slabs a process keeps on several
    devices (the ``partitioned`` benchmark classes).  A slab is a dict with ``device``, ``start``
    (first global row), ``sz`` (its level pitch in elements) and ``first[name]`` (device address of
    element 0 of the field's slab); rows keep the host's pitch."""

    def _copy_rows(self, slab, host, name, first_row, rows, to_device, planes):
        """Padded rows [first_row, first_row + rows) of the slab, `planes` = (first, count) levels."""
        size = host.itemsize
        sy, sz_host = int(self.strides[1]), int(self.strides[2])
        k0, nk = planes
        h_off = ((slab["start"] + first_row) * sy + k0 * sz_host) * size
        d_off = (first_row * sy + k0 * slab["sz"]) * size
        if to_device:
            self._lib.sb200_memcpy2d_h2d(_vp(slab["first"][name] + d_off), slab["sz"] * size,
                                         _vp(host.ctypes.data + h_off), sz_host * size, rows * sy * size, nk, None)
        else:
            self._lib.sb200_memcpy2d_d2h(_vp(host.ctypes.data + h_off), sz_host * size,
                                         _vp(slab["first"][name] + d_off), slab["sz"] * size, rows * sy * size, nk, None)

    def _sync_all(self, slabs):
        for slab in slabs:
            self._lib.sb200_set_device(slab["device"])
            capi.synchronize()
