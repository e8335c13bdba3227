"""Multi-GPU plumbing: J partition of the IJ plane and the width-h halo exchange.

New functionality with respect to the reference, which is single-GPU
(SURVEY.md §2, §8e).  One process per GPU (``torch.distributed``: NCCL on GPUs,
gloo in the CPU tests); the global domain is split along j into contiguous
slabs so rows stay i-contiguous.  Per sweep each rank sends its first / last
``width`` interior rows to the lower / upper neighbour and receives their
counterparts into its own halo rows.  The outermost halos of the global domain
keep the values the host provided (the reference has no boundary condition), so
results are bit-identical to a single-GPU run of the global domain.

STREAM, the basic copy and vertical advection need no exchange: vadv reads
wcon(i+1, j) only, and i is not partitioned.

Two exchanges exist for horizontal diffusion: ``PeerSlabs`` + ``sb200_hdiff_peer`` fuse it into
the sweep (the edge tiles read the neighbours' rows over NVLink peer memory), ``HaloExchange``
moves packed faces with send/recv.  The latter is written against three injected callables (make_buffer, pack,
unpack) so the same neighbour / row arithmetic runs with CUDA kernels + NCCL in
production (``cuda_halo_exchange``) and with NumPy + gloo in the CPU tests.
"""

from typing import Callable, List, Optional, Tuple


def split_rows(ny_global: int, world_size: int) -> List[Tuple[int, int]]:
    """(first row, row count) of every rank; the remainder goes to the first ranks."""
    if world_size < 1 or ny_global < world_size:
        raise ValueError(f"cannot split {ny_global} rows over {world_size} ranks")
    base, extra = divmod(ny_global, world_size)
    rows, start = [], 0
    for rank in range(world_size):
        count = base + (1 if rank < extra else 0)
        rows.append((start, count))
        start += count
    return rows


def neighbours(rank: int, world_size: int) -> Tuple[Optional[int], Optional[int]]:
    """(lower, upper) neighbour in j; None at the ends of the global domain (no periodicity)."""
    lower = rank - 1 if rank > 0 else None
    upper = rank + 1 if rank < world_size - 1 else None
    return lower, upper


def interior_and_boundary_rows(ny: int, reach: int, has_lower: bool, has_upper: bool):
    """Rows that can be computed before the halos arrive, and the strips that cannot.

    A stencil reaching ``reach`` rows in j (2 for horizontal diffusion) can update
    rows [lo, hi) without the neighbours' data; the strips [0, lo) and [hi, ny)
    wait for the exchange.  Returns ((lo, hi), [(start, stop), ...]).
    """
    lo = min(reach, ny) if has_lower else 0
    hi = max(ny - reach, lo) if has_upper else ny
    strips = []
    if lo > 0:
        strips.append((0, lo))
    if hi < ny:
        strips.append((hi, ny))
    return (lo, hi), strips


class HaloExchange:
    """Width-``width`` halo exchange of one field along j between neighbouring ranks.

    make_buffer(nrows) -> a buffer object accepted by ``dist.isend`` / ``dist.irecv``
    pack(field, j0, nrows, buffer)    copy rows [j0, j0+nrows) of every level into buffer
    unpack(field, j0, nrows, buffer)  the inverse
    Row indices are relative to the first interior row (negative = lower halo).
    """

    def __init__(self, dist, rank: int, world_size: int, ny: int, width: int,
                 make_buffer: Callable, pack: Callable, unpack: Callable, group=None):
        counts = [int(ny)]
        if world_size > 1:
            # the neighbours send `width` rows of THEIR slabs: every slab must hold that many.  All
            # ranks take part in the gather before anybody raises, so that all of them refuse together
            counts = [None] * world_size
            dist.all_gather_object(counts, int(ny), group=group)
        if width > ny:
            raise ValueError("halo wider than the local slab")
        if min(counts) < width:
            raise ValueError(f"halo of width {width} wider than the smallest slab ({min(counts)} rows)")
        self.dist = dist
        self.rank, self.world_size = rank, world_size
        self.ny, self.width = ny, width
        self.pack, self.unpack = pack, unpack
        self.group = group
        self.lower, self.upper = neighbours(rank, world_size)
        self.send_lower = make_buffer(width) if self.lower is not None else None
        self.recv_lower = make_buffer(width) if self.lower is not None else None
        self.send_upper = make_buffer(width) if self.upper is not None else None
        self.recv_upper = make_buffer(width) if self.upper is not None else None

    def start(self, field):
        """Pack the edge rows and post the sends / receives; returns the pending requests."""
        ops = []
        P2POp = self.dist.P2POp
        if self.lower is not None:
            self.pack(field, 0, self.width, self.send_lower)
            ops.append(P2POp(self.dist.isend, self.send_lower, self.lower, self.group))
            ops.append(P2POp(self.dist.irecv, self.recv_lower, self.lower, self.group))
        if self.upper is not None:
            self.pack(field, self.ny - self.width, self.width, self.send_upper)
            ops.append(P2POp(self.dist.isend, self.send_upper, self.upper, self.group))
            ops.append(P2POp(self.dist.irecv, self.recv_upper, self.upper, self.group))
        return self.dist.batch_isend_irecv(ops) if ops else []

    def finish(self, field, requests):
        """Wait for the messages and write them into the halo rows."""
        for request in requests:
            request.wait()
        if self.lower is not None:
            self.unpack(field, -self.width, self.width, self.recv_lower)
        if self.upper is not None:
            self.unpack(field, self.ny, self.width, self.recv_upper)

    @property
    def bytes_per_exchange(self):
        """Bytes this rank sends per sweep."""
        total = 0
        for buffer in (self.send_lower, self.send_upper):
            if buffer is not None:
                total += buffer.numel() * buffer.element_size()
        return total


def attach_neighbours(bench, dist, rank, world_size, group=None):
    """Make a horizontal-diffusion instance the sweep of ONE J slab of a domain partitioned over
    ``world_size`` ranks: map the neighbours' ``inp`` slabs (PeerSlabs) and hand them to the
    instance, whose ``run()`` then reads its halo rows from the neighbours (one kernel per sweep)
    and orders its copies against theirs.  Returns the PeerSlabs (close() it at the end)."""
    data = bench.data(0)
    mirrors = bench._device_fields(data)
    interior = bench.interior_ptr(mirrors["inp"][1], data.inp).value
    peers = PeerSlabs(dist, rank, world_size, mirrors["inp"][0].ptr, interior, int(bench.domain[1]),
                      int(bench.strides[2]), group=group)
    bench.peers = peers
    return peers


def global_rows(field_seed, rows, nx, nz, halo=(3, 3, 3), dtype="float64"):
    """Rows ``rows`` (indices into the padded global array) of a synthetic global field
    U[0,1) whose j-th row depends on (field_seed, j) only: every rank can build its own slab --
    and the true halo rows around it -- without ever holding the global field.  Shape
    (nx + 2 hx, len(rows), nz + 2 hz)."""
    import numpy as np

    width, levels = nx + 2 * halo[0], nz + 2 * halo[2]
    out = np.empty((width, len(rows), levels), dtype=dtype)
    for n, j in enumerate(rows):
        out[:, n, :] = np.random.default_rng([field_seed, j]).random((levels, width)).T
    return out


def cuda_halo_exchange(rank, world_size, dtype, nx, ny, nz, hx, sy, sz, width, group=None):
    """HaloExchange whose pack / unpack are the sb200 CUDA kernels and whose buffers are
    torch CUDA tensors (NCCL send/recv over NVLink).  ``field`` is the device address of
    the first interior element (an int)."""
    import ctypes

    import torch
    import torch.distributed as dist

    from . import capi

    lib = capi.library()
    code = capi.dtype_code(dtype)
    torch_dtype = {"float64": torch.float64, "float32": torch.float32}[str(dtype)]
    row_len = nx + 2 * hx

    def make_buffer(nrows):
        return torch.empty(row_len * nrows * nz, dtype=torch_dtype, device="cuda")

    def stream():
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def checked(status, what):
        # raw calls (no stdout/stderr capture on the hot path): a failed launch must not go unnoticed
        if status != 0:
            raise RuntimeError(f"{what} failed (status {status}); see stderr")

    def pack(field, j0, nrows, buffer):
        checked(lib.raw.sb200_pack_rows(code, ctypes.c_void_p(field), ctypes.c_void_p(buffer.data_ptr()),
                                        nx, nz, hx, sy, sz, j0, nrows, stream()), "sb200_pack_rows")

    def unpack(field, j0, nrows, buffer):
        checked(lib.raw.sb200_unpack_rows(code, ctypes.c_void_p(field), ctypes.c_void_p(buffer.data_ptr()),
                                          nx, nz, hx, sy, sz, j0, nrows, stream()), "sb200_unpack_rows")

    return HaloExchange(dist, rank, world_size, ny, width, make_buffer, pack, unpack, group)


class PeerSlabs:
    """Neighbour slabs mapped into this process (CUDA IPC) for the fused halo exchange.

    Every rank publishes the IPC handle of the device allocation that holds its ``inp`` field,
    the offset of the first interior element inside it, its row count and its k stride; each rank opens the
    handles of its lower / upper neighbour.  ``lower`` / ``upper`` are then device addresses of
    the neighbours' first interior elements (or None), valid for TMA and ordinary loads over
    NVLink.  ``sb200_hdiff_peer`` reads the halo rows through them inside the sweep itself.
    """

    def __init__(self, dist, rank, world_size, base_ptr, interior_ptr, ny, sz, group=None):
        import ctypes

        from . import capi

        self._lib = capi.library()
        self._dist, self._group = dist, group
        handle = (ctypes.c_ubyte * 64)()
        self._lib.sb200_ipc_get_handle(ctypes.c_void_p(base_ptr), handle)
        mine = (bytes(handle), int(interior_ptr - base_ptr), int(ny), int(sz))
        everyone = [None] * world_size
        dist.all_gather_object(everyone, mine, group=group)
        self._opened = []
        self.lower = self.upper = None
        self.ny_lower = self.ny_upper = 0
        self.sz_lower = self.sz_upper = 0
        lower, upper = neighbours(rank, world_size)
        for who, neighbour in (("lower", lower), ("upper", upper)):
            if neighbour is None:
                continue
            raw, offset, rows, kstride = everyone[neighbour]
            mapped = ctypes.c_void_p()
            self._lib.sb200_ipc_open_handle((ctypes.c_ubyte * 64).from_buffer_copy(raw),
                                            ctypes.byref(mapped))
            self._opened.append(mapped)
            setattr(self, who, mapped.value + offset)
            setattr(self, "ny_" + who, rows)
            setattr(self, "sz_" + who, kstride)

    def barrier(self):
        """All ranks of the partition have reached this point (host side).  Orders the uploads of
        one rank against the sweeps of its neighbours: the fused exchange reads the neighbours'
        ``inp`` rows straight from their HBM, so whenever ``inp`` changes between sweeps (a new
        upload, a time loop that swaps fields) every rank has to pass a barrier -- or use the
        in-kernel step flags of the iterated mode -- before the next sweep starts."""
        self._dist.barrier(group=self._group)

    def close(self):
        for mapped in self._opened:
            self._lib.sb200_ipc_close_handle(mapped)
        self._opened = []


class TimeLoop:
    """A time loop of horizontal diffusion on one J slab: sweep m reads field X_m and writes
    X_(m+1), with X_0 = the instance's ``inp`` mirror, X_1 = its ``out`` mirror, X_2 = X_0 again...
    (``coeff`` stays).  Halo values never change: both buffers start as copies of ``inp``, sweeps
    write interior points only, so the global boundary acts as a fixed-value condition -- exactly
    what applying the oracle again and again to ``out[interior]`` computes.

    With neighbours (``world_size`` > 1) the sweeps are ``sb200_hdiff_step`` launches: halo rows are
    read from the neighbours' buffers over NVLink and the sweeps of neighbouring GPUs order
    themselves through counters in peer memory (include/sbench_b200.h).  Nothing but kernel
    launches happens per step: no host synchronisation, no barrier, no copy.

    New with respect to the reference, which sweeps once per ``run()`` on one GPU.
    """

    #: own allocation for the 2 x nz counters: large enough that the CUDA allocator does not carve it
    #: out of a block shared with other buffers (an IPC handle always maps a whole block)
    FLAG_BYTES = 2 << 20

    def __init__(self, bench, mirrors, dist=None, rank=0, world_size=1, group=None):
        import ctypes

        from . import capi
        from .tools import fields

        self._ctypes = ctypes
        self._lib = capi.library()
        self._capi = capi
        self.bench = bench
        self.code = capi.dtype_code(bench.dtype)
        self.geometry = bench.geometry()
        data = bench.data(0)
        self._nbytes = fields.nbytes(data.inp)
        self.first = [mirrors["inp"][1], mirrors["out"][1]]
        interior = sum(int(s) * int(h) for s, h in zip(data.inp.strides, bench.halo))
        self.interior = [first + interior for first in self.first]
        self.coeff = bench.interior_ptr(mirrors["coeff"][1], data.coeff).value
        self._lib.sb200_memcpy_d2d(ctypes.c_void_p(self.first[1]), ctypes.c_void_p(self.first[0]),
                                   self._nbytes, None, 1)
        self.count = 0
        self.flags = capi.DeviceBuffer(self.FLAG_BYTES)
        self._lib.sb200_memset(ctypes.c_void_p(self.flags.ptr), 0, self.FLAG_BYTES, None, 1)
        self.peers = [None, None]
        self.peer_flags = None
        self._dist, self._group = dist, group
        if world_size > 1:
            ny, sz = int(bench.domain[1]), int(bench.strides[2])
            bases = [mirrors["inp"][0].ptr, mirrors["out"][0].ptr]
            self.peers = [PeerSlabs(dist, rank, world_size, bases[n], self.interior[n], ny, sz, group=group)
                          for n in range(2)]
            self.peer_flags = PeerSlabs(dist, rank, world_size, self.flags.ptr, self.flags.ptr, 0, 0, group=group)
            dist.barrier(group=group)  # every rank's buffers and counters are in place

    def step(self, stream=None):
        """Enqueue sweep number ``count`` on ``stream``."""
        c = self._ctypes
        vp = c.c_void_p
        m = self.count
        src, dst = m % 2, (m + 1) % 2
        peers = self.peers[src]
        if peers is None:
            status = self._lib.raw.sb200_hdiff(self.code, vp(self.interior[src]), vp(self.coeff),
                                               vp(self.interior[dst]), *self.geometry, 0, None, vp(stream))
        else:
            flags = self.peer_flags
            status = self._lib.raw.sb200_hdiff_step(
                self.code, vp(self.interior[src]), vp(self.coeff), vp(self.interior[dst]),
                vp(peers.lower), peers.ny_lower, peers.sz_lower, vp(peers.upper), peers.ny_upper, peers.sz_upper,
                vp(self.flags.ptr),
                # the lower neighbour's arrived[nz:] (pushed by its upper neighbour: this slab),
                # the upper neighbour's arrived[:nz]
                vp(flags.lower + 4 * int(self.geometry[2]) if flags.lower is not None else None),
                vp(flags.upper if flags.upper is not None else None),
                m, *self.geometry, None, vp(stream))
        if status != 0:
            raise RuntimeError(f"sweep {m} of the time loop failed (status {status}); see stderr")
        self.count += 1

    def download(self, host):
        """The current state X_count (whole padded field) into the host array ``host``."""
        self._capi.memcpy_d2h(host.ctypes.data, self.first[self.count % 2], self._nbytes)
        return host

    def close(self):
        if self._dist is not None and self.peer_flags is not None:
            self._capi.synchronize()
            self._dist.barrier(group=self._group)
        for peers in self.peers + [self.peer_flags]:
            if peers is not None:
                peers.close()
        self.peers = [None, None]
        self.peer_flags = None
