"""The N > 1 plugin path on CPU: `run()` of J-slab instances in world_size-2/3 process groups (gloo).

This is the path `bench.py` times as `e2e` at N > 1 (`distributed.attach_neighbours`): every rank
maps its neighbours' `inp` slabs, uploads its edge rows first, waits at a barrier until the
neighbours' edge rows are in place, then runs the slab pipeline whose edge sweeps read the halo rows
from the NEIGHBOURS' memory, and ends with a barrier.  Here the devices are emulated: device buffers
live in POSIX shared memory (so a neighbouring process can really map them -- the stand-in for a
CUDA IPC handle is the segment's name), copies are memmove, the kernel is the C oracle on those
buffers (tests/test_host_datapath.py).  Each rank's result is checked against the oracle on the
GLOBAL field with bench.py's own checker.  The local j-halo rows towards a neighbour are poisoned
with NaN and never uploaded by the exchange-free path: a correct result proves the neighbour's rows
were read, and were there in time.
"""

import ctypes
import os
import socket
from multiprocessing import shared_memory

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import bench
from stencil_benchmarks_b200 import capi, distributed
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion
from test_host_datapath import FakePartitionDevice, address


class SharedBuffer:
    """A "device allocation" another process can map: a named shared-memory segment."""

    live = {}

    def __init__(self, nbytes):
        self.segment = shared_memory.SharedMemory(create=True, size=nbytes + 64)
        self.view = np.frombuffer(self.segment.buf, dtype=np.uint8)
        self.view[:] = 0xA5
        self.ptr = self.view.ctypes.data
        self.nbytes = nbytes
        SharedBuffer.live[self.ptr] = self

    def release(self):
        SharedBuffer.live.pop(self.ptr, None)
        self.view = None
        self.segment.close()
        self.segment.unlink()


class SharedDevice(FakePartitionDevice):
    """The emulated device plus the three IPC entry points of include/sbench_b200.h."""

    def __init__(self):
        super().__init__()
        self.mapped = {}

    def sb200_ipc_get_handle(self, base, handle):
        name = SharedBuffer.live[address(base)].segment.name.encode()
        assert len(name) < 64
        ctypes.memmove(handle, name + b"\0", len(name) + 1)
        return 0

    def sb200_ipc_open_handle(self, handle, reference):
        name = bytes(handle).split(b"\0", 1)[0].decode()
        segment = shared_memory.SharedMemory(name=name)
        view = np.frombuffer(segment.buf, dtype=np.uint8)
        reference._obj.value = view.ctypes.data
        self.mapped[view.ctypes.data] = (segment, view)
        return 0

    def sb200_ipc_close_handle(self, pointer):
        segment, view = self.mapped.pop(address(pointer))
        del view
        segment.close()
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, domain, chunks, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    buffers = []

    def allocate(nbytes):
        buffers.append(SharedBuffer(nbytes))
        return buffers[-1]

    fake = SharedDevice()

    def memcpy_h2d(dptr, host_ptr, nbytes, stream=None, sync=True):
        ctypes.memmove(dptr, host_ptr, nbytes)

    capi.require_device = lambda: None
    capi.library = lambda: fake
    capi.DeviceBuffer = allocate
    capi.memcpy_h2d = memcpy_h2d
    capi.memcpy_d2h = lambda host_ptr, dptr, nbytes, stream=None, sync=True: ctypes.memmove(host_ptr, dptr, nbytes)
    capi.synchronize = lambda stream=None: None
    try:
        nx, ny_global, nz = domain
        start, ny = distributed.split_rows(ny_global, world)[rank]
        lower, upper = distributed.neighbours(rank, world)
        slab = horizontal_diffusion.Fused(domain=(nx, ny, nz), pinned=False, verify=False, chunks=chunks)
        slab._lib = slab._kernels = fake
        data = slab.data(0)
        # this rank's rows of ONE global row-seeded field; j-halo rows towards a neighbour are NaN
        bench.fill_hdiff_slab(slab, data, start, lower is not None, upper is not None)
        peers = distributed.attach_neighbours(slab, dist, rank, world)
        dist.barrier()
        for _ in range(2):  # the second run re-uploads while the neighbours may still be reading: ordered by the final barrier
            outcome = slab.run()
        parity = bench.edge_parity(slab, data.out, start, "peer, partitioned run()")
        # ... and every row, not only the edge blocks
        halo = tuple(int(h) for h in slab.halo)
        rows = range(start, start + ny + 2 * halo[1])
        g_inp = np.asfortranarray(bench.make_global_rows("inp", rows, nx, nz, halo))
        g_coeff = np.asfortranarray(bench.make_global_rows("coeff", rows, nx, nz, halo))
        from oracle import stencils

        expected = stencils.hdiff(g_inp, g_coeff, halo)
        inner = tuple(slice(h, h + d) for d, h in zip((nx, ny, nz), halo))
        results[rank] = dict(parity_ok=parity["ok"], max_abs_err=parity["max_abs_err"],
                             all_rows_equal=bool(np.array_equal(data.out[inner], expected[inner])),
                             time=outcome["time"], launches=len(fake.launches),
                             mapped=len(fake.mapped), expected_maps=(lower is not None) + (upper is not None))
        dist.barrier()
        peers.close()
        assert not fake.mapped  # every neighbour mapping released
    finally:
        dist.barrier()
        for buffer in buffers:
            buffer.release()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,domain,chunks", [(2, (24, 21, 3), 1), (2, (40, 26, 2), 4), (3, (16, 19, 2), 3)])
def test_partitioned_run_through_the_plugin_gloo(world, domain, chunks):
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, _free_port(), domain, chunks, results), nprocs=world, join=True)
    assert len(results) == world
    for rank in range(world):
        outcome = dict(results[rank])
        assert outcome["parity_ok"] and outcome["all_rows_equal"], (rank, outcome)
        assert outcome["mapped"] == outcome["expected_maps"] and outcome["time"] > 0
