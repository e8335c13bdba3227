"""Multi-GPU host logic on CPU: J partition arithmetic and the halo exchange protocol run with
gloo (world size 2 and 3) and NumPy pack/unpack standing in for the CUDA kernels.  The
exchange must make a J-partitioned horizontal diffusion (oracle per slab) reproduce the
global result bit for bit."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import stencils
from stencil_benchmarks_b200 import distributed


def test_split_rows():
    assert distributed.split_rows(2048, 8) == [(256 * r, 256) for r in range(8)]
    assert distributed.split_rows(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert sum(c for _, c in distributed.split_rows(1001, 7)) == 1001
    with pytest.raises(ValueError):
        distributed.split_rows(3, 4)


def test_neighbours_are_not_periodic():
    assert distributed.neighbours(0, 1) == (None, None)
    assert distributed.neighbours(0, 4) == (None, 1)
    assert distributed.neighbours(2, 4) == (1, 3)
    assert distributed.neighbours(3, 4) == (2, None)


def test_interior_and_boundary_rows():
    f = distributed.interior_and_boundary_rows
    assert f(256, 2, False, False) == ((0, 256), [])
    assert f(256, 2, True, False) == ((2, 256), [(0, 2)])
    assert f(256, 2, False, True) == ((0, 254), [(254, 256)])
    assert f(256, 2, True, True) == ((2, 254), [(0, 2), (254, 256)])
    (lo, hi), strips = f(3, 2, True, True)  # slab thinner than twice the reach
    assert lo <= hi and sorted(set(range(lo, hi)) | {j for a, b in strips for j in range(a, b)}) == [0, 1, 2]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shape, halo, width, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nx, ny_global, nz = shape
        hx, hy, hz = halo
        rng = np.random.default_rng(5)
        g_inp = rng.random((nx + 2 * hx, ny_global + 2 * hy, nz + 2 * hz))
        g_coeff = rng.random(g_inp.shape)
        expected = stencils.hdiff(g_inp, g_coeff)
        start, ny = distributed.split_rows(ny_global, world)[rank]
        # local slab with its own halo; interior + OUTER halos come from the global field, the
        # halos towards neighbours start as garbage and must be filled by the exchange
        rows = slice(start, start + ny + 2 * hy)
        inp = g_inp[:, rows, :].copy()
        coeff = g_coeff[:, rows, :].copy()
        lower, upper = distributed.neighbours(rank, world)
        if lower is not None:
            inp[:, :hy, :] = np.nan
        if upper is not None:
            inp[:, hy + ny:, :] = np.nan

        def make_buffer(nrows):
            return torch.empty((nx + 2 * hx) * nrows * (nz + 2 * hz), dtype=torch.float64)

        def pack(field, j0, nrows, buffer):
            block = field[:, hy + j0:hy + j0 + nrows, :]  # (i, rows, k) -> [k][row][i]
            buffer.copy_(torch.from_numpy(np.ascontiguousarray(block.transpose(2, 1, 0)).reshape(-1)))

        def unpack(field, j0, nrows, buffer):
            block = buffer.numpy().reshape(nz + 2 * hz, nrows, nx + 2 * hx).transpose(2, 1, 0)
            field[:, hy + j0:hy + j0 + nrows, :] = block

        exchange = distributed.HaloExchange(dist, rank, world, ny, width, make_buffer, pack, unpack)
        requests = exchange.start(inp)
        exchange.finish(inp, requests)
        # only `width` rows are exchanged; rows further out in the halo stay unspecified
        lo = hy - width if lower is not None else 0
        hi = hy + ny + width if upper is not None else ny + 2 * hy
        ok_halo = np.array_equal(inp[:, lo:hi, :], g_inp[:, start + lo:start + hi, :])
        inp = np.nan_to_num(inp)
        local = stencils.hdiff(inp, coeff)
        ok_result = np.array_equal(local[hx:-hx, hy:hy + ny, hz:nz + hz],
                                   expected[hx:-hx, hy + start:hy + start + ny, hz:nz + hz])
        results[rank] = (bool(ok_halo), bool(ok_result), exchange.bytes_per_exchange)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,width", [(2, 3), (3, 2)])
def test_halo_exchange_gloo(world, width):
    shape, halo = (12, 17, 3), (3, 3, 1)
    manager = mp.Manager()
    results = manager.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, shape, halo, width, results), nprocs=world, join=True)
    assert len(results) == world
    for rank in range(world):
        ok_halo, ok_result, nbytes = results[rank]
        assert ok_halo, f"rank {rank}: halo rows differ from the neighbour's interior"
        assert ok_result, f"rank {rank}: partitioned hdiff differs from the global result"
        faces = (rank > 0) + (rank < world - 1)
        assert nbytes == faces * width * (shape[0] + 2 * halo[0]) * (shape[2] + 2 * halo[2]) * 8


def _narrow_worker(rank, world, port, results):
    """Unequal slabs where ONE rank holds fewer rows than the halo is wide: every rank must refuse
    (the neighbours would receive rows the narrow slab does not have)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ny = 2 if rank == 1 else 6
        try:
            distributed.HaloExchange(dist, rank, world, ny, 3, lambda n: torch.empty(n), None, None)
        except ValueError as error:
            results[rank] = str(error)
        else:
            results[rank] = "accepted"
    finally:
        dist.destroy_process_group()


def test_halo_exchange_refuses_a_slab_narrower_than_the_halo():
    results = mp.Manager().dict()
    mp.spawn(_narrow_worker, args=(2, _free_port(), results), nprocs=2, join=True)
    assert "wider than the local slab" in results[1]
    assert "smallest slab (2 rows)" in results[0]
