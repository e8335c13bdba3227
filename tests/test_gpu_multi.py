"""Two-GPU parity of the J-partitioned horizontal diffusion: NCCL halo exchange (CUDA pack /
unpack kernels) + per-slab sweeps must reproduce the oracle on the global domain.
Skipped on boxes with fewer than two GPUs."""

import ctypes
import os
import socket

import numpy as np
import pytest

from oracle import stencils
from stencil_benchmarks_b200 import capi

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, domain, mode, results):
    import torch
    import torch.distributed as dist

    from stencil_benchmarks_b200 import distributed
    from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        nx, ny_global, nz = domain
        halo = (3, 3, 3)
        rng = np.random.default_rng(21)
        shape = tuple(d + 2 * h for d, h in zip(domain, halo))
        g_inp, g_coeff = rng.random(shape), rng.random(shape)
        start, ny = distributed.split_rows(ny_global, world)[rank]
        bench = horizontal_diffusion.Fused(domain=(nx, ny, nz), halo=halo, verify=False, device=rank)
        data = bench.data()
        rows = slice(start, start + ny + 6)
        data.inp[...] = g_inp[:, rows, :]
        data.coeff[...] = g_coeff[:, rows, :]
        lower, upper = distributed.neighbours(rank, world)
        if lower is not None:
            data.inp[:, :3, :] = -7.0  # must be replaced by the neighbour's rows
        if upper is not None:
            data.inp[:, 3 + ny:, :] = -7.0
        mirrors = bench._device_fields(data)
        bench.upload(data, mirrors)
        ptr = {n: bench.interior_ptr(mirrors[n][1], h).value for n, h in zip(bench.args, data)}
        _, _, _, _, sy, sz = bench.geometry()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        if mode == "nccl":
            exchange = distributed.cuda_halo_exchange(rank, world, "float64", nx, ny, nz, 3, sy, sz, width=3)
            requests = exchange.start(ptr["inp"])
            exchange.finish(ptr["inp"], requests)
            capi.library().sb200_hdiff(capi.F64, ptr["inp"], ptr["coeff"], ptr["out"], nx, ny, nz, 1, sy, sz,
                                       0, None, stream)
            torch.cuda.synchronize()
        else:
            # fused exchange: the local halo rows keep their poison value, the sweep must read the
            # neighbour's rows through peer memory instead
            dist.barrier()
            torch.cuda.synchronize()
            peers = distributed.PeerSlabs(dist, rank, world, mirrors["inp"][0].ptr, ptr["inp"], ny, sz)
            assert (peers.lower is not None) == (lower is not None)
            assert (peers.upper is not None) == (upper is not None)
            capi.library().sb200_hdiff_peer(capi.F64, ptr["inp"], ptr["coeff"], ptr["out"], peers.lower,
                                            peers.ny_lower, peers.sz_lower, peers.upper, peers.ny_upper,
                                            peers.sz_upper, nx, ny, nz, 1, sy, sz, 0, None, stream)
            torch.cuda.synchronize()
            dist.barrier()
            peers.close()
        bench.download(data, mirrors)
        expected = stencils.hdiff(g_inp, g_coeff)
        ok = np.allclose(data.out[3:-3, 3:3 + ny, 3:-3], expected[3:-3, 3 + start:3 + start + ny, 3:-3],
                         rtol=1e-13, atol=1e-14)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "peer"])
@pytest.mark.parametrize("domain", [(300, 64, 5), (512, 259, 3)])
def test_partitioned_hdiff_matches_global_oracle(mode, domain):
    import torch
    import torch.multiprocessing as mp

    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    results = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), domain, mode, results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}


def _global_rows(field_seed, rows, nx, nz):
    """Rows [rows) of a global (nx+6, ny+6, nz+6) field whose j-th row depends on (seed, j) only,
    so every rank can build its slab -- and the true halo rows around it -- without the whole field."""
    out = np.empty((nx + 6, len(rows), nz + 6))
    for n, j in enumerate(rows):
        out[:, n, :] = np.random.default_rng([field_seed, j]).random((nz + 6, nx + 6)).T
    return out


def _full_size_worker(rank, world, port, domain, mode, results):
    import torch
    import torch.distributed as dist

    from oracle import native
    from stencil_benchmarks_b200 import distributed
    from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 2) // world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        nx, ny_global, nz = domain
        start, ny = distributed.split_rows(ny_global, world)[rank]
        bench = horizontal_diffusion.Fused(domain=(nx, ny, nz), halo=(3, 3, 3), verify=False, device=rank)
        data = bench.data()
        rows = range(start, start + ny + 6)
        true_inp = _global_rows(1, rows, nx, nz)
        data.inp[...] = true_inp
        data.coeff[...] = _global_rows(2, rows, nx, nz)
        lower, upper = distributed.neighbours(rank, world)
        if lower is not None:
            data.inp[:, :3, :] = -7.0  # the sweep has to fetch these rows from the neighbour
        if upper is not None:
            data.inp[:, 3 + ny:, :] = -7.0
        mirrors = bench._device_fields(data)
        bench.upload(data, mirrors)
        ptr = {n: bench.interior_ptr(mirrors[n][1], h).value for n, h in zip(bench.args, data)}
        _, _, _, _, sy, sz = bench.geometry()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        dist.barrier()
        torch.cuda.synchronize()
        if mode == "nccl":
            exchange = distributed.cuda_halo_exchange(rank, world, "float64", nx, ny, nz, 3, sy, sz, width=3)
            exchange.finish(ptr["inp"], exchange.start(ptr["inp"]))
            capi.library().sb200_hdiff(capi.F64, ptr["inp"], ptr["coeff"], ptr["out"], nx, ny, nz, 1, sy, sz,
                                       0, None, stream)
            peers = None
        else:
            peers = distributed.PeerSlabs(dist, rank, world, mirrors["inp"][0].ptr, ptr["inp"], ny, sz)
            capi.library().sb200_hdiff_peer(capi.F64, ptr["inp"], ptr["coeff"], ptr["out"], peers.lower,
                                            peers.ny_lower, peers.sz_lower, peers.upper, peers.ny_upper,
                                            peers.sz_upper, nx, ny, nz, 1, sy, sz, 0, None, stream)
        torch.cuda.synchronize()
        dist.barrier()
        if peers is not None:
            peers.close()
        bench.download(data, mirrors)
        # the oracle sees the rows of the GLOBAL field; stencils are local, so the slab with its
        # true halo rows is all it needs
        data.inp[...] = true_inp
        expected = bench.empty_field()  # same padded strides as the other fields
        expected[...] = 0.0
        native.hdiff(data.inp, data.coeff, expected, (3, 3, 3))
        inner = bench.inner_slice()
        results[rank] = bool(np.allclose(data.out[inner], expected[inner], rtol=1e-13, atol=1e-14))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_strong_scaled_full_size_hdiff(mode):
    """BASELINE's 2048x2048x80 float64 domain split over every GPU of the box."""
    import torch.multiprocessing as mp

    world = min(capi.device_count(), 8)
    if world < 2:
        pytest.skip("needs two GPUs")
    results = mp.Manager().dict()
    mp.spawn(_full_size_worker, args=(world, _free_port(), (2048, 2048, 80), mode, results),
             nprocs=world, join=True)
    assert dict(results) == {rank: True for rank in range(world)}


def _basic_worker(rank, world, port, case, results):
    """J-partitioned basic stencils with a j reach (same machinery as hdiff, width 1)."""
    import torch
    import torch.distributed as dist

    from stencil_benchmarks_b200 import distributed
    from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import basic

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        nx, ny_global, nz = domain = (260, 67, 6)
        halo = (1, 1, 1)
        shape = tuple(d + 2 * h for d, h in zip(domain, halo))
        g_inp = np.random.default_rng(5).random(shape)
        start, ny = distributed.split_rows(ny_global, world)[rank]
        if case == "laplacian":
            bench = basic.Laplacian(domain=(nx, ny, nz), halo=halo, verify=False, device=rank)
            expected = stencils.laplacian(g_inp, halo, along=(True, True, False))
        elif case == "symmetric":
            bench = basic.SymmetricAverage(domain=(nx, ny, nz), halo=halo, axis=1, verify=False, device=rank)
            expected = stencils.symmetric_average(g_inp, halo, axis=1)
        else:
            bench = basic.OnesidedAverage(domain=(nx, ny, nz), halo=halo, axis=1, verify=False, device=rank)
            expected = stencils.onesided_average(g_inp, halo, axis=1)
        data = bench.data()
        data.inp[...] = g_inp[:, start:start + ny + 2, :]
        lower, upper = distributed.neighbours(rank, world)
        if lower is not None:
            data.inp[:, :1, :] = -7.0  # to be replaced by the neighbour's last interior row
        if upper is not None:
            data.inp[:, 1 + ny:, :] = -7.0
        mirrors = bench._device_fields(data)
        bench.upload(data, mirrors)
        pointers = {n: bench.interior_ptr(mirrors[n][1], h) for n, h in zip(bench.args, data)}
        _, _, _, _, sy, sz = bench.geometry()
        exchange = distributed.cuda_halo_exchange(rank, world, "float64", nx, ny, nz, halo[0], sy, sz, width=1)
        exchange.finish(pointers["inp"].value, exchange.start(pointers["inp"].value))
        bench.launch(pointers, 0, None, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        bench.download(data, mirrors)
        ok = np.allclose(data.out[1:-1, 1:1 + ny, 1:-1], expected[1:-1, 1 + start:1 + start + ny, 1:-1],
                         rtol=1e-13, atol=1e-14)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["laplacian", "symmetric", "onesided"])
def test_partitioned_basic_matches_global_oracle(case):
    import torch.multiprocessing as mp

    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    results = mp.Manager().dict()
    mp.spawn(_basic_worker, args=(2, _free_port(), case, results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}


def _time_loop_worker(rank, world, port, domain, steps, results, return_state=False):
    """K sweeps of the time loop (inp/out swapped every step, neighbours ordered by the in-kernel
    step flags) against the oracle applied K times to the global field."""
    import torch
    import torch.distributed as dist

    from oracle import native
    from stencil_benchmarks_b200 import distributed
    from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        nx, ny_global, nz = domain
        halo = (3, 3, 3)
        rng = np.random.default_rng(33)
        shape = tuple(d + 2 * h for d, h in zip(domain, halo))
        # a stable time step: coeff <= 1/32 (with U[0,1) the scheme amplifies rounding ~4x per sweep)
        g_inp, g_coeff = np.asfortranarray(rng.random(shape)), np.asfortranarray(rng.random(shape) * 0.025)
        start, ny = distributed.split_rows(ny_global, world)[rank]
        bench = horizontal_diffusion.Fused(domain=(nx, ny, nz), halo=halo, verify=False, device=rank)
        data = bench.data()
        rows = slice(start, start + ny + 6)
        data.inp[...] = g_inp[:, rows, :]
        data.coeff[...] = g_coeff[:, rows, :]
        lower, upper = distributed.neighbours(rank, world)
        if lower is not None:
            data.inp[:, :3, :] = np.nan  # never read: the neighbour's rows come over NVLink
        if upper is not None:
            data.inp[:, 3 + ny:, :] = np.nan
        mirrors = bench._device_fields(data)
        bench.upload(data, mirrors)
        loop = distributed.TimeLoop(bench, mirrors, dist, rank, world)
        stream = torch.cuda.current_stream().cuda_stream
        # no host synchronisation between the sweeps: rank 1 even starts late
        if rank == 1:
            import time
            time.sleep(0.05)
        for _ in range(steps):
            loop.step(stream)
        torch.cuda.synchronize()
        state = loop.download(bench.empty_field())
        loop.close()
        if return_state:
            results[rank] = (start, ny, np.ascontiguousarray(state[3:-3, 3:3 + ny, 3:-3]))
            return
        x, y = g_inp.copy(order="F"), g_inp.copy(order="F")
        for _ in range(steps):
            native.hdiff(x, g_coeff, y, halo)
            x, y = y, x
        ok = np.allclose(state[3:-3, 3:3 + ny, 3:-3], x[3:-3, 3 + start:3 + start + ny, 3:-3],
                         rtol=1e-12, atol=1e-14)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("domain,steps", [((300, 64, 5), 7), ((512, 259, 3), 12), ((1100, 40, 9), 25),
                                          # slabs of 16 n + 1 rows: the last march segment is a single
                                          # row, so TWO segments per tile exchange rows with the upper
                                          # neighbour
                                          ((520, 66, 4), 20), ((300, 130, 2), 16)])
def test_time_loop_orders_neighbouring_gpus(domain, steps):
    import torch.multiprocessing as mp

    world = min(capi.device_count(), 4)
    if world < 2:
        pytest.skip("needs two GPUs")
    if domain[1] // world < 8:
        world = 2
    results = mp.Manager().dict()
    mp.spawn(_time_loop_worker, args=(world, _free_port(), domain, steps, results), nprocs=world, join=True)
    assert dict(results) == {rank: True for rank in range(world)}


def test_long_time_loop_is_bitwise_the_single_gpu_loop():
    """300 sweeps on every GPU of the box against the same loop on ONE GPU: the arithmetic is the
    same kernel code, so any difference -- a halo row read too early, an edge row overwritten too
    early, once in 300 sweeps -- shows up as a bit difference."""
    import torch.multiprocessing as mp

    from stencil_benchmarks_b200 import distributed
    from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion

    world = min(capi.device_count(), 8)
    if world < 2:
        pytest.skip("needs two GPUs")
    domain, steps = (1100, 37 * world, 6), 300
    results = mp.Manager().dict()
    mp.spawn(_time_loop_worker, args=(world, _free_port(), domain, steps, results, True), nprocs=world, join=True)
    # the same global field (same generator as the workers), swept on one GPU
    rng = np.random.default_rng(33)
    shape = tuple(d + 6 for d in domain)
    g_inp, g_coeff = rng.random(shape), rng.random(shape) * 0.025
    bench = horizontal_diffusion.Fused(domain=domain, halo=(3, 3, 3), verify=False)
    data = bench.data()
    data.inp[...] = g_inp
    data.coeff[...] = g_coeff
    mirrors = bench._device_fields(data)
    bench.upload(data, mirrors)
    loop = distributed.TimeLoop(bench, mirrors)
    for _ in range(steps):
        loop.step()
    capi.synchronize()
    single = loop.download(bench.empty_field())[3:-3, 3:-3, 3:-3]
    assert np.isfinite(single).all()
    for rank in range(world):
        start, ny, state = results[rank]
        assert np.array_equal(state, single[:, start:start + ny, :]), f"rank {rank} differs from the single-GPU loop"


def _partitioned_run_worker(rank, world, port, domain, chunks, results):
    """run() of an instance that sweeps one J slab (attach_neighbours): uploads, ordering against
    the neighbours, fused halo reads and downloads through the plugin call itself."""
    import torch
    import torch.distributed as dist

    from stencil_benchmarks_b200 import distributed
    from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        nx, ny_global, nz = domain
        halo = (3, 3, 3)
        rng = np.random.default_rng(8)
        shape = tuple(d + 2 * h for d, h in zip(domain, halo))
        g_inp, g_coeff = rng.random(shape), rng.random(shape)
        start, ny = distributed.split_rows(ny_global, world)[rank]
        bench = horizontal_diffusion.Fused(domain=(nx, ny, nz), halo=halo, verify=False, device=rank,
                                           chunks=chunks)
        data = bench.data()
        rows = slice(start, start + ny + 6)
        data.inp[...] = g_inp[:, rows, :]
        data.coeff[...] = g_coeff[:, rows, :]
        lower, upper = distributed.neighbours(rank, world)
        if lower is not None:
            data.inp[:, :3, :] = np.nan
        if upper is not None:
            data.inp[:, 3 + ny:, :] = np.nan
        peers = distributed.attach_neighbours(bench, dist, rank, world)
        for _ in range(2):  # the second run re-uploads while the neighbours may still be sweeping
            data.out[...] = 0.0
            result = bench.run()
        assert result["time"] > 0
        peers.close()
        expected = stencils.hdiff(g_inp, g_coeff)
        ok = np.allclose(data.out[3:-3, 3:3 + ny, 3:-3], expected[3:-3, 3 + start:3 + start + ny, 3:-3],
                         rtol=1e-13, atol=1e-14)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chunks", [1, 3])
def test_partitioned_run_through_the_plugin(chunks):
    import torch.multiprocessing as mp

    if capi.device_count() < 2:
        pytest.skip("needs two GPUs")
    results = mp.Manager().dict()
    mp.spawn(_partitioned_run_worker, args=(2, _free_port(), (520, 96, 4), chunks, results), nprocs=2, join=True)
    assert dict(results) == {0: True, 1: True}


@pytest.mark.parametrize("dtype,domain,gpus", [("float64", (300, 64, 5), 2), ("float32", (1030, 131, 3), 2),
                                               ("float64", (520, 259, 4), 0)])
def test_partitioned_plugin_class(dtype, domain, gpus):
    """`stencils b200 horizontal-diffusion partitioned`: one process drives the GPUs, halo rows come
    from the neighbouring device's slab (peer access), the gathered field matches the oracle."""
    from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion

    available = capi.device_count()
    if available < 2:
        pytest.skip("needs two GPUs")
    gpus = gpus or min(available, 8)  # 0: every GPU of the box
    bench = horizontal_diffusion.Partitioned(domain=domain, dtype=dtype, gpus=gpus, verify=False, seed=11,
                                             dry_runs=1)
    data = bench.data()
    before = [np.array(f, copy=True) for f in data]
    result = bench.run()
    assert result["gpus"] == gpus and result["time"] > 0
    expected = stencils.hdiff(before[0], before[1], halo=bench.halo)
    inner = bench.inner_slice()
    np.testing.assert_allclose(data.out[inner], expected[inner], **stencils.tolerances(dtype))
    assert np.array_equal(data.inp, before[0]) and np.array_equal(data.coeff, before[1])
    # halo of out untouched (only interior rows are gathered)
    mask = np.ones(data.out.shape, dtype=bool)
    mask[inner] = False
    assert np.array_equal(data.out[mask], before[2][mask])
