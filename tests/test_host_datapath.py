"""The host data path of the backend classes without a GPU.

`run()` of a backend class is host logic around the kernels: device mirrors with the host's
strides, role-based uploads / downloads, and -- with `chunks > 1` -- a slab pipeline whose row
arithmetic decides which rows travel when.  Here the device is emulated in host memory (device
buffers are NumPy buffers, copies are memmove) and the kernels by the C oracle applied to those
buffers through the very pointers and geometry the C ABI would receive.  If a slab were uploaded
too late, a halo row forgotten or a download misplaced, the result would differ from the oracle
applied to the host fields.  (On the GPU box the same paths run for real: tests/test_gpu_parity.py.)
"""

import ctypes

import numpy as np
import pytest

from oracle import native, stencils
from stencil_benchmarks_b200 import capi, distributed
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import (
    basic,
    horizontal_diffusion,
    vertical_advection,
)
from stencil_benchmarks_b200.tools import fields

I64 = ctypes.c_int64


def address(pointer):
    return pointer.value if isinstance(pointer, ctypes.c_void_p) else int(pointer)


class FakeDevice:
    """Counts what crosses the emulated link and stands in for the library handles."""

    def __init__(self):
        self.h2d = self.d2h = 0
        self.launches = []
        self.handles = 0
        self.raw = self

    # ---- wrapped-style entry points (return value unused) and raw ones (int status) ----------
    def sb200_set_device(self, device):
        return 0

    def _new_handle(self, reference):
        self.handles += 1
        reference._obj.value = self.handles
        return 0

    sb200_stream_create = sb200_event_create = _new_handle

    def sb200_event_record(self, event, stream):
        return 0

    def sb200_stream_wait_event(self, stream, event):
        return 0

    def sb200_synchronize(self, stream):
        return 0

    def sb200_event_elapsed(self, begin, end, reference):
        reference._obj.value = 1e-3
        return 0

    def _copy2d(self, dst, dst_pitch, src, src_pitch, width, height):
        for row in range(height):
            ctypes.memmove(address(dst) + row * dst_pitch, address(src) + row * src_pitch, width)
        return width * height

    def sb200_memcpy2d_h2d(self, dst, dst_pitch, src, src_pitch, width, height, stream):
        self.h2d += self._copy2d(dst, dst_pitch, src, src_pitch, width, height)
        return 0

    def sb200_memcpy2d_d2h(self, dst, dst_pitch, src, src_pitch, width, height, stream):
        self.d2h += self._copy2d(dst, dst_pitch, src, src_pitch, width, height)
        return 0

    # ---- the kernels: the C oracle on the emulated device memory ---------------------------------
    @staticmethod
    def _finish(time_pointer):
        if time_pointer is not None:
            time_pointer._obj.value = 1e-3
        return 0  # status of the bare handle; callers of the wrapped form ignore it

    def sb200_hdiff(self, code, inp, coeff, out, nx, ny, nz, sx, sy, sz, dry_runs, time_pointer, stream):
        assert sx == 1 and code == capi.F64
        self.launches.append((ny, address(out)))
        native.lib().oracle_hdiff_f64(inp, coeff, out, I64(nx), I64(ny), I64(nz), I64(sy), I64(sz))
        return self._finish(time_pointer)

    def sb200_basic(self, kind, code, inp, out, nx, ny, nz, sx, sy, sz, axis, mask, dry_runs, time_pointer, stream):
        assert sx == 1 and code == capi.F64
        self.launches.append((ny, address(out)))
        geometry = [I64(nx), I64(ny), I64(nz), I64(sy), I64(sz)]
        if kind == capi.BASIC_COPY:
            native.lib().oracle_copy_f64(inp, out, *geometry)
        elif kind == capi.BASIC_LAPLACIAN:
            native.lib().oracle_laplacian_f64(inp, out, *geometry, ctypes.c_int(mask))
        else:
            native.lib().oracle_average_f64(inp, out, *geometry, ctypes.c_int(axis),
                                            ctypes.c_int(int(kind == capi.BASIC_SYMMETRIC_AVG)))
        return self._finish(time_pointer)

    def sb200_vadv_components(self, code, ncomp, stage, pos, tens, tensstage, ishift, jshift, wcon, ccol, dcol,
                              nx, ny, nz, sx, sy, sz, variant, dry_runs, time_pointer, stream):
        assert sx == 1 and code == capi.F64
        self.launches.append((ny, address(tensstage[0])))
        for c in range(ncomp):
            native.lib().oracle_vadv_f64(
                ctypes.c_void_p(stage[c]), ctypes.c_void_p(pos[c]), ctypes.c_void_p(tens[c]),
                ctypes.c_void_p(tensstage[c]), wcon, ccol, dcol, I64(nx), I64(ny), I64(nz), I64(sy), I64(sz),
                ctypes.c_int(ishift[c]), ctypes.c_int(jshift[c]))
        return self._finish(time_pointer)


class FakeBuffer:
    def __init__(self, nbytes):
        self.array = np.full(nbytes + 64, 0xA5, dtype=np.uint8)  # poison: unwritten memory shows
        self.ptr = self.array.ctypes.data
        self.nbytes = nbytes


@pytest.fixture
def device(monkeypatch):
    fake = FakeDevice()

    def memcpy_h2d(dptr, host_ptr, nbytes, stream=None, sync=True):
        ctypes.memmove(dptr, host_ptr, nbytes)
        fake.h2d += nbytes

    def memcpy_d2h(host_ptr, dptr, nbytes, stream=None, sync=True):
        ctypes.memmove(host_ptr, dptr, nbytes)
        fake.d2h += nbytes

    monkeypatch.setattr(capi, "require_device", lambda: None)
    monkeypatch.setattr(capi, "DeviceBuffer", FakeBuffer)
    monkeypatch.setattr(capi, "memcpy_h2d", memcpy_h2d)
    monkeypatch.setattr(capi, "memcpy_d2h", memcpy_d2h)
    monkeypatch.setattr(capi, "synchronize", lambda stream=None: None)
    return fake


def on_fake_device(cls, fake, **kwargs):
    bench = cls(pinned=False, verify=False, seed=11, **kwargs)
    bench._lib = bench._kernels = fake
    return bench


def interior(bench):
    return tuple(slice(h, h + d) for d, h in zip(bench.domain, bench.halo))


@pytest.mark.parametrize("chunks", [1, 2, 3, 7, 29])
@pytest.mark.parametrize("halo", [(3, 3, 3), (2, 2, 0)])
def test_hdiff_run_moves_the_right_rows(device, chunks, halo):
    bench = on_fake_device(horizontal_diffusion.Fused, device, domain=(37, 29, 4), halo=halo, chunks=chunks)
    data = bench.data(0)
    inp0, coeff0, out0 = (np.array(f, copy=True) for f in data)
    expected = stencils.hdiff(inp0, coeff0, bench.halo)
    # what the mirrors are set up with is not part of a run's traffic
    mirrors = bench._device_fields(data)
    device.h2d = device.d2h = 0
    result = bench.run()
    assert result["time"] > 0
    inner = interior(bench)
    np.testing.assert_array_equal(data.out[inner], expected[inner])
    # inputs untouched, out's halo and padding keep the host's values
    assert np.array_equal(data.inp, inp0) and np.array_equal(data.coeff, coeff0)
    outside = np.ones(out0.shape, dtype=bool)
    outside[inner] = False
    assert np.array_equal(data.out[outside], out0[outside])
    # every interior row swept exactly once, in `chunks` launches (at most one per row)
    assert sum(rows for rows, _ in device.launches) == 29 and len(device.launches) == min(chunks, 29)
    # the byte counters of the run equal what transfer_bytes() declares (bench.py's e2e keys)
    assert (device.h2d, device.d2h) == bench.transfer_bytes()
    assert mirrors is bench._device_fields(data)  # mirrors are kept between runs


@pytest.mark.parametrize("chunks", [1, 4])
@pytest.mark.parametrize("case", [
    (basic.Copy, {}, lambda f, h: stencils.copy(f, h)),
    (basic.OnesidedAverage, dict(axis=1), lambda f, h: stencils.onesided_average(f, h, 1)),
    (basic.SymmetricAverage, dict(axis=2), lambda f, h: stencils.symmetric_average(f, h, 2)),
    (basic.Laplacian, dict(along_x=True, along_y=True, along_z=True),
     lambda f, h: stencils.laplacian(f, h, (True, True, True))),
], ids=["copy", "onesided_j", "symmetric_k", "laplacian_ijk"])
def test_basic_run_moves_the_right_rows(device, chunks, case):
    cls, parameters, oracle = case
    bench = on_fake_device(cls, device, domain=(21, 13, 6), halo=(1, 2, 1), chunks=chunks, **parameters)
    data = bench.data(0)
    inp0 = np.array(data.inp, copy=True)
    expected = oracle(inp0, bench.halo)
    bench._device_fields(data)
    device.h2d = device.d2h = 0
    bench.run()
    inner = interior(bench)
    np.testing.assert_array_equal(data.out[inner], expected[inner])
    assert np.array_equal(data.inp, inp0)
    assert (device.h2d, device.d2h) == bench.transfer_bytes()


def test_resident_mode_uploads_once(device):
    bench = on_fake_device(horizontal_diffusion.Fused, device, domain=(16, 10, 3), resident=True)
    data = bench.data(0)
    expected = stencils.hdiff(np.array(data.inp), np.array(data.coeff), bench.halo)
    bench.run()
    after_first = (device.h2d, device.d2h)
    bench.run()
    # the second run moves nothing in either direction (verify is off): the fields live in "HBM"
    assert (device.h2d, device.d2h) == after_first
    state = bench.empty_field()  # same padded layout as the mirror
    mirrors = bench._device_fields(data)
    ctypes.memmove(state.ctypes.data, mirrors["out"][1], fields.nbytes(state))
    inner = interior(bench)
    np.testing.assert_array_equal(state[inner], expected[inner])


@pytest.mark.parametrize("chunks", [1, 3])
@pytest.mark.parametrize("all_components", [False, True])
def test_vadv_run_moves_inputs_up_and_only_the_solution_down(device, chunks, all_components):
    bench = on_fake_device(vertical_advection.Thomas, device, domain=(18, 11, 9), halo=(1, 1, 1),
                           chunks=chunks, all_components=all_components)
    data = bench.data(0)
    before = {name: np.array(field, copy=True) for name, field in zip(bench.args, data)}
    components = "uvw" if all_components else "u"
    if all_components:
        expected = dict(zip("uvw", stencils.vadv_all(
            *[tuple(before[c + f] for f in ("stage", "pos", "tens", "tensstage")) for c in "uvw"],
            before["wcon"], bench.halo)))
    else:
        expected = {"u": stencils.vadv(before["ustage"], before["upos"], before["utens"], before["utensstage"],
                                       before["wcon"], bench.halo)}
    bench._device_fields(data)
    device.h2d = device.d2h = 0
    bench.run()
    inner = interior(bench)
    for c in components:
        np.testing.assert_array_equal(getattr(data, c + "tensstage")[inner], expected[c][inner])
    # read-only inputs and the scratch columns come back untouched: scratch never crosses the link
    for name in bench.args:
        if not name.endswith("tensstage"):
            assert np.array_equal(getattr(data, name), before[name]), name
    assert (device.h2d, device.d2h) == bench.transfer_bytes()
    moved_up = sum(1 for name in bench.args if bench.field_roles[name] in ("in", "inout"))
    assert moved_up == (13 if all_components else 5) and device.d2h < device.h2d


# ---- the J-partitioned class: N emulated devices in one process --------------------------------------
class FakePartitionDevice(FakeDevice):
    """Adds what `Partitioned` needs: memset, peer access, and `sb200_hdiff_peer`, whose halo rows
    come from the NEIGHBOURING slabs' memory (here: gathered into a scratch field for the oracle)."""

    def __init__(self):
        super().__init__()
        self.peer_pairs = set()
        self.devices_used = []

    def sb200_set_device(self, device):
        self.devices_used.append(device)
        return 0

    def sb200_memset(self, pointer, value, nbytes, stream, sync):
        ctypes.memset(address(pointer), value, nbytes)
        return 0

    def sb200_enable_peer_access(self, device, peer):
        self.peer_pairs.add((device, peer))
        return 0

    def sb200_event_destroy(self, event):
        return 0

    @staticmethod
    def _rows(pointer, nx, first_row, rows, nz, sy, sz):
        """Rows [first_row, first_row + rows) x i in [-2, nx + 2) x nz levels at `pointer` (interior origin)."""
        base = address(pointer) + 8 * (first_row * sy - 2)
        flat = np.ctypeslib.as_array((ctypes.c_double * 1).from_address(base))
        return np.lib.stride_tricks.as_strided(flat, shape=(nx + 4, rows, nz), strides=(8, 8 * sy, 8 * sz))

    def sb200_hdiff_peer(self, code, inp, coeff, out, lower, ny_lower, sz_lower, upper, ny_upper, sz_upper,
                         nx, ny, nz, sx, sy, sz, dry_runs, time_pointer, stream):
        assert sx == 1 and code == capi.F64
        self.launches.append((ny, address(out)))
        halo = (2, 2, 0)
        field = np.zeros((nx + 4, ny + 4, nz), order="F")
        field[:, 2:ny + 2, :] = self._rows(inp, nx, 0, ny, nz, sy, sz)
        # a neighbour's slab supplies the two rows beyond the edge; at the global boundary they are local
        field[:, :2, :] = (self._rows(lower, nx, ny_lower - 2, 2, nz, sy, sz_lower) if address(lower)
                           else self._rows(inp, nx, -2, 2, nz, sy, sz))
        field[:, ny + 2:, :] = (self._rows(upper, nx, 0, 2, nz, sy, sz_upper) if address(upper)
                                else self._rows(inp, nx, ny, 2, nz, sy, sz))
        weights = np.zeros_like(field)
        weights[2:nx + 2, 2:ny + 2, :] = self._rows(coeff, nx, 0, ny, nz, sy, sz)[2:nx + 2]
        result = np.zeros_like(field)
        native.hdiff(field, weights, result, halo)
        self._rows(out, nx, 0, ny, nz, sy, sz)[2:nx + 2] = result[2:nx + 2, 2:ny + 2, :]
        return self._finish(time_pointer)


@pytest.mark.parametrize("gpus,domain", [(1, (20, 9, 3)), (2, (33, 17, 4)), (3, (20, 10, 2)), (4, (16, 8, 3))])
def test_partitioned_class_scatters_sweeps_and_gathers(monkeypatch, gpus, domain):
    fake = FakePartitionDevice()
    monkeypatch.setattr(capi, "require_device", lambda: None)
    monkeypatch.setattr(capi, "device_count", lambda: 8)
    monkeypatch.setattr(capi, "DeviceBuffer", FakeBuffer)
    monkeypatch.setattr(capi, "synchronize", lambda stream=None: None)
    bench = on_fake_device(horizontal_diffusion.Partitioned, fake, domain=domain, gpus=gpus, device=2)
    data = bench.data(0)
    inp0, coeff0, out0 = (np.array(f, copy=True) for f in data)
    expected = stencils.hdiff(inp0, coeff0, bench.halo)
    result = bench.run()
    inner = interior(bench)
    # internal j-halo rows of the slabs are never uploaded (they stay 0xFF = NaN in "HBM"): equality
    # with the oracle on the GLOBAL field shows the sweeps took them from the neighbouring slabs
    np.testing.assert_array_equal(data.out[inner], expected[inner])
    assert np.array_equal(data.inp, inp0) and np.array_equal(data.coeff, coeff0)
    outside = np.ones(out0.shape, dtype=bool)
    outside[inner] = False
    assert np.array_equal(data.out[outside], out0[outside])
    assert result["gpus"] == gpus and result["time"] > 0
    assert sorted(rows for rows, _ in fake.launches) == sorted(
        count for _, count in distributed.split_rows(domain[1], gpus))
    assert set(fake.devices_used) == set(range(2, 2 + gpus))
    assert fake.peer_pairs == {(a, b) for a in range(2, 2 + gpus) for b in (a - 1, a + 1) if 2 <= b < 2 + gpus}


def test_time_loop_alternates_the_two_buffers(device, monkeypatch):
    """distributed.TimeLoop on one (emulated) device: sweep m reads X_m and writes X_(m+1), the two
    buffers alternate, the halo keeps its values -- equal to the oracle applied again and again."""

    def memcpy_d2d(dst, src, nbytes, stream, sync):
        ctypes.memmove(address(dst), address(src), nbytes)
        return 0

    def memset(pointer, value, nbytes, stream, sync):
        ctypes.memset(address(pointer), value, nbytes)
        return 0

    device.sb200_memcpy_d2d, device.sb200_memset = memcpy_d2d, memset
    monkeypatch.setattr(capi, "library", lambda: device)
    bench = on_fake_device(horizontal_diffusion.Fused, device, domain=(24, 15, 3))
    data = bench.data(0)
    data.coeff[...] *= 0.025  # a stable time step (bench.TIME_LOOP_COEFF_SCALE)
    mirrors = bench._device_fields(data)
    bench.upload(data, mirrors)
    loop = distributed.TimeLoop(bench, mirrors)
    state = np.array(data.inp, copy=True)
    inner = interior(bench)
    for sweep in range(1, 6):
        loop.step()
        state[inner] = stencils.hdiff(state, np.array(data.coeff), bench.halo)[inner]
        got = loop.download(bench.empty_field())
        np.testing.assert_array_equal(got, state, err_msg=f"after sweep {sweep}")
    assert loop.count == 5 and len({out for _, out in device.launches}) == 2
    loop.close()


def test_data_sets_cycle_with_their_own_mirrors(device):
    """`data_sets` (base.py:47-51, :152): run() cycles through independent field sets; every set has
    its own device mirrors, allocated once, and every sweep lands in the set it belongs to."""
    bench = on_fake_device(basic.Laplacian, device, domain=(12, 9, 4), halo=(1, 1, 1), data_sets=3)
    expected = [stencils.laplacian(np.array(bench.data(n).inp), bench.halo) for n in range(3)]
    for _ in range(5):
        bench.run()
    inner = interior(bench)
    for n in range(3):
        np.testing.assert_array_equal(bench.data(n).out[inner], expected[n][inner])
    assert len(bench._device) == 3 and len(device.launches) == 5
    # three distinct output mirrors were written: sets 0 and 1 twice, set 2 once
    written = [out for _, out in device.launches]
    assert len(set(written)) == 3 and written[0] == written[3] and written[1] == written[4]


def test_wall_timer_and_dry_runs(device):
    bench = on_fake_device(basic.Copy, device, domain=(10, 6, 3), halo=(0, 0, 0), timers="wall", dry_runs=2)
    result = bench.run()
    assert result["time"] > 0 and result["bandwidth"] > 0
    np.testing.assert_array_equal(bench.data(0).out, bench.data(0).inp)
    # two C calls: the warm-up call (which the library repeats dry_runs times) and the wall-timed one
    assert len(device.launches) == 2


@pytest.mark.parametrize("gpus", [1, 2, 3, 5])
@pytest.mark.parametrize("case", [
    (basic.PartitionedCopy, {}, lambda f, h: stencils.copy(f, h)),
    (basic.PartitionedOnesidedAverage, dict(axis=1), lambda f, h: stencils.onesided_average(f, h, 1)),
    (basic.PartitionedSymmetricAverage, dict(axis=1), lambda f, h: stencils.symmetric_average(f, h, 1)),
    (basic.PartitionedSymmetricAverage, dict(axis=2), lambda f, h: stencils.symmetric_average(f, h, 2)),
    (basic.PartitionedLaplacian, dict(along_x=True, along_y=True, along_z=True),
     lambda f, h: stencils.laplacian(f, h, (True, True, True))),
], ids=["copy", "onesided_j", "symmetric_j", "symmetric_k", "laplacian_ijk"])
def test_partitioned_basic_stencils(monkeypatch, gpus, case):
    """The basic stencils over 1-5 emulated devices: the j halo of a slab (rows of the neighbouring
    slabs) arrives with the scatter; gathered result equal to the oracle on the global field."""
    cls, parameters, oracle = case
    fake = FakePartitionDevice()
    monkeypatch.setattr(capi, "require_device", lambda: None)
    monkeypatch.setattr(capi, "device_count", lambda: 8)
    monkeypatch.setattr(capi, "DeviceBuffer", FakeBuffer)
    monkeypatch.setattr(capi, "synchronize", lambda stream=None: None)
    bench = on_fake_device(cls, fake, domain=(19, 11, 4), halo=(1, 1, 2), gpus=gpus, device=1, **parameters)
    data = bench.data(0)
    inp0, out0 = np.array(data.inp, copy=True), np.array(data.out, copy=True)
    expected = oracle(inp0, bench.halo)
    result = bench.run()
    inner = interior(bench)
    np.testing.assert_array_equal(data.out[inner], expected[inner])
    assert np.array_equal(data.inp, inp0)
    outside = np.ones(out0.shape, dtype=bool)
    outside[inner] = False
    assert np.array_equal(data.out[outside], out0[outside])
    assert result["gpus"] == gpus and result["time"] > 0 and "bandwidth" in result
    assert sorted(rows for rows, _ in fake.launches) == sorted(n for _, n in distributed.split_rows(11, gpus))
    assert set(fake.devices_used) == set(range(1, 1 + gpus))
    bench.run()  # buffers and events are allocated once
    assert len(fake.launches) == 2 * gpus and fake.handles == 2 * gpus


def test_partitioned_basic_parameter_errors():
    from stencil_benchmarks_b200.benchmark import ParameterError

    kwargs = dict(pinned=False, verify=False, domain=(8, 3, 2))
    with pytest.raises(ParameterError, match="at least one row"):
        basic.PartitionedCopy(gpus=4, **kwargs)
    with pytest.raises(ParameterError, match="gpus must be at least 1"):
        basic.PartitionedCopy(gpus=0, **kwargs)
    with pytest.raises(ParameterError, match="chunks / resident"):
        basic.PartitionedLaplacian(gpus=2, chunks=2, **kwargs)
    with pytest.raises(ParameterError, match="positive halo"):
        basic.PartitionedOnesidedAverage(gpus=2, axis=1, halo=(1, 0, 1), **kwargs)
