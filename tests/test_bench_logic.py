"""CPU checks of the checkers bench.py relies on: the row-seeded global field and the windowed
oracle of the time loop."""

import numpy as np
import pytest

import bench
from oracle import native
from stencil_benchmarks_b200 import distributed


def test_global_rows_are_consistent_across_ranks():
    """A slab's rows, its true halo rows and its neighbours' rows are the same numbers whoever
    builds them (bench.py: every rank regenerates what it needs of ONE global field)."""
    nx, nz, halo = 20, 3, (3, 3, 3)
    whole = distributed.global_rows(1, range(0, 46), nx, nz, halo)
    for start, count in distributed.split_rows(40, 3):
        slab = distributed.global_rows(1, range(start, start + count + 6), nx, nz, halo)
        assert np.array_equal(slab, whole[:, start:start + count + 6, :])
    other = distributed.global_rows(2, range(0, 46), nx, nz, halo)
    assert not np.array_equal(whole, other)
    assert whole.shape == (nx + 6, 46, nz + 6) and 0 <= whole.min() and whole.max() < 1
    scaled = bench.make_global_rows("coeff", range(5, 9), nx, nz, halo, coeff_scale=0.025)
    assert np.array_equal(scaled, distributed.global_rows(bench.FIELD_SEEDS["coeff"], range(5, 9), nx, nz, halo) * 0.025)


@pytest.mark.parametrize("rows", [range(3, 11), range(30, 38), range(51, 59)])
def test_windowed_time_loop_oracle_equals_the_global_one(rows):
    """time_loop_expected iterates the oracle on a window around the rows it is asked for; the
    window's artificial edges must never reach them -- also where the window is clipped by the
    global boundary (first and last rows)."""
    nx, ny, nz, halo, steps = 24, 56, 2, (3, 3, 3), 5

    def make_rows(name, which):
        return bench.make_global_rows(name, which, nx, nz, halo, bench.TIME_LOOP_COEFF_SCALE)

    padded = ny + 6
    x = np.asfortranarray(make_rows("inp", range(padded)))
    coeff = np.asfortranarray(make_rows("coeff", range(padded)))
    y = x.copy(order="F")
    for _ in range(steps):
        native.hdiff(x, coeff, y, halo)
        x, y = y, x
    got = bench.time_loop_expected(make_rows, halo, padded, rows, steps)
    assert np.array_equal(got, x[:, rows[0]:rows[-1] + 1, :])
    # the window really is smaller than the domain for rows in the middle
    if rows[0] == 30:
        assert 2 * (2 * steps + halo[1]) + len(rows) < padded


def test_edge_parity_detects_a_wrong_halo_row():
    """edge_parity compares against the oracle on the TRUE global rows: an `out` computed from a
    slab whose halo rows were not exchanged fails it."""
    nx, ny, nz, halo = 40, 24, 2, (3, 3, 3)

    class Slab:
        domain = (nx, ny, nz)
    Slab.halo = halo
    start = 24  # second slab of a global domain
    rows = range(start, start + ny + 6)
    fields = {name: np.asfortranarray(bench.make_global_rows(name, rows, nx, nz, halo)) for name in ("inp", "coeff")}
    out = np.zeros_like(fields["inp"])
    native.hdiff(fields["inp"], fields["coeff"], out, halo)
    assert bench.edge_parity(Slab, out, start, "test")["ok"]
    stale = fields["inp"].copy(order="F")
    stale[:, :3, :] = 0.5  # halo rows that never arrived
    native.hdiff(stale, fields["coeff"], out, halo)
    result = bench.edge_parity(Slab, out, start, "test")
    assert not result["ok"] and result["max_abs_err"] > 1e-6


def test_reference_arm_falls_back_to_the_oracle_port(monkeypatch, capsys):
    """Without oracle/_ref (a tree built where the reference was absent) `--impl reference` still
    prints a line: the C restatement on the host cores, kind "port"."""
    import argparse
    import json

    from oracle import ref_cpu

    monkeypatch.setattr(ref_cpu, "available", lambda: False)
    monkeypatch.setitem(bench.WORKLOADS, "hdiff", dict(bench.WORKLOADS["hdiff"], domain=(64, 48, 5)))
    monkeypatch.setitem(bench.WORKLOADS, "vadv", dict(bench.WORKLOADS["vadv"], domain=(32, 16, 12)))
    for workload, gpus in (("hdiff", 2), ("vadv", 1)):
        args = argparse.Namespace(workload=workload, steps=2, warmup=1, gpus=gpus, no_extras=False)
        assert bench.run_reference(args) == 0
        line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
        # both arms of the driver's comparison carry the same config at every N
        assert line["config"] == bench.workload_config(workload, gpus, "peer", "weak", False)
        assert line["n_gpus"] == gpus and ("J slabs" in line["cpu_baseline"]["sample"]) == (gpus > 1)
        assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port"
        assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
        assert line["also"] == {"unavailable": "oracle/_ref is not built"}


def test_baselines_beside_the_gpu_numbers(monkeypatch):
    """cpu_baseline is always there; the optional comparisons run inside the wall-clock allowance
    only, and a failing baseline does not raise."""
    import argparse

    fake = dict(name="k", isa="native", mean_s=0.05, sweeps=5, threads=16, tried=["k 50.0 ms"])
    monkeypatch.setattr(bench, "time_reference", lambda *a, **k: dict(fake))
    monkeypatch.setattr(bench, "cpu_other_configs", lambda *a, **k: {"x": 1})
    monkeypatch.setattr(bench, "reference_gpu", lambda workload, ms: {"ms": 2 * ms})
    args = argparse.Namespace(workload="hdiff", optional_until_s=1e9)
    keys = bench.baselines_beside(args, 8_000_000_000, 1.25)
    assert keys["cpu_baseline"]["value"] == pytest.approx(160.0) and keys["cpu_baseline"]["cores"] == 16
    assert keys["cpu_baseline"]["kind"] == "reference" and "reference OpenMP kernel k" in keys["cpu_baseline"]["sample"]
    assert keys["cpu_baseline"]["also"] == {"x": 1} and keys["reference_gpu"] == {"ms": 2.5}

    args.optional_until_s = 0.0
    keys = bench.baselines_beside(args, 8_000_000_000, 1.25)
    assert keys["cpu_baseline"]["value"] == pytest.approx(160.0)
    assert "skipped" in keys["cpu_baseline"]["also"]["unavailable"] and "skipped" in keys["reference_gpu"]["unavailable"]

    def broken(*a, **k):
        raise RuntimeError("no compiler")

    monkeypatch.setattr(bench, "time_reference", broken)
    monkeypatch.setattr(bench, "cpu_other_configs", broken)
    args.optional_until_s = 1e9
    keys = bench.baselines_beside(args, 8_000_000_000, 1.25)
    assert keys["cpu_baseline"]["value"] is None and "no compiler" in keys["cpu_baseline"]["sample"]
    assert "no compiler" in keys["cpu_baseline"]["also"]["unavailable"]
