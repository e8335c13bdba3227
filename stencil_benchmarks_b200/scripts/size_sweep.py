import statistics, sys
sys.path.insert(0, '.')
from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import horizontal_diffusion as hd, vertical_advection as va, basic
for cls, name, kw, nz in [(hd.Fused, 'hdiff', {}, 80), (va.Thomas, 'vadv', {}, 80), (va.Thomas, 'vadv', {}, 160), (basic.Copy, 'copy', dict(halo=(1,1,1)), 80), (basic.Laplacian, 'lap-ij', dict(halo=(1,1,1)), 80)]:
    for dtype in ('float64', 'float32'):
        row = []
        for e in range(5, 12):
            n = 2 ** e
            b = cls(domain=(n, n, nz), dtype=dtype, verify=False, dry_runs=2, resident=True, **kw)
            ts = [b.run() for _ in range(7)]
            t = statistics.median(r['time'] for r in ts)
            row.append(f"{n}^2: {b.algorithmic_bytes / t / 1e9:7.0f} GB/s ({t*1e6:7.1f} us)")
            del b
        print(f"{name:7s} nz={nz:3d} {dtype}: " + " | ".join(row), flush=True)
