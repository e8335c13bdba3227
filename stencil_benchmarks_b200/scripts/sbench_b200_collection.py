"""Curated B200 parameter sets -> CSV, in the format of the reference's collection scripts.

Modelled on stencil_benchmarks/scripts/sbench_h100_collection.py:56-157: one click command per
benchmark family, every configuration swept over the domains 32^2 ... 2048^2 x 80 through the
reference's `tools.multirun.run_scaling_benchmark` (multirun.py:75-92), results written as CSV
that `sbench-analyze` reads.  Needs the reference package importable.

    python -m stencil_benchmarks_b200.scripts.sbench_b200_collection horizontal-diffusion-bandwidth out.csv
"""

import sys

try:
    import click
    from stencil_benchmarks.tools.multirun import Configuration, default_kwargs, run_scaling_benchmark
except ImportError as error:  # pragma: no cover - needs the reference
    sys.exit(f"the collection script drives the reference's multirun tool, which is not importable: {error}")

from stencil_benchmarks_b200.benchmarks_collection.stencils.b200 import (
    basic,
    horizontal_diffusion as hdiff,
    vertical_advection as vadv,
)


def domains(k=80):
    for exponent in range(5, 12):
        yield 2**exponent, 2**exponent, k


common = default_kwargs(verify=False, dry_runs=1, alignment=128, dtype="float64")

# family -> (levels k, extra keyword arguments, [(name, class, class arguments), ...])
FAMILIES = {
    "basic-bandwidth": (80, dict(halo=(1, 1, 1)), [
        ("empty", basic.Empty, {}),
        ("copy", basic.Copy, {}),
        ("avg-i", basic.OnesidedAverage, dict(axis=0)),
        ("avg-j", basic.OnesidedAverage, dict(axis=1)),
        ("avg-k", basic.OnesidedAverage, dict(axis=2)),
        ("sym-avg-i", basic.SymmetricAverage, dict(axis=0)),
        ("sym-avg-j", basic.SymmetricAverage, dict(axis=1)),
        ("sym-avg-k", basic.SymmetricAverage, dict(axis=2)),
        ("lap-ij", basic.Laplacian, dict(along_x=True, along_y=True, along_z=False)),
    ]),
    "horizontal-diffusion-bandwidth": (80, {}, [("fused", hdiff.Fused, {})]),
    "vertical-advection-bandwidth": (160, {}, [
        ("thomas-onchip", vadv.Thomas, dict(coefficients="auto")),
        ("thomas-global", vadv.Thomas, dict(coefficients="global")),
        ("thomas-uvw", vadv.Thomas, dict(coefficients="auto", all_components=True)),
    ]),
    # one domain partitioned over 1, 2, 4, 8 GPUs of the box (as many as it has): strong scaling
    "horizontal-diffusion-multi-gpu": (80, {}, [
        (f"partitioned-{gpus}", hdiff.Partitioned, dict(gpus=gpus)) for gpus in (1, 2, 4, 8)
    ]),
}


@click.group()
def main():
    pass


def _family_command(family):
    levels, extra, members = FAMILIES[family]

    @main.command(name=family)
    @click.argument("output", type=click.Path())
    @click.option("--executions", "-e", type=int, default=101)
    @click.option("--option", "-o", multiple=True)
    def command(output, executions, option):
        from stencil_benchmarks_b200 import capi

        kwargs = common(option, **extra)
        configurations = [Configuration(cls, name=name, **cls_kwargs, **kwargs)
                          for name, cls, cls_kwargs in members
                          if cls_kwargs.get("gpus", 1) <= max(capi.device_count(), 1)]
        run_scaling_benchmark(configurations, executions, domain_range=domains(levels)).to_csv(output)

    return command


for _family in FAMILIES:
    _family_command(_family)


if __name__ == "__main__":
    main()
