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


@click.group()
def main():
    pass


common_kwargs = default_kwargs(verify=False, dry_runs=1, alignment=128, dtype="float64")


def domains(k=80):
    for exponent in range(5, 12):
        yield 2**exponent, 2**exponent, k


@main.command()
@click.argument("output", type=click.Path())
@click.option("--executions", "-e", type=int, default=101)
@click.option("--option", "-o", multiple=True)
def basic_bandwidth(output, executions, option):
    kwargs = common_kwargs(option, halo=(1, 1, 1))
    configurations = [
        Configuration(basic.Empty, name="empty", **kwargs),
        Configuration(basic.Copy, name="copy", **kwargs),
        Configuration(basic.OnesidedAverage, name="avg-i", axis=0, **kwargs),
        Configuration(basic.OnesidedAverage, name="avg-j", axis=1, **kwargs),
        Configuration(basic.OnesidedAverage, name="avg-k", axis=2, **kwargs),
        Configuration(basic.SymmetricAverage, name="sym-avg-i", axis=0, **kwargs),
        Configuration(basic.SymmetricAverage, name="sym-avg-j", axis=1, **kwargs),
        Configuration(basic.SymmetricAverage, name="sym-avg-k", axis=2, **kwargs),
        Configuration(basic.Laplacian, name="lap-ij", along_x=True, along_y=True, along_z=False, **kwargs),
    ]
    run_scaling_benchmark(configurations, executions, domain_range=domains()).to_csv(output)


@main.command()
@click.argument("output", type=click.Path())
@click.option("--executions", "-e", type=int, default=101)
@click.option("--option", "-o", multiple=True)
def horizontal_diffusion_bandwidth(output, executions, option):
    kwargs = common_kwargs(option)
    configurations = [Configuration(hdiff.Fused, name="fused", **kwargs)]
    run_scaling_benchmark(configurations, executions, domain_range=domains()).to_csv(output)


@main.command()
@click.argument("output", type=click.Path())
@click.option("--executions", "-e", type=int, default=101)
@click.option("--option", "-o", multiple=True)
def vertical_advection_bandwidth(output, executions, option):
    kwargs = common_kwargs(option)
    configurations = [
        Configuration(vadv.Thomas, name="thomas-onchip", coefficients="auto", **kwargs),
        Configuration(vadv.Thomas, name="thomas-global", coefficients="global", **kwargs),
    ]
    run_scaling_benchmark(configurations, executions, domain_range=domains(k=160)).to_csv(output)


if __name__ == "__main__":
    main()
