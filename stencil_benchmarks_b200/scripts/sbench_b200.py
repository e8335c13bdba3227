"""`sbench` with the B200 backend registered.

The reference's entry point (stencil_benchmarks/scripts/sbench.py:34-38) imports its benchmark
collection and starts the CLI; this one additionally imports the B200 collection, whose classes
then show up as `stencils b200 ...` / `stream b200 native`.  Needs the reference package
(`stencil_benchmarks`) importable.

    python -m stencil_benchmarks_b200.scripts.sbench_b200 stencils b200 horizontal-diffusion fused --help
"""

import sys


def main():
    try:
        import stencil_benchmarks.benchmarks_collection  # noqa: F401
        from stencil_benchmarks.cli import main as cli_main
    except ImportError as error:
        sys.exit(f"the sbench CLI belongs to the reference package, which is not importable: {error}")
    import stencil_benchmarks_b200.benchmarks_collection  # noqa: F401

    cli_main()


if __name__ == "__main__":
    main()
