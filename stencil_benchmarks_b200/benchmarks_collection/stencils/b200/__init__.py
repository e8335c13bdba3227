from . import basic, horizontal_diffusion, vertical_advection  # noqa: F401
