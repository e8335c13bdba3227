from . import b200  # noqa: F401
