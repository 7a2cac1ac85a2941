from . import mesh, utils  # noqa: F401
