from . import various  # noqa: F401
