"""B200 back end of the drop-in package: ctypes binding of the C ABI (``cabi``), autograd
functions over it (``ops``), data-parallel helpers (``ddp``)."""
from . import cabi, ops  # noqa: F401
from .config import config  # noqa: F401
