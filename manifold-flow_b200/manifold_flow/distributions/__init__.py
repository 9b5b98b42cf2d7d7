from .base import Distribution, NoMeanException  # noqa: F401
from .normal import RescaledNormal, StandardNormal  # noqa: F401
