"""Export list of the reference (manifold_flow/transforms/__init__.py:1-55): hot-path classes are
implemented on the B200 kernels, the rest are importable stubs that raise on construction."""
from .base import InverseNotAvailable, InputOutsideDomain, Transform, CompositeTransform, MultiscaleCompositeTransform, InverseTransform

from .out_of_scope import (
    MaskedAffineAutoregressiveTransform,
    MaskedPiecewiseLinearAutoregressiveTransform,
    MaskedPiecewiseQuadraticAutoregressiveTransform,
    MaskedPiecewiseCubicAutoregressiveTransform,
    MaskedPiecewiseRationalQuadraticAutoregressiveTransform,
    QRLinear,
    SVDLinear,
    CompositeCDFTransform,
    LeakyReLU,
    Logit,
    PiecewiseLinearCDF,
    PiecewiseQuadraticCDF,
    PiecewiseCubicCDF,
    PiecewiseRationalQuadraticCDF,
    Sigmoid,
    Tanh,
    HouseholderSequence,
    AffineCouplingTransform,
    AdditiveCouplingTransform,
    PiecewiseLinearCouplingTransform,
    PiecewiseQuadraticCouplingTransform,
    PiecewiseCubicCouplingTransform,
    ElementwisePiecewiseRationalQuadraticTransform,
    ConditionalAffineScalarTransform,
    SphericalCoordinates,
)

from .linear import NaiveLinear, Linear
from .lu import LULinear
from .nonlinearities import LogTanh
from .normalization import BatchNorm, ActNorm
from .permutations import Permutation, RandomPermutation, ReversePermutation, MaskBasedPermutation
from .coupling import CouplingTransform, PiecewiseRationalQuadraticCouplingTransform
from .standard import IdentityTransform, AffineScalarTransform
from .reshape import SqueezeTransform, ReshapeTransform
from .conv import OneByOneConvolution
from .projections import Projection, ProjectionSplit
from .partial import PartialTransform
from . import splines
