"""Names the reference exports that no BASELINE config reaches.
They import so that ``experiments/architectures`` loads unchanged, and fail loudly if constructed."""


def _stub(name, why):
    class _OutOfScope:
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(f"manifold_flow.transforms.{name} is outside the B200 hot-path scope: {why}")

    _OutOfScope.__name__ = _OutOfScope.__qualname__ = name
    return _OutOfScope


_COUPLING = "--outertransform/--innertransform default to rq-coupling and no shipped config overrides them"
_AUTOREG = "the *-autoregressive transform types are opt-in flags no config uses"
_CDF = "only reached through preprocessing='realnvp' / apply_unconditional_transform=True, which create_model never sets"

AffineCouplingTransform = _stub("AffineCouplingTransform", _COUPLING)
AdditiveCouplingTransform = _stub("AdditiveCouplingTransform", _COUPLING)
PiecewiseLinearCouplingTransform = _stub("PiecewiseLinearCouplingTransform", _COUPLING)
PiecewiseQuadraticCouplingTransform = _stub("PiecewiseQuadraticCouplingTransform", _COUPLING)
PiecewiseCubicCouplingTransform = _stub("PiecewiseCubicCouplingTransform", _COUPLING)
MaskedAffineAutoregressiveTransform = _stub("MaskedAffineAutoregressiveTransform", _AUTOREG)
MaskedPiecewiseLinearAutoregressiveTransform = _stub("MaskedPiecewiseLinearAutoregressiveTransform", _AUTOREG)
MaskedPiecewiseQuadraticAutoregressiveTransform = _stub("MaskedPiecewiseQuadraticAutoregressiveTransform", _AUTOREG)
MaskedPiecewiseCubicAutoregressiveTransform = _stub("MaskedPiecewiseCubicAutoregressiveTransform", _AUTOREG)
MaskedPiecewiseRationalQuadraticAutoregressiveTransform = _stub("MaskedPiecewiseRationalQuadraticAutoregressiveTransform", _AUTOREG)
QRLinear = _stub("QRLinear", "unreachable in the reference (uses removed torch.trtrs)")
SVDLinear = _stub("SVDLinear", "--lineartransform svd is not used by any config")
HouseholderSequence = _stub("HouseholderSequence", "only used by SVDLinear / QRLinear")
CompositeCDFTransform = _stub("CompositeCDFTransform", _CDF)
LeakyReLU = _stub("LeakyReLU", _CDF)
Logit = _stub("Logit", _CDF)
Sigmoid = _stub("Sigmoid", _CDF)
Tanh = _stub("Tanh", _CDF)
PiecewiseLinearCDF = _stub("PiecewiseLinearCDF", _CDF)
PiecewiseQuadraticCDF = _stub("PiecewiseQuadraticCDF", _CDF)
PiecewiseCubicCDF = _stub("PiecewiseCubicCDF", _CDF)
PiecewiseRationalQuadraticCDF = _stub("PiecewiseRationalQuadraticCDF", _CDF)
ElementwisePiecewiseRationalQuadraticTransform = _stub("ElementwisePiecewiseRationalQuadraticTransform", "not constructed by experiments/architectures")
ConditionalAffineScalarTransform = _stub("ConditionalAffineScalarTransform", "not constructed by experiments/architectures")
SphericalCoordinates = _stub("SphericalCoordinates", "FOM baseline (--specified) only, no BASELINE config")
