from .resnet import ConvResidualNet, ResidualNet  # noqa: F401


def _out_of_scope(name):
    class _Stub:
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(
                f"manifold_flow.nn.{name} is outside the B200 hot-path scope (SURVEY.md section 2, row 18); "
                "only ResidualNet and ConvResidualNet are provided."
            )

    _Stub.__name__ = _Stub.__qualname__ = name
    return _Stub


MLP = _out_of_scope("MLP")
UNet = _out_of_scope("UNet")
ConvAttentionNet = _out_of_scope("ConvAttentionNet")
ModifiedConvEncoder = _out_of_scope("ModifiedConvEncoder")
ConvEncoder = _out_of_scope("ConvEncoder")
ConvDecoder = _out_of_scope("ConvDecoder")
Conv2dSameSize = _out_of_scope("Conv2dSameSize")
SylvesterFlowConvEncoderNet = _out_of_scope("SylvesterFlowConvEncoderNet")
SylvesterFlowConvDecoderNet = _out_of_scope("SylvesterFlowConvDecoderNet")
