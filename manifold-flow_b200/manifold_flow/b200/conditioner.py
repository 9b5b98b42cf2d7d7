"""Dense contractions of the conditioner networks (kernel (b) of the north star).

Round-1 state: the GEMMs / convolutions are issued as plain library calls (cuBLAS / cuDNN through
ATen) in full fp32 - TF32 is switched off around them unless ``config.conditioner_math == "tf32"`` -
so that the spline kernels see bit-faithful parameters.  The hand-written tcgen05 + TMA
implicit-GEMM kernels replace these calls behind the same four functions.
"""
import contextlib

import torch
from torch.nn import functional as F

from .cabi import require_cuda
from .config import config


def _math_mode():
    """cuDNN convolutions default to TF32 (torch.backends.cudnn.allow_tf32 = True), which is not
    parity-clean (1e-3 relative).  The flags are process-global and are also read by the BACKWARD
    kernels that autograd launches later, so they are set (not scoped) on every call."""
    want_tf32 = config.conditioner_math == "tf32"
    if torch.backends.cudnn.allow_tf32 != want_tf32:
        torch.backends.cudnn.allow_tf32 = want_tf32
    if torch.backends.cuda.matmul.allow_tf32 != want_tf32:
        torch.backends.cuda.matmul.allow_tf32 = want_tf32
    return contextlib.nullcontext()


def linear(x, layer):
    require_cuda(x)
    with _math_mode():
        return F.linear(x, layer.weight, layer.bias)


def conv(x, layer):
    require_cuda(x)
    with _math_mode():
        return F.conv2d(x, layer.weight, layer.bias, padding=layer.padding)


def residual_block_linear(x, lin0, lin1, gate_layer, context):
    """x + [glu]( lin1(relu(lin0(relu(x)))) , gate(context) )   nn/resnet.py:27-40"""
    require_cuda(x)
    with _math_mode():
        t = F.linear(F.relu(x), lin0.weight, lin0.bias)
        t = F.linear(F.relu(t), lin1.weight, lin1.bias)
        if gate_layer is not None:
            t = F.glu(torch.cat((t, F.linear(context, gate_layer.weight, gate_layer.bias)), dim=1), dim=1)
        return x + t


def residual_block_conv(x, conv0, conv1, gate_layer, context):
    """nn/resnet.py:91-104"""
    require_cuda(x)
    with _math_mode():
        t = F.conv2d(F.relu(x), conv0.weight, conv0.bias, padding=1)
        t = F.conv2d(F.relu(t), conv1.weight, conv1.bias, padding=1)
        if gate_layer is not None:
            t = F.glu(torch.cat((t, F.conv2d(context, gate_layer.weight, gate_layer.bias)), dim=1), dim=1)
        return x + t
