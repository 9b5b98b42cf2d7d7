"""Dense contractions of the conditioner networks (kernel (b) of the north star).

Image conditioners (``ConvResidualNet``): every 3x3 / 1x1 convolution - forward and data gradient - runs
on the tcgen05 tensor cores through ``mflow_conv2d_tf32x3`` (TMA-fed implicit GEMM on NCHW tensors,
3xTF32 split for fp32 accuracy, pre-activation ReLU, bias, residual add and the ReLU mask of the backward
pass fused).  Still library calls in this round: the weight gradients of those convolutions (cuDNN), the
small vector conditioners (``ResidualNet``: cuBLAS GEMMs with M = batch) and the GLU context path; TF32 is
switched off around them unless ``config.conditioner_math == "tf32"``.
"""
import contextlib

import torch
from torch.nn import functional as F

from . import ops
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


def _use_tc(x, layer):
    return config.conv_tensor_cores and config.conditioner_math == "fp32" and ops.conv_tc_supported(x, layer.weight)


def conv(x, layer):
    require_cuda(x)
    with _math_mode():
        if _use_tc(x, layer):
            return ops.conv_tc(x, layer.weight, layer.bias)
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
        if gate_layer is None and _use_tc(x, conv0):
            t = ops.conv_tc(x, conv0.weight, conv0.bias, relu_in=True)
            return ops.conv_tc(t, conv1.weight, conv1.bias, residual=x, relu_in=True)
        t = F.conv2d(F.relu(x), conv0.weight, conv0.bias, padding=1)
        t = F.conv2d(F.relu(t), conv1.weight, conv1.bias, padding=1)
        if gate_layer is not None:
            t = F.glu(torch.cat((t, F.conv2d(context, gate_layer.weight, gate_layer.bias)), dim=1), dim=1)
        return x + t
