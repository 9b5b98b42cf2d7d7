"""Dense contractions of the conditioner networks (kernel (b) of the north star).

Image conditioners (``ConvResidualNet``): every 3x3 / 1x1 convolution - forward and data gradient - runs
on the tcgen05 tensor cores through ``mflow_conv2d_tf32x3`` (TMA-fed implicit GEMM on NCHW tensors,
3xTF32 split for fp32 accuracy, pre-activation ReLU, bias, residual add and the ReLU mask of the backward
pass fused); their weight gradients through ``mflow_conv2d_wgrad_tf32x3``.  Vector conditioners (``ResidualNet``)
take the same kernels from ``config.linear_tc_min_rows`` rows on (``residual_net_planes``); below that they are
cuBLAS GEMMs with M = batch (launch-latency bound either way), with TF32 switched off unless
``config.conditioner_math == "tf32"``.
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


class _PlanesOut:
    """Set by the coupling layer around its conditioner call: ``residual_net_planes`` then returns its NCHW result
    [rows / w^2, n_out, w, w] as it is (the spline kernel reads that layout in place, ``mflow_rqs_desc.p_hw``) instead of
    transposing it back to [rows, n_out] - that copy, and the transposition of the parameter gradient on the way back,
    moved more bytes than the spline itself (a quarter of the cfg3 step at 65 536 rows)."""
    enabled = False


@contextlib.contextmanager
def planes_output():
    prev, _PlanesOut.enabled = _PlanesOut.enabled, True
    try:
        yield
    finally:
        _PlanesOut.enabled = prev


def residual_net_planes(net, inputs, context):
    """ResidualNet (nn/resnet.py:43-72) on the tensor-core convolution kernels, or None when the shape does not qualify.

    The batch rows become the pixels of [rows / w^2, C, w, w] tensors (one blocked transposition in, one out); every
    Linear layer is then a 1x1 convolution with the pre-activation ReLU, the bias and the skip connection fused,
    its data gradient runs through the same kernel and its weight gradient through the pixels-as-K kernel."""
    rows = inputs.shape[0]
    if not (config.conv_tensor_cores and config.conditioner_math == "fp32" and inputs.dim() == 2 and rows >= config.linear_tc_min_rows):
        return None
    w = next((c for c in (32, 16, 8, 4) if rows % (c * c) == 0), 0)
    if not w:
        return None
    require_cuda(inputs)

    def to_planes(t):
        return t.reshape(rows // (w * w), w * w, t.shape[1]).transpose(1, 2).contiguous().view(-1, t.shape[1], w, w)

    def lin(x, layer, **kw):
        weight = layer.weight
        return ops.conv_tc(x, weight.view(weight.shape[0], weight.shape[1], 1, 1), layer.bias, weight_owner=weight, **kw)

    h = lin(to_planes(inputs if context is None else torch.cat((inputs, context), dim=1)), net.initial_layer)
    ctx_planes = None if context is None else to_planes(context)
    for block in net.blocks:
        lin0, lin1 = block.linear_layers
        # the block input feeds the branch and the skip connection: their gradients are added (and the sum's absolute
        # maximum published) by one launch instead of autograd's accumulation plus an absmax pass
        hb, hs = ops.fork(h) if (h.requires_grad and torch.is_grad_enabled()) else (h, h)
        t = lin(hb, lin0, relu_in=True)
        if context is None:
            h = lin(t, lin1, relu_in=True, residual=hs)
        else:  # glu(cat(t, gate)) = t * sigmoid(gate)   resnet.py:38-39
            h = ops.glu_residual(hs, lin(t, lin1, relu_in=True), lin(ctx_planes, block.context_layer))
    out = lin(h, net.final_layer)
    if _PlanesOut.enabled:
        return out
    return out.view(out.shape[0], out.shape[1], w * w).transpose(1, 2).reshape(rows, out.shape[1])


def residual_net_fused(net, inputs, context):
    """ResidualNet (nn/resnet.py:43-72) as one fused launch (``mflow_mlp_forward``: all layers, ReLUs, GLU context gates and
    skip connections; the identity features are read in place from the coupling layer's input), or None when the shape
    does not qualify (hidden > 128 features, context that needs a gradient, ...)."""
    if not (config.fuse_vector_conditioner and config.conditioner_math == "fp32" and ops.mlp_fused_supported(net, inputs, context)):
        return None
    require_cuda(inputs)
    return ops.residual_net_fused(net, inputs, context)


def residual_block_conv(x, conv0, conv1, gate_layer, context):
    """nn/resnet.py:91-104"""
    require_cuda(x)
    with _math_mode():
        if gate_layer is None and _use_tc(x, conv0):
            xb, xs = ops.fork(x) if (x.requires_grad and torch.is_grad_enabled()) else (x, x)
            t = ops.conv_tc(xb, conv0.weight, conv0.bias, relu_in=True)
            return ops.conv_tc(t, conv1.weight, conv1.bias, residual=xs, relu_in=True)
        t = F.conv2d(F.relu(x), conv0.weight, conv0.bias, padding=1)
        t = F.conv2d(F.relu(t), conv1.weight, conv1.bias, padding=1)
        if gate_layer is not None:
            t = F.glu(torch.cat((t, F.conv2d(context, gate_layer.weight, gate_layer.bias)), dim=1), dim=1)
        return x + t
