"""Conditioner networks of the coupling layers (reference: manifold_flow/nn/resnet.py).

Parameter names, shapes, creation order (hence seeded initialisation) and forward semantics match
``ResidualNet`` (:43-72) / ``ConvResidualNet`` (:107-142) so reference checkpoints load unchanged.
Batch norm and dropout are not part of any BASELINE config (``batchnorm = False``, ``dropout = 0``)
and are rejected.  The dense contractions are dispatched through ``manifold_flow.b200.conditioner``.
"""
import torch
from torch import nn
from torch.nn import functional as F
from torch.nn import init

from ..b200 import conditioner as K


def _reject(dropout_probability, use_batch_norm):
    if use_batch_norm:
        raise NotImplementedError("use_batch_norm=True is outside the B200 hot-path scope (no BASELINE config uses it)")
    if dropout_probability:
        raise NotImplementedError("dropout_probability > 0 is outside the B200 hot-path scope (no BASELINE config uses it)")


class ResidualBlock(nn.Module):
    """relu -> linear -> relu -> linear (-> GLU gate from the context) + skip.  resnet.py:9-40"""

    def __init__(self, features, context_features, activation=F.relu, dropout_probability=0.0, use_batch_norm=False,
                 zero_initialization=True):
        super().__init__()
        _reject(dropout_probability, use_batch_norm)
        if activation is not F.relu:
            raise NotImplementedError("only ReLU conditioners are in scope")
        if context_features is not None:
            self.context_layer = nn.Linear(context_features, features)
        self.linear_layers = nn.ModuleList([nn.Linear(features, features) for _ in range(2)])
        if zero_initialization:
            init.uniform_(self.linear_layers[-1].weight, -1e-3, 1e-3)
            init.uniform_(self.linear_layers[-1].bias, -1e-3, 1e-3)

    def forward(self, inputs, context=None):
        gate = self.context_layer if context is not None else None
        return K.residual_block_linear(inputs, self.linear_layers[0], self.linear_layers[1], gate, context)


class ResidualNet(nn.Module):
    """resnet.py:43-72"""

    def __init__(self, in_features, out_features, hidden_features, context_features=None, num_blocks=2, activation=F.relu,
                 dropout_probability=0.0, use_batch_norm=False):
        super().__init__()
        self.hidden_features = hidden_features
        self.context_features = context_features
        self.initial_layer = nn.Linear(in_features + (context_features or 0), hidden_features)
        self.blocks = nn.ModuleList(
            [ResidualBlock(hidden_features, context_features, activation, dropout_probability, use_batch_norm) for _ in range(num_blocks)]
        )
        self.final_layer = nn.Linear(hidden_features, out_features)

    def forward(self, inputs, context=None):
        out = K.residual_net_planes(self, inputs, context)   # large batches: tensor-core path
        if out is not None:
            return out
        out = K.residual_net_fused(self, inputs, context)    # everything else: the whole net in one launch
        if out is not None:
            return out
        temps = K.linear(inputs if context is None else torch.cat((inputs, context), dim=1), self.initial_layer)
        for block in self.blocks:
            temps = block(temps, context=context)
        return K.linear(temps, self.final_layer)


class ConvResidualBlock(nn.Module):
    """relu -> 3x3 conv -> relu -> 3x3 conv (-> GLU gate) + skip.  resnet.py:75-104"""

    def __init__(self, channels, context_channels=None, activation=F.relu, dropout_probability=0.0, use_batch_norm=False,
                 zero_initialization=True):
        super().__init__()
        _reject(dropout_probability, use_batch_norm)
        if activation is not F.relu:
            raise NotImplementedError("only ReLU conditioners are in scope")
        if context_channels is not None:
            self.context_layer = nn.Conv2d(context_channels, channels, kernel_size=1, padding=0)
        self.conv_layers = nn.ModuleList([nn.Conv2d(channels, channels, kernel_size=3, padding=1) for _ in range(2)])
        if zero_initialization:
            init.uniform_(self.conv_layers[-1].weight, -1e-3, 1e-3)
            init.uniform_(self.conv_layers[-1].bias, -1e-3, 1e-3)

    def forward(self, inputs, context=None):
        gate = self.context_layer if context is not None else None
        return K.residual_block_conv(inputs, self.conv_layers[0], self.conv_layers[1], gate, context)


class ConvResidualNet(nn.Module):
    """resnet.py:107-142"""

    def __init__(self, in_channels, out_channels, hidden_channels, context_channels=None, num_blocks=2, activation=F.relu,
                 dropout_probability=0.0, use_batch_norm=False):
        super().__init__()
        self.context_channels = context_channels
        self.hidden_channels = hidden_channels
        self.initial_layer = nn.Conv2d(in_channels + (context_channels or 0), hidden_channels, kernel_size=1, padding=0)
        self.blocks = nn.ModuleList(
            [ConvResidualBlock(hidden_channels, context_channels, activation, dropout_probability, use_batch_norm)
             for _ in range(num_blocks)]
        )
        self.final_layer = nn.Conv2d(hidden_channels, out_channels, kernel_size=1, padding=0)

    def forward(self, inputs, context=None):
        return K.conv(self.hidden(inputs, context), self.final_layer)

    def hidden(self, inputs, context=None):
        """Everything but the final 1x1 convolution (:130-140).  The coupling layer calls this in evaluation passes and
        runs the final layer with its spline in the epilogue (``ops.conv_rqs``)."""
        if context is not None and context.dim() < inputs.dim():  # vector context -> constant image planes (:130-133)
            if context.dim() != 2 or inputs.dim() != 4:
                raise ValueError("context must be [B, F] for [B, C, H, W] inputs")
            context = context[:, :, None, None].expand(-1, -1, inputs.size(2), inputs.size(3))
        temps = K.conv(inputs if context is None else torch.cat((inputs, context), dim=1), self.initial_layer)
        for block in self.blocks:
            temps = block(temps, context)
        return temps
