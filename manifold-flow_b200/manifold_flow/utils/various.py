"""Mask builders and small helpers the model factories import
(reference: manifold_flow/utils/various.py; only the functions the hot path and
experiments/architectures use are provided)."""
import numpy as np
import torch


def product(x):
    """utils/various.py `product`: number of elements of an int or a shape tuple."""
    try:
        out = 1
        for v in x:
            out *= int(v)
        return out
    except TypeError:
        return x


def is_int(x):
    return isinstance(x, (int, np.integer)) and not isinstance(x, bool)


def is_bool(x):
    return isinstance(x, bool)


def is_positive_int(x):
    return is_int(x) and x > 0


def is_nonnegative_int(x):
    return is_int(x) and x >= 0


def sum_except_batch(x, num_batch_dims=1):
    """:329-334"""
    if not is_nonnegative_int(num_batch_dims):
        raise TypeError("Number of batch dimensions must be a non-negative integer.")
    dims = list(range(num_batch_dims, x.ndimension()))
    return torch.sum(x, dim=dims) if dims else x


def split_leading_dim(x, shape):
    return torch.reshape(x, torch.Size(shape) + x.shape[1:])


def get_num_parameters(model):
    return sum(p.numel() for p in model.parameters())


def _mask(features, active):
    mask = torch.zeros(features, dtype=torch.uint8)
    mask[active] = 1
    return mask


def create_split_binary_mask(features, n_active):
    """:397-407 - first n_active entries set."""
    return _mask(features, slice(0, n_active))


def create_alternating_binary_mask(features, even=True):
    """:410-421 - even (or odd) positions set."""
    return _mask(features, slice(0 if even else 1, None, 2))


def create_mid_split_binary_mask(features):
    """:424-434 - first ceil(features / 2) entries set."""
    return _mask(features, slice(0, (features + 1) // 2))


def create_random_binary_mask(features):
    """:437-450"""
    n = (features + 1) // 2
    idx = torch.multinomial(torch.ones(features), num_samples=n, replacement=False)
    return _mask(features, idx)


def create_mlt_channel_mask(features, channels_per_level=(1, 2, 4, 8), resolution=64):
    """:453-469 - the first `channels` channels of every multi-scale level's flattened output."""
    mask = torch.zeros(features, dtype=torch.uint8)
    pos, remaining, res = 0, features, resolution
    for channels in channels_per_level:
        remaining //= 2
        res //= 2
        mask[pos : pos + int(channels) * res * res] = 1
        pos += remaining
    assert pos <= features
    return mask
