"""Abstract flow (reference: manifold_flow/flows/base.py:7-34)."""
import functools
import logging

import torch
from torch import nn

logger = logging.getLogger(__name__)


def _map_tensors(obj, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, tuple):
        return tuple(_map_tensors(o, fn) for o in obj)
    if isinstance(obj, list):
        return [_map_tensors(o, fn) for o in obj]
    return obj


def host_staged(method):
    """Device placement at the model boundary.  ``experiments/evaluate.py:347-349`` loads the checkpoint with
    ``map_location=cpu``, never moves the model and feeds CPU tensors (:211-224), then calls ``.detach().numpy()`` on
    the results.  The kernels only run on the GPU, so a public call that receives host tensors moves the model to the
    current CUDA device (once, in place), copies the arguments to it, runs the same CUDA path and returns the results
    on the caller's device.  This is placement, not a CPU implementation: without a CUDA device the call raises."""

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        tensors = [a for a in list(args) + list(kwargs.values()) if torch.is_tensor(a)]
        if not tensors or all(t.is_cuda for t in tensors):
            return method(self, *args, **kwargs)
        if not torch.cuda.is_available():
            from ..b200.cabi import MflowError

            raise MflowError("manifold_flow (B200 build) needs a CUDA device: there is no CPU implementation of the flow")
        home = next((t.device for t in tensors if not t.is_cuda), torch.device("cpu"))
        param = next(self.parameters(), None)
        if param is not None and not param.is_cuda:
            logger.info("Moving the model to cuda:%d (host tensors were passed to a CUDA-only model)", torch.cuda.current_device())
            self.to(torch.device("cuda", torch.cuda.current_device()))
            param = next(self.parameters())
        dev = param.device if param is not None else torch.device("cuda", torch.cuda.current_device())
        to_dev = lambda t: t.to(dev, dtype=torch.float32 if t.is_floating_point() else None)
        with torch.cuda.device(dev):
            out = method(self, *_map_tensors(args, to_dev), **{k: _map_tensors(v, to_dev) for k, v in kwargs.items()})
        return _map_tensors(out, lambda t: t.to(home))

    return wrapper


class BaseFlow(nn.Module):
    def forward(self, x, context=None):
        raise NotImplementedError

    def encode(self, x, context=None):
        raise NotImplementedError

    def decode(self, u, context=None):
        raise NotImplementedError

    def project(self, x, context=None):
        return self.decode(self.encode(x, context), context)

    def log_prob(self, x, context=None):
        raise NotImplementedError

    def sample(self, u=None, n=1, context=None):
        raise NotImplementedError

    def _report_model_parameters(self):
        n_all = sum(p.numel() for p in self.parameters())
        n_train = sum(p.numel() for p in self.parameters() if p.requires_grad)
        logger.info("Model has %.1f M parameters (%.1f M trainable) with an estimated size of %.1f MB", n_all / 1e6, n_train / 1e6, n_all * 4 / 1e6)
