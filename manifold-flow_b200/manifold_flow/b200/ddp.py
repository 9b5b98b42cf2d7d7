"""Batch-sharded data parallelism: one process per GPU, replicated parameters, one NCCL all-reduce of
the gradients per step.

The reference's only multi-GPU mechanism is single-process ``nn.parallel.data_parallel``
(experiments/training/trainer.py:577,624,674), which re-broadcasts every parameter on every step.
Here parameters stay resident on every rank; gradients accumulate straight into a few large flat
buckets (no per-parameter copies) and each bucket is all-reduced asynchronously as soon as the last
of its gradients has been produced, so the transfer over NVLink/NVSwitch overlaps the rest of the
backward pass.  Works with any ``torch.distributed`` backend (``nccl`` on GPUs, ``gloo`` in the CPU
tests)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_batch(n_global, rank, world):
    """Contiguous shard [lo, hi) of a global batch for this rank (remainder to the first ranks)."""
    base, rem = divmod(n_global, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_parameters(module, src=0):
    """Make every rank start from rank ``src``'s parameters and buffers."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return
    from .config import invalidate_caches

    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)
    invalidate_caches()   # collectives write behind the version counters: drop dense L / U, split weights, composed maps


class GradientAllReducer:
    """Averages the gradients of ``params`` over all ranks.

    ``params`` is the parameter group being trained in the current phase (e.g. only
    ``model.inner_transform.parameters()`` in phase 2 of the sequential M-flow training,
    experiments/train.py:287-297), in any order.  Gradients live in flat fp32 buckets of about
    ``bucket_mb`` MB filled in reverse parameter order (the order backward produces them).

    ``overlap=True``: a bucket is all-reduced (async, on the backend's own stream) from the post-accumulate hook of
    its last gradient, so the transfer runs under the rest of backward.  Because the hooks fire inside ``backward()``
    the same code is CUDA-graph capturable: capture ``zero_grad(); loss.backward(); finish()`` and the replayed graph
    contains the NCCL kernels forked after each bucket's last gradient (``bench.py`` does that).
    ``overlap=False``: all buckets are reduced in ``finish()`` - the mode for gradient accumulation over several
    ``backward()`` calls.

    Contract: zero the gradients with ``reducer.zero_grad()``.  ``optimizer.zero_grad()`` / ``model.zero_grad()``
    default to ``set_to_none=True``, which drops the bucket views; the hooks (and ``finish()``) detect a gradient that
    no longer aliases its slice, copy it in and re-attach it, so replicas cannot silently diverge.  A second
    ``backward()`` before ``finish()`` with ``overlap=True`` raises (the bucket would already be on the wire)."""

    def __init__(self, params, bucket_mb=32.0, overlap=True):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        # NCCL averages in the reduction itself (no separate division pass); gloo has no AVG
        self._avg = self.world > 1 and dist.get_backend() == "nccl"
        self.buckets = []  # (flat, [params])
        self._views = {}   # id(param) -> (bucket index, view)
        cap = int(bucket_mb * (1 << 20) / 4)
        cur, cur_n = [], 0
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > cap:
                self._close(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._close(cur)
        self._handles = []
        self._issued = [False] * len(self.buckets)
        self._remaining = [len(pl) for _, pl in self.buckets]
        self._hook_handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params] if self.world > 1 else []

    def close(self):
        """Detach from the parameters (remove the hooks); gradients keep living in the buckets until they are reset."""
        for h in self._hook_handles:
            h.remove()
        self._hook_handles = []

    def _close(self, plist):
        n = sum(p.numel() for p in plist)
        flat = torch.zeros(n, dtype=plist[0].dtype, device=plist[0].device)
        off = 0
        bi = len(self.buckets)
        for p in plist:
            view = flat[off : off + p.numel()].view_as(p)
            p.grad = view  # autograd accumulates in place into the bucket
            self._views[id(p)] = (bi, view)
            off += p.numel()
        self.buckets.append((flat, plist))

    def _aliases(self, p):
        g, (_, view) = p.grad, self._views[id(p)]
        return g is not None and g.data_ptr() == view.data_ptr() and g.shape == view.shape

    def _reduce(self, bi):
        flat = self.buckets[bi][0]
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._issued[bi] = True
        return dist.all_reduce(flat, op=op, async_op=True)

    def _hook(self, p):
        bi, view = self._views[id(p)]
        if not self._aliases(p):  # someone dropped the view (zero_grad(set_to_none=True)): autograd allocated a fresh tensor
            view.copy_(p.grad)
            p.grad = view
        if not self.overlap:
            return
        if self._issued[bi]:
            raise RuntimeError("GradientAllReducer(overlap=True): a second backward() reached a bucket that is already being "
                               "all-reduced; call finish() after every backward(), or construct with overlap=False to accumulate")
        self._remaining[bi] -= 1
        if self._remaining[bi] == 0:
            self._handles.append(self._reduce(bi))

    def zero_grad(self):
        """Zero the buckets, re-attach any ``p.grad`` that lost its view, re-arm the hooks."""
        for flat, _ in self.buckets:
            flat.zero_()
        for p in self.params:
            if not self._aliases(p):
                p.grad = self._views[id(p)][1]
        self._remaining = [len(pl) for _, pl in self.buckets]
        self._issued = [False] * len(self.buckets)
        self._handles = []

    def finish(self):
        """Call after backward: issues the outstanding all-reduces, waits for all of them, and averages."""
        if self.world == 1:
            return
        for p in self.params:  # gradients produced without the hook having seen them (never happens with autograd; cheap)
            if p.grad is not None and not self._aliases(p):
                _, view = self._views[id(p)]
                view.copy_(p.grad)
                p.grad = view
        # buckets not yet on the wire: overlap=False, or parameters that received no gradient this step
        for bi in range(len(self.buckets)):
            if not self._issued[bi]:
                self._handles.append(self._reduce(bi))
        for h in self._handles:
            h.wait()
        self._handles = []
        if not self._avg:
            for flat, _ in self.buckets:
                flat.div_(self.world)
        # re-arm for the next step (a caller that zeroes with optimizer.zero_grad() never calls reducer.zero_grad())
        self._remaining = [len(pl) for _, pl in self.buckets]
        self._issued = [False] * len(self.buckets)

    def payload_bytes(self):
        return sum(flat.numel() * flat.element_size() for flat, _ in self.buckets)
