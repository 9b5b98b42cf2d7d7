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
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)


class GradientAllReducer:
    """Averages the gradients of ``params`` over all ranks.

    ``params`` is the parameter group being trained in the current phase (e.g. only
    ``model.inner_transform.parameters()`` in phase 2 of the sequential M-flow training,
    experiments/train.py:287-297), in any order.  Gradients live in flat fp32 buckets of about
    ``bucket_mb`` MB filled in reverse parameter order (the order backward produces them)."""

    def __init__(self, params, bucket_mb=32.0, overlap=True):
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        self.buckets = []  # (flat, [params])
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
        self._pending = []
        self._handles = []
        if self.overlap:
            for bi, (_, plist) in enumerate(self.buckets):
                for p in plist:
                    p.register_post_accumulate_grad_hook(self._make_hook(bi))
        self._remaining = [len(pl) for _, pl in self.buckets]

    def _close(self, plist):
        n = sum(p.numel() for p in plist)
        flat = torch.zeros(n, dtype=plist[0].dtype, device=plist[0].device)
        off = 0
        for p in plist:
            p.grad = flat[off : off + p.numel()].view_as(p)  # autograd accumulates in place into the bucket
            off += p.numel()
        self.buckets.append((flat, plist))

    def _make_hook(self, bi):
        def hook(_param):
            self._remaining[bi] -= 1
            if self._remaining[bi] == 0:
                flat = self.buckets[bi][0]
                self._handles.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))

        return hook

    def zero_grad(self):
        for flat, _ in self.buckets:
            flat.zero_()
        self._remaining = [len(pl) for _, pl in self.buckets]

    def finish(self):
        """Call after backward: waits for (or issues) the all-reduces and divides by the world size."""
        if self.world == 1:
            return
        if self.overlap:
            # parameters that received no gradient this step never fire their hook
            for bi, rem in enumerate(self._remaining):
                if rem > 0:
                    self._handles.append(dist.all_reduce(self.buckets[bi][0], op=dist.ReduceOp.SUM, async_op=True))
            for h in self._handles:
                h.wait()
            self._handles = []
        else:
            for flat, _ in self.buckets:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        for flat, _ in self.buckets:
            flat.div_(self.world)

    def payload_bytes(self):
        return sum(flat.numel() * flat.element_size() for flat, _ in self.buckets)
