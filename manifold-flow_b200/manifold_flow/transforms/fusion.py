"""Chain evaluation with linear-glue fusion (kernel (c) of the north star).

ActNorm, the channel permutation + LU map of the 1x1 convolution, ``RandomPermutation`` +
``LULinear`` on vectors are all per-pixel affine maps over the channel axis.  A run of such layers is
composed into ONE dense C x C matrix (a handful of tiny differentiable torch ops on [C, C] tensors,
rebuilt once per optimizer step) and applied by ONE ``mflow_chanmix_apply`` launch that reads and
writes the activation exactly once - the reference issues ~10 ATen launches and 4 activation
round trips for the same step (normalization.py:99-131, conv.py:16-55, lu.py:44-95,
permutations.py:73).
"""
import torch

from ..b200 import ops
from ..b200.config import config

MAX_FUSED_CHANNELS = 96  # chanmix weight-gradient kernel limit (shared memory)

_pass_id = 0


def begin_pass(model=None):
    """Called at the start of every model-level forward: dense L/U and composed matrices built
    under autograd are reused by all transforms evaluated within the same pass.  The first pass of a
    model also groups its LU layers for batched matrix builds (``lu.LUGroup``)."""
    global _pass_id
    _pass_id += 1
    ops._amax_pool.reset()   # a model-level pass starts on a fresh block of zeroed amax slots
    if model is not None and config.batch_lu_builds and not model.__dict__.get("_lu_groups_ready", False):
        from .lu import assign_lu_groups

        assign_lu_groups(model)
        model.__dict__["_lu_groups_ready"] = True
    return _pass_id


def current_pass():
    return _pass_id


def member_views(holder, slot, stacked):
    """Per-member views of the stacked [G, ...] tensors a group cache returned, through ONE ``unbind`` per tensor.
    ``stacked[k][i]`` is a ``select`` whose backward allocates a zero [G, ...] tensor, copies one slice and adds it to the
    running gradient: three launches per member, map and step (297 of cfg4's launches); ``unbind``'s backward is a single
    ``stack``.  Memoised on the identity of the cached tuple (rebuilt whenever the group cache rebuilds)."""
    ent = holder.__dict__.get(slot)
    if ent is None or ent[0] is not stacked:
        ent = (stacked, tuple(t.unbind(0) for t in stacked))
        holder.__dict__[slot] = ent
    return ent[1]


def _linear_capable(t, x):
    return getattr(t, "_affine_map", None) is not None and t._can_fuse(x)


def apply_affine(maps, x):
    """Compose a run of per-pixel affine maps (M, m, logabsdet) and apply them in one launch."""
    mat, bias, lad = maps[0]
    for m2, b2, l2 in maps[1:]:
        bias = m2 @ bias + b2
        mat = m2 @ mat
        lad = lad + l2
    spatial = x.shape[2] * x.shape[3] if x.dim() == 4 else 1
    return ops.chanmix(x, mat, bias), (lad * spatial if spatial != 1 else lad)


def apply_chain(transforms, x, context, inverse):
    """Evaluate ``transforms`` in order (reversed when ``inverse``) and sum their log-dets.  Runs of
    affine layers are merged into one launch (``config.fuse_linear_glue``; off = one launch each)."""
    from .base import _add_lad

    seq = transforms[::-1] if inverse else transforms
    total = None
    run = []  # pending affine maps (M, m, lad) to compose
    run_layers = []

    def flush(x, total):
        if not run:
            return x, total
        maps = list(run)
        if len(run_layers) == 2:
            # (ActNorm, 1x1 convolution) pairs of a model are composed for all pairs at once (PairGroup)
            act, conv = (run_layers[0], run_layers[1]) if not inverse else (run_layers[1], run_layers[0])
            pairs = getattr(act, "_pair_group", None)
            if pairs is not None and getattr(act, "_pair_partner", None) is conv and pairs.ready():
                mat, bias, lad = member_views(pairs, "_members_inv" if inverse else "_members_fwd", pairs.composed(inverse))
                maps = [(mat[act._pair_index], bias[act._pair_index], lad[act._pair_index])]
        y, lad = apply_affine(maps, x)
        run.clear()
        run_layers.clear()
        return y, _add_lad(total, lad)

    for t in seq:
        if _linear_capable(t, x):
            if not inverse and getattr(t, "_needs_data_init", None) is not None and t._needs_data_init():
                x, total = flush(x, total)  # ActNorm's data-dependent init must see its real input
                t._initialize(x)
            run.append(t._affine_map(inverse))
            run_layers.append(t)
            if not config.fuse_linear_glue:
                x, total = flush(x, total)
            continue
        x, total = flush(x, total)
        x, lad = t._apply_transform(x, context, inverse)
        total = _add_lad(total, lad)
    x, total = flush(x, total)
    return x, total


class VersionedCache:
    """Caches tensors derived from parameters, keyed on the parameters' in-place version counters
    (bumped by optimizer steps / load_state_dict).  Under autograd the entry is additionally bound
    to the current model-level pass so that a freed graph is never reused."""

    def __init__(self):
        self._key = None
        self._value = None

    def get(self, params, build):
        key = (config.cache_epoch,) + tuple((p.data_ptr(), p._version) for p in params) + (
            torch.is_grad_enabled() and any(p.requires_grad for p in params),)
        if key[-1]:
            key = key + (current_pass(),)
            if current_pass() == 0:
                return build()
        if key != self._key:
            self._value = build()
            self._key = key
        return self._value


class PairGroup:
    """Composed maps of all (ActNorm -> 1x1 convolution) pairs - or (RandomPermutation -> LULinear) pairs of the
    vector flows - of one width: second(first(x)) forward, first^-1(second^-1(y)) inverse, as a few batched products
    instead of four small launches per pair."""

    def __init__(self, pairs, kind):
        self.pairs = list(pairs)   # [(first, second)]
        self.kind = kind           # "actnorm" | "perm"
        self._composed = {False: VersionedCache(), True: VersionedCache()}
        self._idx = None
        self._perm_mats = {}

    def ready(self):
        return self.kind == "perm" or self.pairs[0][0]._an_group.ready()

    def _first_maps(self, inverse, device):
        if self.kind == "actnorm":
            return tuple(t.index_select(0, self._idx[0]) for t in self.pairs[0][0]._an_group.affine_maps(inverse))
        # load_state_dict overwrites the permutation buffers in place (version bump): part of the key
        key = (bool(inverse), str(device), config.cache_epoch) + tuple((p._permutation.data_ptr(), p._permutation._version) for p, _ in self.pairs)
        if key not in self._perm_mats:
            self._perm_mats.clear()
            n = self.pairs[0][0]._permutation.numel()
            eye = torch.eye(n, dtype=torch.float32, device=device)
            self._perm_mats[key] = torch.stack([eye[(p._inverse_permutation if inverse else p._permutation).to(device), :]
                                                for p, _ in self.pairs])
        return self._perm_mats[key], None, None

    def composed(self, inverse):
        first0, second0 = self.pairs[0]

        def build():
            device = second0.lower_entries.device
            if self._idx is None or self._idx[1].device != device:
                ia = torch.tensor([a._an_index for a, _ in self.pairs], device=device) if self.kind == "actnorm" else None
                self._idx = (ia, torch.tensor([c._lu_index for _, c in self.pairs], device=device))
            ma, ba, la = self._first_maps(inverse, device)
            mc, bc, lc = (t.index_select(0, self._idx[1]) for t in second0._lu_group.affine_maps(inverse))
            if not inverse:   # x -> first -> second
                bias = bc if ba is None else torch.bmm(mc, ba[:, :, None]).squeeze(2) + bc
                return torch.bmm(mc, ma), bias, (lc if la is None else la + lc)
            bias = torch.bmm(ma, bc[:, :, None]).squeeze(2)   # y -> second^-1 -> first^-1
            return torch.bmm(ma, mc), (bias if ba is None else bias + ba), (lc if la is None else la + lc)

        params = tuple(p for _, c in self.pairs for p in (c.bias,) + c._lu_params())
        if self.kind == "actnorm":
            params = params + tuple(p for a, _ in self.pairs for p in (a.log_scale, a.shift))
            params = params + tuple(c.permutation._permutation for _, c in self.pairs)
        else:
            params = params + tuple(p._permutation for p, _ in self.pairs)
        return self._composed[bool(inverse)].get(params, build)
