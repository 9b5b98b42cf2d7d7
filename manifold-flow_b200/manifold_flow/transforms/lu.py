"""LU-parameterised linear layer (reference: manifold_flow/transforms/lu.py, linear.py:32-132).

W = L U with L unit lower triangular and U upper triangular with diagonal softplus(.) + eps
(lu.py:44-54,118-120).  Small layers (features <= 96) are dense C x C affine maps that
``fusion.apply_chain`` composes with their neighbours into one channel-mix launch.  Large layers
(``LULinear(3840)`` of the image post-processing, architectures/image_transforms.py:240-258) keep
their triangles dense on the device - rebuilt once per optimizer step instead of on every call via
host index arrays (lu.py:18-20,44-54) - and run as plain library GEMMs / triangular solves.
"""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F
from torch.nn import init

from ..b200 import ops
from ..b200.cabi import require_cuda
from ..b200.config import config
from ..utils import various
from .base import Transform
from .fusion import MAX_FUSED_CHANNELS, VersionedCache


class Linear(Transform):
    """Base of linear transforms: ``features``, ``bias`` and the (unused by default) eval cache flag
    of the reference (linear.py:32-48)."""

    def __init__(self, features, using_cache=False):
        if not various.is_positive_int(features):
            raise TypeError("Number of features must be a positive integer.")
        super().__init__()
        self.features = features
        self.bias = nn.Parameter(torch.zeros(features))
        self.using_cache = using_cache

    def use_cache(self, mode=True):
        if not various.is_bool(mode):
            raise TypeError("Mode must be boolean.")
        self.using_cache = mode  # dense factors are always cached per parameter version here


class LULinear(Linear):
    def __init__(self, features, using_cache=False, identity_init=True, eps=1e-3):
        super().__init__(features, using_cache)
        self.eps = eps
        n_tri = ((features - 1) * features) // 2
        self.lower_entries = nn.Parameter(torch.zeros(n_tri))
        self.upper_entries = nn.Parameter(torch.zeros(n_tri))
        self.unconstrained_upper_diag = nn.Parameter(torch.zeros(features))
        # flat positions of the strictly-lower / strictly-upper entries in row-major order
        # (np.tril_indices / np.triu_indices order, lu.py:18-19); not part of the state_dict
        rows, cols = np.tril_indices(features, k=-1)
        self.register_buffer("_lower_pos", torch.from_numpy(rows * features + cols), persistent=False)
        rows, cols = np.triu_indices(features, k=1)
        self.register_buffer("_upper_pos", torch.from_numpy(rows * features + cols), persistent=False)
        self._dense_cache = VersionedCache()
        self._map_cache = {False: VersionedCache(), True: VersionedCache()}
        self._initialize(identity_init)

    def _initialize(self, identity_init):
        init.zeros_(self.bias)
        if identity_init:  # lu.py:33-37
            init.zeros_(self.lower_entries)
            init.zeros_(self.upper_entries)
            init.constant_(self.unconstrained_upper_diag, float(np.log(np.exp(1 - self.eps) - 1)))
        else:
            stdv = 1.0 / np.sqrt(self.features)
            init.uniform_(self.lower_entries, -stdv, stdv)
            init.uniform_(self.upper_entries, -stdv, stdv)
            init.uniform_(self.unconstrained_upper_diag, -stdv, stdv)

    # ---- dense factors -------------------------------------------------------------------
    @property
    def upper_diag(self):
        return F.softplus(self.unconstrained_upper_diag) + self.eps

    def logabsdet(self):
        group = getattr(self, "_lu_group", None)
        if group is not None:
            return group.factors()[2][self._lu_index]
        return torch.sum(torch.log(self.upper_diag))

    def _lu_params(self):
        return (self.lower_entries, self.upper_entries, self.unconstrained_upper_diag)

    def _create_lower_upper(self):
        group = getattr(self, "_lu_group", None)
        if group is not None:
            lower, upper, _ = group.factors()
            return lower[self._lu_index], upper[self._lu_index]

        def build():
            n = self.features
            eye = torch.eye(n, dtype=self.bias.dtype, device=self.bias.device)
            lower = eye.reshape(-1).index_put((self._lower_pos,), self.lower_entries).view(n, n)
            upper = torch.diag(self.upper_diag).reshape(-1).index_put((self._upper_pos,), self.upper_entries).view(n, n)
            return lower, upper

        return self._dense_cache.get(self._lu_params(), build)

    def weight(self):
        group = getattr(self, "_lu_group", None)
        if group is not None:
            return group.products()[self._lu_index]
        lower, upper = self._create_lower_upper()
        return lower @ upper

    def weight_inverse(self):
        group = getattr(self, "_lu_group", None)
        if group is not None:
            return group.inverses()[self._lu_index]
        lower, upper = self._create_lower_upper()
        eye = torch.eye(self.features, dtype=lower.dtype, device=lower.device)
        inv_lower = torch.linalg.solve_triangular(lower, eye, upper=False, unitriangular=True)
        return torch.linalg.solve_triangular(upper, inv_lower, upper=True)

    # ---- fused small path ------------------------------------------------------------------
    def _can_fuse(self, x):
        return self.features <= MAX_FUSED_CHANNELS and x.dim() == 2

    def _affine_map(self, inverse):
        group = getattr(self, "_lu_group", None)
        if group is not None:
            from .fusion import member_views

            mat, bias, lad = member_views(group, "_members_inv" if inverse else "_members_fwd", group.affine_maps(inverse))
            return mat[self._lu_index], bias[self._lu_index], lad[self._lu_index]

        def build():
            if not inverse:
                return self.weight(), self.bias, self.logabsdet()
            w_inv = self.weight_inverse()
            return w_inv, -(w_inv @ self.bias), -self.logabsdet()

        return self._map_cache[bool(inverse)].get(self._lu_params() + (self.bias,), build)

    # ---- large path: library GEMMs on cached dense triangles -------------------------------
    def _apply_transform(self, inputs, context, inverse):
        require_cuda(inputs)
        if inputs.dim() != 2 or inputs.shape[1] != self.features:
            raise ValueError("Expected inputs of shape [N, {}], got {}".format(self.features, tuple(inputs.shape)))
        if self._can_fuse(inputs):
            from .fusion import apply_affine

            return apply_affine([self._affine_map(inverse)], inputs)
        lower, upper = self._create_lower_upper()
        wide_tc = (self.features >= 1024 and config.conv_tensor_cores and config.conditioner_math == "fp32")
        if wide_tc:
            return self._apply_wide(inputs, lower, upper, inverse)
        if not inverse:  # lu.py:56-73
            outputs = F.linear(F.linear(inputs, upper), lower, self.bias)
            return outputs, self.logabsdet()
        outputs = (inputs - self.bias).t()  # lu.py:75-95
        if self.features >= 1024:   # wide layers: blocked substitution on GEMMs (b200/ops.py)
            outputs = ops.tri_solve(lower, outputs, upper=False, unit=True)
            outputs = ops.tri_solve(upper, outputs, upper=True)
        else:
            outputs = torch.linalg.solve_triangular(lower, outputs, upper=False, unitriangular=True)
            outputs = torch.linalg.solve_triangular(upper, outputs, upper=True)
        return outputs.t(), -self.logabsdet()

    # ---- wide layers on the tensor-core kernels ---------------------------------------------------
    def _param_version(self):
        return tuple(p._version for p in self._lu_params())

    def _apply_wide(self, inputs, lower, upper, inverse):
        """LULinear(3840) of the image post-processing (architectures/image_transforms.py:240-258) on the tcgen05 kernels:
        with the batch rows laid out as pixels, y = L (U x) + b is two 1x1 convolutions (lu.py:56-73); their data and weight
        gradients run through the same kernels.  The inverse (lu.py:75-95) is two triangular solves: while the parameters
        change every step they run as blocked substitutions; once the same parameters are seen again (evaluation,
        sampling, the phases of the sequential / alternating schedules that freeze this flow) the triangles are inverted
        once and the solves become two more 1x1 convolutions."""
        n = self.features
        ver = self._param_version()
        if not inverse:
            planes, meta = ops.rows_to_planes(inputs)
            t = ops.wide_linear(planes, upper, None, self._dense_cache, ("U", ver))
            t = ops.wide_linear(t, lower, self.bias, self._dense_cache, ("L", ver))
            return ops.planes_to_rows(t, meta), self.logabsdet()
        inv = self._wide_inverses(lower, upper, ver)
        if inv is None:
            outputs = (inputs - self.bias).t()
            outputs = ops.tri_solve(lower, outputs, upper=False, unit=True)
            outputs = ops.tri_solve(upper, outputs, upper=True)
            return outputs.t(), -self.logabsdet()
        inv_lower, inv_upper = inv
        planes, meta = ops.rows_to_planes(inputs - self.bias)
        t = ops.wide_solve(planes, lower, inv_lower, False, self._dense_cache, ("Linv", ver))
        t = ops.wide_solve(t, upper, inv_upper, True, self._dense_cache, ("Uinv", ver))
        return ops.planes_to_rows(t, meta), -self.logabsdet()

    def _wide_inverses(self, lower, upper, ver):
        """(L^-1, U^-1) for the current parameters, or None while they are not worth computing: the first call with a new
        parameter version answers None (substitution), the second one inverts and caches."""
        state = self.__dict__.setdefault("_wide_inv_state", {"seen": None, "ver": None, "inv": None})
        key = (ver, config.cache_epoch, str(lower.device))
        if state["ver"] == key:
            return state["inv"]
        if state["seen"] != key:
            state["seen"] = key
            return None
        with torch.no_grad():
            eye = torch.eye(self.features, dtype=lower.dtype, device=lower.device)
            inv_lower = torch.linalg.solve_triangular(lower.detach(), eye, upper=False, unitriangular=True)
            inv_upper = torch.linalg.solve_triangular(upper.detach(), eye, upper=True)
        state["ver"], state["inv"] = key, (inv_lower, inv_upper)
        return state["inv"]

    # reference spelling of the two directions
    def forward_no_cache(self, inputs, full_jacobian=False):
        return self.forward(inputs, full_jacobian=full_jacobian)

    def inverse_no_cache(self, inputs, full_jacobian=False):
        return self.inverse(inputs, full_jacobian=full_jacobian)


class LUGroup:
    """Dense factors, products and inverses of SEVERAL LU layers of the same width in one batched pass.

    A level of the image flow holds six 1x1 convolutions (twelve evaluations per step: forward and inverse), the
    inner flow eight ``LULinear(64)``: built one by one that is ~35 tiny launches per layer and direction, i.e.
    thousands of 2.5 us kernels per step.  Stacking the packed parameters lets the same torch ops (index_put,
    softplus, bmm, batched triangular solves) produce all members' matrices at once; members read their slice
    (a view).  Gradients flow back through the stack to every member's own parameters."""

    def __init__(self, layers):
        self.layers = list(layers)
        self._factors = VersionedCache()
        self._products = VersionedCache()
        self._inverses = VersionedCache()
        self._maps = {False: VersionedCache(), True: VersionedCache()}
        self._inv_perm = None

    def _params(self):
        return tuple(p for layer in self.layers for p in layer._lu_params())

    def factors(self):
        def build():
            first = self.layers[0]
            n, g = first.features, len(self.layers)
            lower_entries = torch.stack([layer.lower_entries for layer in self.layers])
            upper_entries = torch.stack([layer.upper_entries for layer in self.layers])
            upper_diag = F.softplus(torch.stack([layer.unconstrained_upper_diag for layer in self.layers])) + first.eps
            rows = torch.arange(g, device=lower_entries.device)[:, None]
            eye = torch.eye(n, dtype=lower_entries.dtype, device=lower_entries.device)
            lower = eye.expand(g, n, n).reshape(g, n * n).index_put((rows, first._lower_pos[None, :]), lower_entries).view(g, n, n)
            upper = torch.diag_embed(upper_diag).reshape(g, n * n).index_put((rows, first._upper_pos[None, :]), upper_entries).view(g, n, n)
            return lower, upper, torch.log(upper_diag).sum(dim=1)

        return self._factors.get(self._params(), build)

    def products(self):
        def build():
            lower, upper, _ = self.factors()
            return torch.bmm(lower, upper)

        return self._products.get(self._params(), build)

    def inverses(self):
        def build():
            lower, upper, _ = self.factors()
            n = lower.shape[-1]
            eye = torch.eye(n, dtype=lower.dtype, device=lower.device).expand_as(lower)
            inv_lower = torch.linalg.solve_triangular(lower, eye, upper=False, unitriangular=True)
            return torch.linalg.solve_triangular(upper, inv_lower, upper=True)

        return self._inverses.get(self._params(), build)

    def _inverse_permutations(self):
        """[G, n] inverse channel permutations of the members (1x1 convolutions), or None for plain LULinear."""
        first = self.layers[0]
        if getattr(first, "permutation", None) is None:
            return None
        device = first.lower_entries.device
        key = (str(device),) + tuple((layer.permutation._permutation.data_ptr(), layer.permutation._permutation._version)
                                     for layer in self.layers)
        if self._inv_perm is None or self._inv_perm[0] != key:   # load_state_dict rewrites the permutation buffers
            self._inv_perm = (key, torch.stack([layer.permutation._inverse_permutation for layer in self.layers]).to(device))
        return self._inv_perm[1]

    def affine_maps(self, inverse):
        """Stacked per-pixel affine maps (M [G,n,n], m [G,n], logabsdet [G]) of the members: y = W x[perm] + b forward
        (conv.py:33-41; no permutation for LULinear, lu.py:56-73) or x = (W^-1 (y - b))[inv_perm] (conv.py:43-55)."""

        def build():
            bias = torch.stack([layer.bias for layer in self.layers])
            lad = self.factors()[2]
            inv_perm = self._inverse_permutations()
            g, n = bias.shape
            if not inverse:
                mat = self.products()
                if inv_perm is not None:
                    mat = torch.gather(mat, 2, inv_perm[:, None, :].expand(g, n, n))      # W[:, inv_perm]
                return mat, bias, lad
            mat = self.inverses()
            if inv_perm is not None:
                mat = torch.gather(mat, 1, inv_perm[:, :, None].expand(g, n, n))          # W^-1[inv_perm, :]
            return mat, -torch.bmm(mat, bias[:, :, None]).squeeze(2), -lad

        perms = tuple(layer.permutation._permutation for layer in self.layers if getattr(layer, "permutation", None) is not None)
        return self._maps[bool(inverse)].get(self._params() + tuple(layer.bias for layer in self.layers) + perms, build)


def assign_lu_groups(model):
    """Group the fusable LU layers of ``model`` (same class, width, eps, device) so that their dense matrices are
    built together (``LUGroup``).  Idempotent; layers used outside a model-level pass keep their own path."""
    buckets = {}
    for module in model.modules():
        if isinstance(module, LULinear) and module.features <= MAX_FUSED_CHANNELS:
            key = (type(module), module.features, float(module.eps), str(module.lower_entries.device), module.lower_entries.dtype)
            buckets.setdefault(key, []).append(module)
    for layers in buckets.values():
        if len(layers) < 2:
            continue
        group = LUGroup(layers)
        for index, layer in enumerate(layers):
            object.__setattr__(layer, "_lu_group", group)
            object.__setattr__(layer, "_lu_index", index)
    from .normalization import ActNorm, ActNormGroup

    buckets = {}
    for module in model.modules():
        if isinstance(module, ActNorm) and module.log_scale.numel() <= MAX_FUSED_CHANNELS:
            key = (module.log_scale.numel(), str(module.log_scale.device), module.log_scale.dtype)
            buckets.setdefault(key, []).append(module)
    for layers in buckets.values():
        if len(layers) < 2:
            continue
        group = ActNormGroup(layers)
        for index, layer in enumerate(layers):
            object.__setattr__(layer, "_an_group", group)
            object.__setattr__(layer, "_an_index", index)
    # adjacent (ActNorm, 1x1 convolution) members of a chain, both grouped: compose all such pairs of a width together
    from .base import CompositeTransform
    from .conv import OneByOneConvolution
    from .fusion import PairGroup

    from .permutations import Permutation

    pair_buckets = {}
    for module in model.modules():
        if not isinstance(module, CompositeTransform):
            continue
        chain = list(module._transforms)
        for first, second in zip(chain[:-1], chain[1:]):
            if getattr(second, "_lu_group", None) is None or getattr(first, "_pair_group", None) is not None:
                continue
            if (isinstance(first, ActNorm) and isinstance(second, OneByOneConvolution) and getattr(first, "_an_group", None) is not None
                    and first.log_scale.numel() == second.features):
                pair_buckets.setdefault(("actnorm", first._an_group, second._lu_group), []).append((first, second))
            elif (isinstance(first, Permutation) and type(second) is LULinear and first._dim == 1
                  and first._permutation.numel() == second.features):
                pair_buckets.setdefault(("perm", None, second._lu_group), []).append((first, second))
    for (kind, _, _), pairs in pair_buckets.items():
        if len(pairs) < 2:
            continue
        group = PairGroup(pairs, kind)
        for index, (first, second) in enumerate(pairs):
            object.__setattr__(first, "_pair_group", group)
            object.__setattr__(first, "_pair_index", index)
            object.__setattr__(first, "_pair_partner", second)
