"""Autograd functions over the C ABI of ``libmflow_b200.so``.

Each function is one fused launch (forward) plus one or two fused launches (backward); inputs must be
float32 CUDA tensors.  There is no CPU path: a missing library or a CPU tensor raises, and so does an
unsupported shape - with one documented exception: the weight gradient of a 3x3 convolution with more than 112
output channels (no BASELINE configuration has one) is a cuDNN call (``_ConvTC.backward``).
"""
import contextlib
import ctypes
import math
import weakref

import torch

from . import cabi
from .config import config
from .cabi import ConvDesc, MixDesc, MlpDesc, RqsBoxDesc, RqsDesc, check, lib, ptr, require_cuda, stream


def _c(t):
    return t if t is None or t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------------------------
# (a) rational-quadratic spline
# --------------------------------------------------------------------------------------------
class SplineGeometry:
    """Where the transformed (and identity) features of a coupling layer live inside the full
    activation tensor, as an arithmetic progression: feature c at ``offset + c * step`` (in units
    of channels / features).  All masks the reference factories build are of this form
    (alternating: utils/various.py:410-421; mid split: :424-434)."""

    __slots__ = ("n_transform", "t_offset", "t_step", "n_identity", "i_offset", "i_step")

    def __init__(self, n_transform, t_offset, t_step, n_identity, i_offset, i_step):
        self.n_transform, self.t_offset, self.t_step = n_transform, t_offset, t_step
        self.n_identity, self.i_offset, self.i_step = n_identity, i_offset, i_step

    @staticmethod
    def from_indices(transform_idx, identity_idx):
        def prog(idx):
            idx = [int(v) for v in idx]
            if len(idx) == 0:
                return 0, 1
            step = idx[1] - idx[0] if len(idx) > 1 else 1
            if step < 1 or any(idx[j + 1] - idx[j] != step for j in range(len(idx) - 1)):
                return None
            return idx[0], step

        t, i = prog(transform_idx), prog(identity_idx)
        if t is None or i is None:
            return None
        return SplineGeometry(len(transform_idx), t[0], t[1], len(identity_idx), i[0], i[1])


ROWTILE_MAX_FEATURES = 92   # rows x planes spline layout: the kernel stages [128, D] row tiles in 48 KB of shared memory


def _rqs_desc(x, params, geom, num_bins, tail_bound, min_w, min_h, min_d, divisor, inverse):
    """Descriptor for a coupling layer on the full activation tensor ``x`` ([B, C, H, W] or [B, D],
    contiguous) and the conditioner output ``params`` ([B, C_t*P, H, W] or [B, C_t*P], contiguous - or, for vectors whose
    conditioner ran on the convolution kernels, its NCHW output [B / hw, C_t*P, h, w]: "rows x planes")."""
    p = 3 * num_bins - 1
    d = RqsDesc()
    b = x.shape[0]
    feats = x.shape[1]
    inner = x.shape[2] * x.shape[3] if x.dim() == 4 else 1
    d.n_outer, d.n_chan, d.n_inner = b, geom.n_transform, inner
    d.x_so = d.y_so = feats * inner
    d.x_sc = d.y_sc = geom.t_step * inner
    d.p_so = geom.n_transform * p * inner
    d.p_sc = p * inner
    d.p_sp = inner
    d.p_si = 1
    d.num_bins, d.inverse = num_bins, int(bool(inverse))
    d.tail_bound, d.min_bin_width, d.min_bin_height, d.min_derivative = tail_bound, min_w, min_h, min_d
    d.wh_divisor = divisor
    d.n_id_chan, d.id_off, d.id_sc = geom.n_identity, geom.i_offset * inner, geom.i_step * inner
    if params is not None and x.dim() == 2 and params.dim() == 4:   # rows x planes: unshifted row pointers, the kernel stages whole rows
        d.p_hw, d.t_off = params.shape[2] * params.shape[3], geom.t_offset
        return d, 0
    d.p_hw = d.t_off = 0
    return d, geom.t_offset * inner


def _rqs_shape(d):
    return f"[{d.n_outer}x{d.n_chan}x{d.n_inner},K={d.num_bins}{',planes' if d.p_hw else ''}]"


def _rqs_bytes(d, backward):
    """Algorithmic HBM bytes of one spline launch: forward / inverse read P params + x
    and write y per element (+ 4 B per sample for the log-det); backward reads params, x, g_y and
    writes g_params, g_x."""
    p = 3 * d.num_bins - 1
    n = d.n_outer * d.n_chan * d.n_inner
    return n * 4 * (2 * p + 3) if backward else n * 4 * (p + 2) + 4 * d.n_outer


class _RqsCoupling(torch.autograd.Function):
    """(x_full, params) -> (y_full, logabsdet[B]).  coupling.py:58-185 + rational_quadratic.py."""

    @staticmethod
    def forward(ctx, x, params, geom, num_bins, tail_bound, min_w, min_h, min_d, divisor, inverse):
        require_cuda(x, params)
        x, params = _c(x), _c(params)
        p = 3 * num_bins - 1
        expected = x.shape[0] * geom.n_transform * p * (x.shape[2] * x.shape[3] if x.dim() == 4 else 1)
        if params.numel() != expected:
            raise ValueError(f"conditioner output has {params.numel()} values, expected {expected}")
        if x.dim() == 2 and params.dim() == 4:   # rows x planes: [B / hw, C_t*P, h, w] with hw a power of two dividing B
            hw = params.shape[2] * params.shape[3]
            if params.shape[1] != geom.n_transform * p or hw & (hw - 1) or x.shape[0] % hw or x.shape[1] > ROWTILE_MAX_FEATURES:
                raise ValueError(f"planes-layout conditioner output {tuple(params.shape)} does not match {x.shape[0]} rows x {geom.n_transform} features")
        desc, off = _rqs_desc(x, params, geom, num_bins, tail_bound, min_w, min_h, min_d, divisor, inverse)
        y = torch.empty_like(x)
        lad = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        ws = cabi.workspace("rqs", lib().mflow_rqs_workspace_bytes(ctypes.byref(desc)), x.device)
        xo = ctypes.c_void_p(x.data_ptr() + 4 * off)
        yo = ctypes.c_void_p(y.data_ptr() + 4 * off)
        # identity features are addressed relative to the same shifted base
        desc.id_off -= off
        tok = cabi.kernel_timer.begin(("rqs_inv" if inverse else "rqs_fwd") + _rqs_shape(desc), _rqs_bytes(desc, False)) if cabi.kernel_timer else None
        check(lib().mflow_rqs_apply(ctypes.byref(desc), xo, ptr(params), yo, None, ptr(lad), None, None, ptr(ws), stream()),
              "mflow_rqs_apply")
        if tok:
            cabi.kernel_timer.end(tok)
        cabi.count_launch()
        ctx.desc, ctx.off = desc, off
        ctx.save_for_backward(x, params, y if inverse else None)
        return y, lad

    @staticmethod
    def backward(ctx, g_y, g_lad):
        x, params, y = ctx.saved_tensors
        desc, off = ctx.desc, ctx.off
        g_y = None if g_y is None else _c(g_y)
        g_lad = None if g_lad is None else _c(g_lad)
        g_x = torch.empty_like(x)
        g_p = torch.empty_like(params)
        sh = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr() + 4 * off)
        tok = cabi.kernel_timer.begin("rqs_bwd" + _rqs_shape(desc), _rqs_bytes(desc, True)) if cabi.kernel_timer else None
        # NCHW parameters came from a tensor-core convolution, whose backward wants max |g_p| for its operand scale: the
        # kernel publishes it on the side (one absmax pass over the largest tensor of the step less per coupling)
        slot = _amax_pool.take(x.device) if (params.dim() == 4 and config.conv_operand_format == "fp16") else None
        check(lib().mflow_rqs_backward(ctypes.byref(desc), sh(x), sh(y), ptr(params), sh(g_y), ptr(g_lad), None, sh(g_x),
                                       ptr(g_p), ptr(slot), stream()), "mflow_rqs_backward")
        if tok:
            cabi.kernel_timer.end(tok)
        cabi.count_launch()
        if slot is not None:
            _publish_amax(g_p, slot)
        return (g_x, g_p) + (None,) * 8


def rqs_coupling(x, params, geom, num_bins, tail_bound, min_w, min_h, min_d, divisor, inverse):
    return _RqsCoupling.apply(x, params, geom, num_bins, float(tail_bound), float(min_w), float(min_h), float(min_d),
                              float(divisor), bool(inverse))


class _RqsElementwise(torch.autograd.Function):
    """Flat [N] inputs with [N, 3K-1] parameter rows -> (outputs[N], logabsdet[N]): the function-level
    API ``unconstrained_rational_quadratic_spline`` (rational_quadratic.py:16-64)."""

    @staticmethod
    def forward(ctx, x, params, num_bins, tail_bound, min_w, min_h, min_d, inverse, want_bins):
        require_cuda(x, params)
        x, params = _c(x), _c(params)
        n, p = x.numel(), 3 * num_bins - 1
        d = RqsDesc()
        d.n_outer, d.n_chan, d.n_inner = n, 1, 1
        d.x_so = d.y_so = 1
        d.x_sc = d.y_sc = 1
        d.p_so, d.p_sc, d.p_sp, d.p_si = p, p, 1, 1
        d.num_bins, d.inverse = num_bins, int(bool(inverse))
        d.tail_bound, d.min_bin_width, d.min_bin_height, d.min_derivative, d.wh_divisor = tail_bound, min_w, min_h, min_d, 1.0
        y = torch.empty_like(x)
        lad = torch.empty_like(x)
        bins = torch.empty(n, dtype=torch.int32, device=x.device) if want_bins else None
        check(lib().mflow_rqs_apply(ctypes.byref(d), ptr(x), ptr(params), ptr(y), ptr(lad), None, None, ptr(bins), None, stream()),
              "mflow_rqs_apply")
        cabi.count_launch()
        ctx.desc = d
        ctx.save_for_backward(x, params, y if inverse else None)
        if want_bins:
            ctx.mark_non_differentiable(bins)
            return y, lad, bins
        return y, lad

    @staticmethod
    def backward(ctx, g_y, g_lad, *_):
        x, params, y = ctx.saved_tensors
        g_x, g_p = torch.empty_like(x), torch.empty_like(params)
        check(lib().mflow_rqs_backward(ctypes.byref(ctx.desc), ptr(x), ptr(y), ptr(params), ptr(None if g_y is None else _c(g_y)),
                                       None, ptr(None if g_lad is None else _c(g_lad)), ptr(g_x), ptr(g_p), None, stream()),
              "mflow_rqs_backward")
        cabi.count_launch()
        return (g_x, g_p) + (None,) * 7


def rqs_elementwise(x, params, num_bins, tail_bound, min_w=1e-3, min_h=1e-3, min_d=1e-3, inverse=False, return_bins=False):
    return _RqsElementwise.apply(x, params, num_bins, float(tail_bound), float(min_w), float(min_h), float(min_d), bool(inverse),
                                 bool(return_bins))


class _RqsBox(torch.autograd.Function):
    """Flat [N] inputs with [N, 3K+1] parameter rows on a box [left, right] x [bottom, top]:
    ``rational_quadratic_spline`` (rational_quadratic.py:67-168)."""

    @staticmethod
    def forward(ctx, x, params, num_bins, box, mins, inverse, want_bins):
        require_cuda(x, params)
        x, params = _c(x), _c(params)
        d = RqsBoxDesc(n=x.numel(), num_bins=num_bins, inverse=int(bool(inverse)), left=box[0], right=box[1], bottom=box[2], top=box[3],
                       min_bin_width=mins[0], min_bin_height=mins[1], min_derivative=mins[2])
        y, lad = torch.empty_like(x), torch.empty_like(x)
        bins = torch.empty(x.numel(), dtype=torch.int32, device=x.device) if want_bins else None
        check(lib().mflow_rqs_box_apply(ctypes.byref(d), ptr(x), ptr(params), ptr(y), ptr(lad), ptr(bins), stream()), "mflow_rqs_box_apply")
        cabi.count_launch()
        ctx.desc = d
        ctx.save_for_backward(x, params, y if inverse else None)
        if want_bins:
            ctx.mark_non_differentiable(bins)
            return y, lad, bins
        return y, lad

    @staticmethod
    def backward(ctx, g_y, g_lad, *_):
        x, params, y = ctx.saved_tensors
        g_x, g_p = torch.empty_like(x), torch.empty_like(params)
        check(lib().mflow_rqs_box_backward(ctypes.byref(ctx.desc), ptr(x), ptr(y), ptr(params), ptr(None if g_y is None else _c(g_y)),
                                           ptr(None if g_lad is None else _c(g_lad)), ptr(g_x), ptr(g_p), stream()),
              "mflow_rqs_box_backward")
        cabi.count_launch()
        return (g_x, g_p) + (None,) * 5


def rqs_box(x, params, num_bins, left, right, bottom, top, min_w=1e-3, min_h=1e-3, min_d=1e-3, inverse=False, return_bins=False):
    return _RqsBox.apply(x, params, int(num_bins), (float(left), float(right), float(bottom), float(top)),
                         (float(min_w), float(min_h), float(min_d)), bool(inverse), bool(return_bins))


def rqs_search(knots, values):
    """The kernels' bin search on given knots ([n, K+1], [n]) -> int32 [n] (test entry)."""
    require_cuda(knots, values)
    knots, values = _c(knots), _c(values)
    bins = torch.empty(values.numel(), dtype=torch.int32, device=values.device)
    check(lib().mflow_rqs_search(ptr(knots), ptr(values), ptr(bins), values.numel(), knots.shape[-1] - 1, stream()), "mflow_rqs_search")
    cabi.count_launch()
    return bins


# --------------------------------------------------------------------------------------------
# (b) conditioner convolutions on tcgen05 tensor cores (3xTF32)
# --------------------------------------------------------------------------------------------
TC_WIDTHS = (4, 8, 16, 32)


def conv_tc_supported(x, weight):
    return (x.dim() == 4 and x.shape[2] == x.shape[3] and x.shape[3] in TC_WIDTHS and weight.shape[2] == weight.shape[3]
            and weight.shape[2] in (1, 3))


def _split_weight(w_taps):
    """[taps, N, K] fp32 -> (hi, lo), zero padded to N % 128 == 0 and K % 32 == 0: hi = TF32 truncation
    of w, lo = (w - hi) rounded to TF32, so that hi + lo reproduces w to ~2^-22."""
    taps, n, k = w_taps.shape
    n_pad, k_pad = (n + 127) // 128 * 128, (k + 31) // 32 * 32
    wt = torch.zeros(taps, n_pad, k_pad, dtype=torch.float32, device=w_taps.device)
    wt[:, :n, :k] = w_taps
    hi = (wt.view(torch.int32) & -8192).view(torch.float32)
    lo = (((wt - hi).view(torch.int32) + 4096) & -8192).view(torch.float32)
    return hi.contiguous(), lo.contiguous(), n_pad, k_pad


def _split_weight_f16(w_taps):
    """[taps, N, K] fp32 -> (hi, lo, unscale): fp16 halves of the weights scaled per output row n by the power of two that
    puts the row's largest magnitude into [2^13, 2^14) (fp16 keeps 11 significant bits only inside its 5-bit exponent
    range), zero padded to N % 128 == 0 and K % 32 == 0; unscale[n] = 2^-s_n is applied to the accumulator by the kernel's
    epilogue.  hi + lo reproduces the scaled weight to ~2^-22 of the row maximum."""
    taps, n, k = w_taps.shape
    n_pad, k_pad = (n + 127) // 128 * 128, (k + 31) // 32 * 32
    amax = w_taps.abs().amax(dim=(0, 2))
    _, exp = torch.frexp(amax)                       # amax = m * 2^exp with m in [0.5, 1): floor(log2 amax) = exp - 1
    shift = torch.where(amax > 0, 14 - exp, torch.zeros_like(exp)).clamp(-100, 100).to(torch.float32)
    wt = torch.zeros(taps, n_pad, k_pad, dtype=torch.float32, device=w_taps.device)
    wt[:, :n, :k] = w_taps * torch.exp2(shift)[None, :, None]
    hi = wt.to(torch.float16)
    lo = (wt - hi.to(torch.float32)).to(torch.float16)
    unscale = torch.ones(n_pad, dtype=torch.float32, device=w_taps.device)
    unscale[:n] = torch.exp2(-shift)
    return hi.contiguous(), lo.contiguous(), unscale, n_pad, k_pad


class _AmaxPool:
    """Zero-initialised device scalars for the kernels' published absolute maxima.  One ``torch.zeros`` launch provides
    1024 of them (a launch per scalar was 700 fill kernels per step); a slot is handed out once and never reused, the
    views keep their block alive for as long as a consumer (a saved backward context) holds one."""

    def __init__(self):
        self._block, self._next, self._key = None, 0, None

    def take(self, device):
        key = (device, torch.cuda.current_stream(device).cuda_stream, torch.cuda.is_current_stream_capturing())
        if self._block is None or self._next >= self._block.numel() or self._key != key:
            self._block, self._next, self._key = torch.zeros(1024, dtype=torch.float32, device=device), 0, key
        self._next += 1
        return self._block[self._next - 1 : self._next]

    def reset(self):
        self._block = None


_amax_pool = _AmaxPool()


def tensor_amax(x, fresh=False):
    """Device scalar holding max |x|: the operand scale of the fp16-split kernels.  Kernels of this library publish it with
    their output (attribute ``_mflow_amax`` of the tensor they return: valid for exactly that tensor object at that
    version); for any other tensor one ``mflow_absmax`` pass computes it (and remembers it on the tensor)."""
    tag = getattr(x, "_mflow_amax", None)
    if tag is not None and not fresh and tag[1] == x._version and tag[2] == x.data_ptr():
        return tag[0]
    out = _amax_pool.take(x.device)
    check(lib().mflow_absmax(ptr(x), x.numel(), ptr(out), stream()), "mflow_absmax")
    cabi.count_launch()
    x._mflow_amax = (out, x._version, x.data_ptr())
    return out


def _publish_amax(y, slot):
    y._mflow_amax = (slot, y._version, y.data_ptr())


class _WeightCache:
    """Split / padded operand copies of a convolution weight, rebuilt when the parameter changes.
    Entries are bound to the tensor object itself (weak reference), not just to its id(): Python reuses
    ids, and a new tensor can land on the same address with the same version counter."""

    def __init__(self):
        self._entries = {}

    def get(self, weight, transposed, owner=None, fmt="tf32", version=None):
        """owner: the long-lived tensor (parameter) that `weight` is a fresh view of, if any - the entry is then
        bound to the owner, so that a new view object per call does not miss.
        version: for weights DERIVED from parameters on every pass (the dense L / U of a wide LULinear): ``owner`` is then
        any long-lived object (the module, plus a tag) and ``version`` a hashable built from the parameters' version
        counters - the fresh tensor's address does not take part."""
        anchor = weight if owner is None else owner
        key = (id(anchor), transposed, fmt) if version is None else (id(anchor), transposed, fmt, version[0])
        ver = ((weight.data_ptr(), anchor._version, tuple(weight.shape), config.cache_epoch) if version is None
               else (version, tuple(weight.shape), config.cache_epoch))
        ent = self._entries.get(key)
        if ent is None or ent[0]() is not anchor or ent[1] != ver:
            if len(self._entries) > 4096:
                self._entries = {k: v for k, v in self._entries.items() if v[0]() is not None}
            with torch.no_grad():
                co, ci, kh, kw = weight.shape
                if not transposed:  # forward: W[tap][co][ci]
                    taps = weight.permute(2, 3, 0, 1).reshape(kh * kw, co, ci)
                else:  # data gradient: correlation with the taps flipped and channels swapped
                    taps = weight.flip(2, 3).permute(2, 3, 1, 0).reshape(kh * kw, ci, co)
                ent = (weakref.ref(anchor), ver, _split_weight_f16(taps) if fmt == "fp16" else _split_weight(taps))
            self._entries[key] = ent
        return ent[2]


_weight_cache = _WeightCache()


def _conv_format():
    fmt = config.conv_operand_format
    if fmt not in ("fp16", "tf32"):
        raise ValueError("config.conv_operand_format must be 'fp16' or 'tf32', got {!r}".format(fmt))
    return fmt


def _conv_tc_launch(x, split, c_out, kernel, bias, residual, gate, relu_in, x_amax=None):
    """One convolution launch.  fp16 format: ``x_amax`` is the device scalar max |x| (``tensor_amax``); the launch publishes
    max |y| on the returned tensor for the next consumer."""
    f16 = len(split) == 5
    b, c_in, h, w = x.shape
    n_pad, k_pad = split[-2], split[-1]
    d = ConvDesc(batch=b, c_in=c_in, c_out=c_out, height=h, width=w, kernel=kernel, relu_in=int(relu_in), relu_out=0,
                 n_pad=n_pad, k_pad=k_pad)
    y = torch.empty((b, c_out, h, w), dtype=torch.float32, device=x.device)
    out_amax = _amax_pool.take(x.device) if f16 else None
    tok = None
    if cabi.kernel_timer:  # useful (fp32-equivalent) FLOPs: 2 * pixels * c_in * c_out * taps - NOT the three tensor passes
        tok = cabi.kernel_timer.begin(f"conv{kernel}x{kernel}[{b}x{c_in}->{c_out}x{h}x{w}]", 2 * b * h * w * c_in * c_out * kernel * kernel)
    if f16:
        hi, lo, unscale = split[0], split[1], split[2]
        check(lib().mflow_conv2d_f16x3(ctypes.byref(d), ptr(x), ptr(hi), ptr(lo), ptr(unscale), ptr(x_amax), ptr(bias), ptr(residual),
                                       ptr(gate), ptr(y), ptr(out_amax), stream()), "mflow_conv2d_f16x3")
    else:
        hi, lo = split[0], split[1]
        check(lib().mflow_conv2d_tf32x3(ctypes.byref(d), ptr(x), ptr(hi), ptr(lo), ptr(bias), ptr(residual), ptr(gate), ptr(y), stream()),
              "mflow_conv2d_tf32x3")
    if tok:
        cabi.kernel_timer.end(tok)
    cabi.count_launch()
    if f16:
        _publish_amax(y, out_amax)
    return y


def conv_rqs_supported(hidden, weight, x_full, num_bins):
    """Can the final 1x1 convolution of a conditioner run with the coupling layer's spline in its epilogue?"""
    return (config.fuse_spline_epilogue and config.conv_tensor_cores and config.conditioner_math == "fp32"
            and config.conv_operand_format == "fp16" and num_bins == 11 and weight.shape[2] == 1 and weight.shape[0] % 32 == 0
            and conv_tc_supported(hidden, weight) and x_full.dim() == 4 and not torch.is_grad_enabled())


def conv_rqs(hidden, weight, bias, x_full, geom, num_bins, tail_bound, min_w, min_h, min_d, divisor, inverse):
    """(y_full, logabsdet[B]) of an image coupling layer whose conditioner ends in ``conv2d(hidden, weight, bias)`` (1x1):
    one launch computes the convolution and evaluates the spline on its accumulator rows (no autograd: evaluation only)."""
    require_cuda(hidden, weight, bias, x_full)
    hidden, x_full = _c(hidden), _c(x_full)
    b, c_in, h, w = hidden.shape
    c_out = weight.shape[0]
    split = _weight_cache.get(weight, False, None, "fp16")
    hi, lo, unscale, n_pad, k_pad = split
    d = ConvDesc(batch=b, c_in=c_in, c_out=c_out, height=h, width=w, kernel=1, relu_in=0, relu_out=0, n_pad=n_pad, k_pad=k_pad)
    sd, off = _rqs_desc(x_full, None, geom, num_bins, tail_bound, min_w, min_h, min_d, divisor, inverse)
    sd.id_off -= off
    y = torch.empty_like(x_full)
    lad = torch.empty(b, dtype=torch.float32, device=x_full.device)
    ws = cabi.workspace("conv_rqs", lib().mflow_conv2d_rqs_workspace_bytes(ctypes.byref(d)), x_full.device)
    xo = ctypes.c_void_p(x_full.data_ptr() + 4 * off)
    yo = ctypes.c_void_p(y.data_ptr() + 4 * off)
    tok = None
    if cabi.kernel_timer:
        tok = cabi.kernel_timer.begin(f"conv1x1+rqs[{b}x{c_in}->{c_out}x{h}x{w}]", 2 * b * h * w * c_in * c_out)
    check(lib().mflow_conv2d_rqs_f16x3(ctypes.byref(d), ptr(hidden), ptr(hi), ptr(lo), ptr(unscale), ptr(tensor_amax(hidden)), ptr(bias),
                                       ctypes.byref(sd), xo, yo, ptr(lad), ptr(ws), stream()), "mflow_conv2d_rqs_f16x3")
    if tok:
        cabi.kernel_timer.end(tok)
    cabi.count_launch(2)
    return y, lad


def conv_wgrad_supported(x, c_out, kernel):
    b, c_in, h, w = x.shape
    return h == w and w in (4, 8, 16, 32) and w >= config.conv_wgrad_min_width and b >= 1 and (kernel == 1 or c_out <= 112)


def conv_wgrad_tc(g, x, kernel, relu_in, with_bias=False, x_amax=None, g_amax=None):
    """g_w[co, ci, ky, kx] of y = conv2d(relu?(x), W): pixels-as-K GEMM on the tensor cores (error-compensated split of
    both operands, fp16 or TF32 format), deterministic.
    with_bias: also return the bias gradient sum_{b,h,w} g (computed inside the same kernel) as (g_w, g_b)."""
    require_cuda(g, x)
    g, x = _c(g), _c(x)
    f16 = _conv_format() == "fp16"
    if f16:
        x_amax = tensor_amax(x) if x_amax is None else x_amax
        g_amax = tensor_amax(g) if g_amax is None else g_amax
    b, c_in, h, w = x.shape
    c_out = g.shape[1]
    d = ConvDesc(batch=b, c_in=c_in, c_out=c_out, height=h, width=w, kernel=kernel, relu_in=int(relu_in), relu_out=0, n_pad=0, k_pad=0)
    gw = torch.empty((c_out, c_in, kernel, kernel), dtype=torch.float32, device=x.device)
    gb = torch.empty(c_out, dtype=torch.float32, device=x.device) if with_bias else None
    ws = cabi.workspace("conv_wgrad", lib().mflow_conv2d_wgrad_workspace_bytes(ctypes.byref(d)), x.device)
    tok = None
    if cabi.kernel_timer:
        tok = cabi.kernel_timer.begin(f"wgrad{kernel}x{kernel}[{b}x{c_in}->{c_out}x{h}x{w}]", 2 * b * h * w * c_in * c_out * kernel * kernel)
    if f16:
        check(lib().mflow_conv2d_wgrad_f16x3(ctypes.byref(d), ptr(x), ptr(g), ptr(x_amax), ptr(g_amax), ptr(gw), ptr(gb), ptr(ws), stream()),
              "mflow_conv2d_wgrad_f16x3")
    else:
        check(lib().mflow_conv2d_wgrad_tf32x3(ctypes.byref(d), ptr(x), ptr(g), ptr(gw), ptr(gb), ptr(ws), stream()), "mflow_conv2d_wgrad_tf32x3")
    if tok:
        cabi.kernel_timer.end(tok)
    cabi.count_launch(2)
    return (gw, gb) if with_bias else gw


class _GradScope:
    """``inputs_only``: backward functions skip parameter gradients.  ``ctx.needs_input_grad`` says whether a parameter
    requires grad, not whether this backward call was asked for it: ``torch.autograd.grad(y, x)`` of the full-Jacobian
    passes (transforms/base.py:batch_jacobian) would otherwise run every weight-gradient kernel D times for nothing."""
    inputs_only = False


@contextlib.contextmanager
def input_grads_only():
    prev, _GradScope.inputs_only = _GradScope.inputs_only, True
    try:
        yield
    finally:
        _GradScope.inputs_only = prev


class _ConvTC(torch.autograd.Function):
    """y = conv2d(relu?(x), W) + b (+ residual), 3x3 pad 1 or 1x1, NCHW fp32, on the tensor cores.
    Backward: data gradient through the same kernel (transposed / flipped weights, ReLU mask fused in the
    epilogue); weight gradient as a library call (cuDNN) for now; bias gradient a reduction."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, relu_in, owner=None, version=None):
        require_cuda(x, weight, bias, residual)
        x = _c(x)
        residual = None if residual is None else _c(residual)
        kernel = weight.shape[2]
        fmt = _conv_format()
        x_amax = tensor_amax(x) if fmt == "fp16" else None
        y = _conv_tc_launch(x, _weight_cache.get(weight, False, owner, fmt, version), weight.shape[0], kernel, bias, residual, None, relu_in, x_amax)
        ctx.owner, ctx.fmt, ctx.x_amax, ctx.version = owner, fmt, x_amax, version
        ctx.save_for_backward(x, weight)
        ctx.relu_in, ctx.has_bias, ctx.has_res = relu_in, bias is not None, residual is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g = _c(g)
        gx = gw = gb = None
        kernel = weight.shape[2]
        g_amax = tensor_amax(g) if ctx.fmt == "fp16" else None   # one scale for the data and the weight gradient
        if ctx.needs_input_grad[0]:
            gx = _conv_tc_launch(g, _weight_cache.get(weight, True, ctx.owner, ctx.fmt, ctx.version), weight.shape[1], kernel, None, None,
                                 x if ctx.relu_in else None, False, g_amax)
        want_w = ctx.needs_input_grad[1] and not _GradScope.inputs_only
        want_bias = ctx.has_bias and ctx.needs_input_grad[2] and not _GradScope.inputs_only
        if want_w and config.conv_wgrad_tensor_cores and conv_wgrad_supported(x, weight.shape[0], kernel):
            if want_bias:   # the weight-gradient kernel sums g over the pixels on the side
                gw, gb = conv_wgrad_tc(g, x, kernel, ctx.relu_in, with_bias=True, x_amax=ctx.x_amax, g_amax=g_amax)
            else:
                gw = conv_wgrad_tc(g, x, kernel, ctx.relu_in, x_amax=ctx.x_amax, g_amax=g_amax)
        elif want_w:   # shapes the tensor-core kernel does not take: library call
            xin = torch.relu(x) if ctx.relu_in else x
            gw = torch.ops.aten.convolution_backward(g, xin, weight, None, [1, 1], [kernel // 2] * 2, [1, 1], False, [0, 0], 1,
                                                     [False, True, False])[1]
        if want_bias and gb is None:
            gb = channel_sum(g)
        return gx, gw, gb, (g if ctx.has_res and ctx.needs_input_grad[3] else None), None, None, None


# --------------------------------------------------------------------------------------------
# (b') vector conditioner: the whole ResidualNet in one launch (nn/resnet.py:9-72)
# --------------------------------------------------------------------------------------------
def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def mlp_fused_supported(net, inputs, context):
    n_ctx = 0 if context is None else context.shape[1]
    return (inputs.dim() == 2 and net.hidden_features <= 128 and inputs.shape[1] + n_ctx <= 128 and len(net.blocks) <= 4
            and net.final_layer.out_features <= 8192 and (context is None or (context.dim() == 2 and not context.requires_grad))
            and inputs.stride(1) >= 1 and inputs.stride(0) >= 1)


class _MlpFused(torch.autograd.Function):
    """(x view, ctx, weights..., biases...) -> conditioner output [rows, n_out].  x may be a strided 2-D view (the identity
    features of a coupling layer): it is read in place."""

    @staticmethod
    def forward(ctx, x, context, n_layers, n_blocks, *params):
        weights, biases = params[:n_layers], params[n_layers:]
        require_cuda(x, context, *params)
        context = None if context is None else _c(context)
        weights, biases = [_c(w) for w in weights], [_c(b) for b in biases]
        rows, n_in = x.shape
        n_ctx = 0 if context is None else context.shape[1]
        d = MlpDesc(rows=rows, n_in=n_in, n_ctx=n_ctx, hidden=weights[0].shape[0], n_out=weights[-1].shape[0], n_blocks=n_blocks)
        out = torch.empty((rows, d.n_out), dtype=torch.float32, device=x.device)
        check(lib().mflow_mlp_forward(ctypes.byref(d), ptr(x), x.stride(0), x.stride(1), ptr(context), n_ctx, _ptr_array(weights),
                                      _ptr_array(biases), ptr(out), stream()), "mflow_mlp_forward")
        cabi.count_launch()
        ctx.desc, ctx.n_layers = d, n_layers
        ctx.save_for_backward(x, context, *weights, *biases)
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, context, *params = ctx.saved_tensors
        n_layers, d = ctx.n_layers, ctx.desc
        weights, biases = params[:n_layers], params[n_layers:]
        g_out = _c(g_out)
        g_x = torch.empty((d.rows, d.n_in), dtype=torch.float32, device=x.device) if ctx.needs_input_grad[0] else None
        g_w = [torch.empty_like(w) for w in weights]
        g_b = [torch.empty_like(b) for b in biases]
        ws = cabi.workspace("mlp", 4 * lib().mflow_mlp_workspace_floats(ctypes.byref(d)), x.device)
        check(lib().mflow_mlp_backward(ctypes.byref(d), ptr(x), x.stride(0), x.stride(1), ptr(context), d.n_ctx, _ptr_array(weights),
                                       _ptr_array(biases), ptr(g_out), ptr(g_x), _ptr_array(g_w), _ptr_array(g_b), ptr(ws), stream()),
              "mflow_mlp_backward")
        cabi.count_launch(2)
        return (g_x, None, None, None) + tuple(g_w) + tuple(g_b)


def residual_net_fused(net, inputs, context):
    """ResidualNet forward through ``mflow_mlp_forward`` (one launch; two for the backward)."""
    layers = [net.initial_layer]
    for block in net.blocks:
        layers += [block.linear_layers[0], block.linear_layers[1]]
        if context is not None:
            layers.append(block.context_layer)
    layers.append(net.final_layer)
    params = [layer.weight for layer in layers] + [layer.bias for layer in layers]
    return _MlpFused.apply(inputs, context, len(layers), len(net.blocks), *params)


class _Fork(torch.autograd.Function):
    """x -> (x, x): the input of a residual block taken once for the branch and once for the skip connection, so that the
    two gradients meet HERE and are added by one launch that also publishes max |sum| - autograd's own accumulation is
    an ATen add, after which the consuming convolution needs an absmax pass over the sum for its operand scale."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, g_branch, g_skip):
        if g_branch is None or g_skip is None:
            return g_skip if g_branch is None else g_branch
        a, b = _c(g_branch), _c(g_skip)
        if a.numel() % 4 or a.dtype != torch.float32 or not a.is_cuda:
            return a + b
        out = torch.empty_like(a)
        slot = _amax_pool.take(a.device) if config.conv_operand_format == "fp16" else None
        check(lib().mflow_add_amax(ptr(a), ptr(b), ptr(out), ptr(slot), a.numel(), stream()), "mflow_add_amax")
        cabi.count_launch()
        if slot is not None:
            _publish_amax(out, slot)
        return out


def fork(x):
    """(x, x) as two autograd outputs whose gradients are summed by ``mflow_add_amax``; the views inherit x's amax tag."""
    a, b = _Fork.apply(x)
    tag = getattr(x, "_mflow_amax", None)
    if tag is not None and tag[1] == x._version and tag[2] == x.data_ptr():
        a._mflow_amax = (tag[0], a._version, a.data_ptr())
        b._mflow_amax = (tag[0], b._version, b.data_ptr())
    return a, b


class _GluResidual(torch.autograd.Function):
    """out = h + t * sigmoid(g): the gated residual of a ResidualBlock with context (nn/resnet.py:36-40) in one launch each
    way; the absolute maxima of the results ride along for the fp16-split convolutions that consume them."""

    @staticmethod
    def forward(ctx, h, t, g):
        require_cuda(h, t, g)
        h, t, g = _c(h), _c(t), _c(g)
        out = torch.empty_like(h)
        slot = _amax_pool.take(h.device) if config.conv_operand_format == "fp16" else None
        check(lib().mflow_glu_residual(ptr(h), ptr(t), ptr(g), ptr(out), ptr(slot), h.numel(), stream()), "mflow_glu_residual")
        cabi.count_launch()
        if slot is not None:
            _publish_amax(out, slot)
        ctx.save_for_backward(t, g)
        return out

    @staticmethod
    def backward(ctx, g_out):
        t, g = ctx.saved_tensors
        g_out = _c(g_out)
        g_t, g_g = torch.empty_like(t), torch.empty_like(g)
        f16 = config.conv_operand_format == "fp16"
        s_t = _amax_pool.take(t.device) if f16 else None
        s_g = _amax_pool.take(t.device) if f16 else None
        check(lib().mflow_glu_residual_backward(ptr(t), ptr(g), ptr(g_out), ptr(g_t), ptr(g_g), ptr(s_t), ptr(s_g), t.numel(), stream()),
              "mflow_glu_residual_backward")
        cabi.count_launch()
        if f16:
            _publish_amax(g_t, s_t)
            _publish_amax(g_g, s_g)
        return g_out, g_t, g_g


def glu_residual(h, t, g):
    """h + t * sigmoid(g) for same-shaped contiguous fp32 CUDA tensors with numel % 4 == 0."""
    return _GluResidual.apply(h, t, g)


# --------------------------------------------------------------------------------------------
# blocked triangular solve for the wide LU layers (inverse direction of LULinear, lu.py:75-95)
# --------------------------------------------------------------------------------------------
def _blocked_trisolve(mat, rhs, upper, unit, block):
    """X with mat X = rhs; mat [n, n] triangular (any strides), rhs [n, m].  Substitution over diagonal blocks:
    the blocks are inverted once (one small batched solve), everything else is GEMMs on the trailing
    right-hand sides - 2 n / block launches of well-shaped matrix products instead of cuBLAS trsm's ~30 narrow
    kernels (1 ms per 3840 x 3840 solve with 256 right-hand sides; 0.3 ms this way)."""
    n = mat.shape[0]
    x = rhs.clone()
    starts = list(range(0, n, block))
    nfull = n // block
    dinv = [None] * len(starts)
    if nfull:
        diag = torch.stack([mat[i * block:(i + 1) * block, i * block:(i + 1) * block] for i in range(nfull)])
        eye = torch.eye(block, dtype=mat.dtype, device=mat.device).expand(nfull, block, block)
        inv = torch.linalg.solve_triangular(diag, eye, upper=upper, unitriangular=unit)
        for i in range(nfull):
            dinv[i] = inv[i]
    if nfull < len(starts):
        s0 = starts[-1]
        tail = mat[s0:, s0:]
        dinv[-1] = torch.linalg.solve_triangular(tail, torch.eye(n - s0, dtype=mat.dtype, device=mat.device), upper=upper, unitriangular=unit)
    out = torch.empty_like(x)   # solved blocks land here; x keeps the running right-hand side
    order = range(len(starts) - 1, -1, -1) if upper else range(len(starts))
    for i in order:
        lo, hi = starts[i], min(starts[i] + block, n)
        xi = out[lo:hi]
        torch.mm(dinv[i], x[lo:hi], out=xi)
        if upper and lo > 0:
            x[:lo].addmm_(mat[:lo, lo:hi], xi, alpha=-1.0)
        elif not upper and hi < n:
            x[hi:].addmm_(mat[hi:, lo:hi], xi, alpha=-1.0)
    return out


class _TriSolve(torch.autograd.Function):
    """X = T^{-1} B for a dense triangular T [n, n] and B [n, m] (fp32, CUDA)."""

    @staticmethod
    def forward(ctx, tri, rhs, upper, unit, block):
        require_cuda(tri, rhs)
        x = _blocked_trisolve(tri, _c(rhs), upper, unit, block)
        ctx.save_for_backward(tri, x)
        ctx.cfg = (upper, unit, block)
        return x

    @staticmethod
    def backward(ctx, g):
        tri, x = ctx.saved_tensors
        upper, unit, block = ctx.cfg
        g_rhs = _blocked_trisolve(tri.t(), _c(g), not upper, unit, block)    # T^{-T} g
        g_tri = None
        if ctx.needs_input_grad[0]:
            g_tri = -(g_rhs @ x.t())   # dense; the caller's graph keeps only the entries that are parameters
        return g_tri, (g_rhs if ctx.needs_input_grad[1] else None), None, None, None


WIDE_CHUNK = 960   # reduction length per launch of a wide Linear layer (30 pipeline steps of 32 channels)


def _wide_product(planes, weight, transposed, bias, owner, version, fmt, amax):
    """planes [n, K, w, w] x weight [J, K] (or weight^T when ``transposed``: planes [n, J, w, w]) -> [n, J | K, w, w] on the
    tensor-core convolution kernel, the reduction cut into chunks of WIDE_CHUNK channels that are chained through the
    kernel's residual input: the tensor core accumulates with truncation, so one 3840-term chain loses ~1e-5
    relative, four 960-term chains joined by round-to-nearest adds stay at the convolution kernels' ~2e-6."""
    n_red = weight.shape[0] if transposed else weight.shape[1]
    n_out = weight.shape[1] if transposed else weight.shape[0]
    out = None
    for c0 in range(0, n_red, WIDE_CHUNK):
        c1 = min(n_red, c0 + WIDE_CHUNK)
        xs = _c(planes[:, c0:c1])
        w = (weight[c0:c1, :] if transposed else weight[:, c0:c1])[:, :, None, None]
        split = _weight_cache.get(w, transposed, owner, fmt, ((version[0], c0), version[1]))
        out = _conv_tc_launch(xs, split, n_out, 1, bias if c0 == 0 else None, out, None, False, amax)
    return out


class _WideLinear(torch.autograd.Function):
    """y = W x + b on plane-layout activations for a WIDE layer (the 3840-feature LU triangles of the image
    post-processing, lu.py:56-73): chunked tensor-core products forward and for the data gradient, one pixels-as-K
    weight-gradient launch.  ``owner`` / ``version`` key the cached fp16 split of the (per-pass rebuilt) dense weight."""

    @staticmethod
    def forward(ctx, planes, weight, bias, owner, version):
        require_cuda(planes, weight, bias)
        planes = _c(planes)
        fmt = _conv_format()
        x_amax = tensor_amax(planes) if fmt == "fp16" else None
        out = _wide_product(planes, weight, False, bias, owner, version, fmt, x_amax)
        ctx.save_for_backward(planes, weight)
        ctx.cfg = (owner, version, fmt, x_amax, bias is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        planes, weight = ctx.saved_tensors
        owner, version, fmt, x_amax, has_bias = ctx.cfg
        g = _c(g)
        g_amax = tensor_amax(g) if fmt == "fp16" else None
        g_x = _wide_product(g, weight, True, None, owner, version, fmt, g_amax) if ctx.needs_input_grad[0] else None
        g_w = g_b = None
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            g_w, g_b = conv_wgrad_tc(g, planes, 1, False, with_bias=True, x_amax=x_amax, g_amax=g_amax)
            g_w = g_w.view(weight.shape)
        return g_x, g_w, (g_b if has_bias else None), None, None


def wide_linear(planes, weight, bias, owner, version):
    return _WideLinear.apply(planes, weight, bias, owner, version)


class _WideSolve(torch.autograd.Function):
    """X = T^-1 B on plane-layout right-hand sides [n, C, w, w] with a PRECOMPUTED inverse (the triangle's parameters have
    not changed since it was built): chunked 1x1 convolutions on the tensor cores.  Backward: g_B = T^-T g (the same
    kernel with the transposed inverse), g_T = -(g_B) X^T over all pixels (the weight-gradient kernel)."""

    @staticmethod
    def forward(ctx, rhs, tri, inv, upper, owner, version):
        require_cuda(rhs, inv)
        fmt = _conv_format()
        rhs = _c(rhs)
        x_amax = tensor_amax(rhs) if fmt == "fp16" else None
        out = _wide_product(rhs, inv, False, None, owner, version, fmt, x_amax)
        ctx.save_for_backward(inv, out)
        ctx.cfg = (owner, version, fmt)
        return out

    @staticmethod
    def backward(ctx, g):
        inv, out = ctx.saved_tensors
        owner, version, fmt = ctx.cfg
        g = _c(g)
        g_amax = tensor_amax(g) if fmt == "fp16" else None
        g_rhs = _wide_product(g, inv, True, None, owner, version, fmt, g_amax)
        g_tri = None
        if ctx.needs_input_grad[1]:
            g_tri = -conv_wgrad_tc(g_rhs, out, 1, False).view(inv.shape)
        return g_rhs, g_tri, None, None, None, None


def wide_solve(rhs_planes, tri, inv, upper, owner, version):
    return _WideSolve.apply(rhs_planes, tri, inv, bool(upper), owner, version)


def tri_solve(tri, rhs, upper, unit=False, block=256):
    return _TriSolve.apply(tri, rhs, bool(upper), bool(unit), int(block))


def channel_sum(x):
    """sum over all dims but 1 of a contiguous [B, C, ...] tensor (bias gradient of a convolution)."""
    require_cuda(x)
    x = _c(x)
    batch, chan = x.shape[0], x.shape[1]
    inner = x[0, 0].numel() if batch and chan else 0
    out = torch.empty(chan, dtype=torch.float32, device=x.device)
    ws = cabi.workspace("channel_sum", lib().mflow_channel_sum_workspace_bytes(batch, chan, inner), x.device)
    check(lib().mflow_channel_sum(ptr(x), ptr(out), batch, chan, inner, ptr(ws), stream()), "mflow_channel_sum")
    cabi.count_launch(2)
    return out


def conv_tc(x, weight, bias=None, residual=None, relu_in=False, weight_owner=None, weight_version=None):
    return _ConvTC.apply(x, weight, bias, residual, bool(relu_in), weight_owner, weight_version)


def rows_to_planes(t):
    """[rows, C] -> ([n, C, w, w] with n * w * w >= rows (zero padded), meta): the batch rows become the pixels of square
    feature maps, which is the layout the tensor-core convolution kernels take - a Linear layer is then a 1x1 convolution."""
    rows, c = t.shape
    w = 32 if rows > 256 else (16 if rows > 64 else (8 if rows > 16 else 4))
    n = (rows + w * w - 1) // (w * w)
    if n * w * w != rows:
        t = torch.cat((t, t.new_zeros(n * w * w - rows, c)), dim=0)
    return t.reshape(n, w * w, c).transpose(1, 2).contiguous().view(n, c, w, w), (rows, w)


def planes_to_rows(p, meta):
    rows, w = meta
    n, c = p.shape[0], p.shape[1]
    return p.view(n, c, w * w).transpose(1, 2).reshape(n * w * w, c)[:rows]


# --------------------------------------------------------------------------------------------
# (c) linear glue
# --------------------------------------------------------------------------------------------
def _mix_desc(x):
    d = MixDesc()
    if x.dim() == 4:
        b, c, h, w = x.shape
        d.n_outer, d.n_chan, d.n_inner = b, c, h * w
        d.x_so = d.y_so = c * h * w
        d.x_sc = d.y_sc = h * w
        d.x_si = d.y_si = 1
    else:
        b, c = x.shape
        d.n_outer, d.n_chan, d.n_inner = b, c, 1
        d.x_so = d.y_so = c
        d.x_sc = d.y_sc = 1
        d.x_si = d.y_si = 1
    return d


class _ChanMix(torch.autograd.Function):
    """y[b, :, hw] = M x[b, :, hw] + m for [B, C, H, W] or [B, C] activations."""

    @staticmethod
    def forward(ctx, x, mat, bias):
        require_cuda(x, mat, bias)
        x, mat, bias = _c(x), _c(mat), _c(bias)
        d = _mix_desc(x)
        y = torch.empty_like(x)
        check(lib().mflow_chanmix_apply(ctypes.byref(d), ptr(x), ptr(mat), ptr(bias), ptr(y), stream()), "mflow_chanmix_apply")
        cabi.count_launch()
        ctx.desc = d
        ctx.save_for_backward(x, mat)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, mat = ctx.saved_tensors
        d = ctx.desc
        g_y = _c(g_y)
        g_x = g_mat = g_bias = None
        if ctx.needs_input_grad[0]:
            g_x = torch.empty_like(x)
            mat_t = mat.t().contiguous()
            check(lib().mflow_chanmix_apply(ctypes.byref(d), ptr(g_y), ptr(mat_t), None, ptr(g_x), stream()), "mflow_chanmix_apply")
            cabi.count_launch()
        if (ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2])) and not _GradScope.inputs_only:
            g_mat = torch.empty_like(mat)
            g_bias = torch.empty(mat.shape[0], dtype=torch.float32, device=mat.device)
            ws = cabi.workspace("wgrad", lib().mflow_chanmix_wgrad_workspace_bytes(ctypes.byref(d)), x.device)
            check(lib().mflow_chanmix_wgrad(ctypes.byref(d), ptr(x), ptr(g_y), ptr(g_mat), ptr(g_bias), ptr(ws), stream()),
                  "mflow_chanmix_wgrad")
            cabi.count_launch()
            if not ctx.has_bias:
                g_bias = None
        return g_x, g_mat, g_bias


def chanmix(x, mat, bias=None):
    return _ChanMix.apply(x, mat, bias)


class _Squeeze(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inverse, scale, shift):
        require_cuda(x)
        x = _c(x)
        b, c, h, w = x.shape
        if inverse:
            if c < 4 or c % 4 != 0:
                raise ValueError("Invalid number of channel dimensions.")
            y = torch.empty((b, c // 4, h * 2, w * 2), dtype=x.dtype, device=x.device)
        else:
            if h % 2 != 0 or w % 2 != 0:
                raise ValueError("Input image size not compatible with the factor.")
            y = torch.empty((b, c * 4, h // 2, w // 2), dtype=x.dtype, device=x.device)
        check(lib().mflow_squeeze2x2(ptr(x), ptr(y), b, c, h, w, int(inverse), scale, shift, stream()), "mflow_squeeze2x2")
        cabi.count_launch()
        ctx.inverse, ctx.scale = inverse, scale
        return y

    @staticmethod
    def backward(ctx, g_y):
        g_y = _c(g_y)
        b, c, h, w = g_y.shape
        # forward y = s * squeeze(x) + t  ->  g_x = unsqueeze(g_y) * s ; inverse: g_x = squeeze(g_y) / s
        if ctx.inverse:
            g_x = torch.empty((b, c * 4, h // 2, w // 2), dtype=g_y.dtype, device=g_y.device)
            check(lib().mflow_squeeze2x2(ptr(g_y), ptr(g_x), b, c, h, w, 0, 1.0 / ctx.scale, 0.0, stream()), "mflow_squeeze2x2")
        else:
            g_x = torch.empty((b, c // 4, h * 2, w * 2), dtype=g_y.dtype, device=g_y.device)
            check(lib().mflow_squeeze2x2(ptr(g_y), ptr(g_x), b, c, h, w, 1, 1.0 / ctx.scale, 0.0, stream()), "mflow_squeeze2x2")
        cabi.count_launch()
        return g_x, None, None, None


def squeeze2x2(x, inverse=False, scale=1.0, shift=0.0):
    return _Squeeze.apply(x, bool(inverse), float(scale), float(shift))


class _AffineScalar(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, shift, inverse):
        require_cuda(x)
        x = _c(x)
        y = torch.empty_like(x)
        check(lib().mflow_affine_scalar(ptr(x), ptr(y), x.numel(), scale, shift, int(inverse), stream()), "mflow_affine_scalar")
        cabi.count_launch()
        ctx.factor = (1.0 / scale) if inverse else scale
        return y

    @staticmethod
    def backward(ctx, g_y):
        g_y = _c(g_y)
        g_x = torch.empty_like(g_y)
        check(lib().mflow_affine_scalar(ptr(g_y), ptr(g_x), g_y.numel(), ctx.factor, 0.0, 0, stream()), "mflow_affine_scalar")
        cabi.count_launch()
        return g_x, None, None, None


def affine_scalar(x, scale, shift, inverse=False):
    return _AffineScalar.apply(x, float(scale), float(shift), bool(inverse))


class _PermuteRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, perm, inv_perm):
        require_cuda(x)
        x = _c(x)
        y = torch.empty_like(x)
        check(lib().mflow_permute_rows(ptr(x), ptr(perm), ptr(y), x.shape[0], x.shape[1], stream()), "mflow_permute_rows")
        cabi.count_launch()
        ctx.save_for_backward(inv_perm)
        return y

    @staticmethod
    def backward(ctx, g_y):
        (inv_perm,) = ctx.saved_tensors
        g_y = _c(g_y)
        g_x = torch.empty_like(g_y)
        check(lib().mflow_permute_rows(ptr(g_y), ptr(inv_perm), ptr(g_x), g_y.shape[0], g_y.shape[1], stream()), "mflow_permute_rows")
        cabi.count_launch()
        return g_x, None, None


def permute_rows(x, perm, inv_perm):
    """y[:, j] = x[:, perm[j]] for a 2-D x (int64 CUDA index tensors)."""
    return _PermuteRows.apply(x, perm, inv_perm)


class _LogTanh(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inverse, cut_point):
        require_cuda(x)
        x = _c(x)
        rows, n = x.shape[0], x[0].numel()
        y = torch.empty_like(x)
        lad = torch.empty(rows, dtype=torch.float32, device=x.device)
        check(lib().mflow_logtanh_apply(ptr(x), ptr(y), ptr(lad), rows, n, int(inverse), cut_point, stream()), "mflow_logtanh_apply")
        cabi.count_launch()
        ctx.inverse, ctx.cut = inverse, cut_point
        ctx.save_for_backward(x)
        return y, lad

    @staticmethod
    def backward(ctx, g_y, g_lad):
        (x,) = ctx.saved_tensors
        g_x = torch.empty_like(x)
        check(lib().mflow_logtanh_backward(ptr(x), ptr(None if g_y is None else _c(g_y)), ptr(None if g_lad is None else _c(g_lad)),
                                           ptr(g_x), x.shape[0], x[0].numel(), int(ctx.inverse), ctx.cut, stream()),
              "mflow_logtanh_backward")
        cabi.count_launch()
        return g_x, None, None


def logtanh(x, inverse=False, cut_point=1.0):
    return _LogTanh.apply(x, bool(inverse), float(cut_point))


# --------------------------------------------------------------------------------------------
# likelihood epilogues
# --------------------------------------------------------------------------------------------
class _GaussLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, v, add0, add1, eps, clip):
        require_cuda(u, v, add0, add1)
        if u.stride(-1) != 1 or (v is not None and v.stride(-1) != 1):
            u, v = u.contiguous(), (None if v is None else v.contiguous())
        rows, n_u = u.shape
        n_v = 0 if v is None else v.shape[1]
        out = torch.empty(rows, dtype=torch.float32, device=u.device)
        check(lib().mflow_gauss_logprob(ptr(u), n_u, u.stride(0), ptr(v), n_v, 0 if v is None else v.stride(0), eps, clip,
                                        ptr(None if add0 is None else _c(add0)), ptr(None if add1 is None else _c(add1)), ptr(out),
                                        rows, stream()), "mflow_gauss_logprob")
        cabi.count_launch()
        ctx.eps, ctx.clip = eps, clip
        ctx.save_for_backward(u, v)
        ctx.has = (add0 is not None, add1 is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        u, v = ctx.saved_tensors
        g = _c(g)
        rows, n_u = u.shape
        n_v = 0 if v is None else v.shape[1]
        g_u = torch.empty((rows, n_u), dtype=torch.float32, device=u.device)
        g_v = None if v is None else torch.empty((rows, n_v), dtype=torch.float32, device=u.device)
        check(lib().mflow_gauss_logprob_backward(ptr(u), n_u, u.stride(0), ptr(v), n_v, 0 if v is None else v.stride(0), ctx.eps,
                                                 ctx.clip, ptr(g), ptr(g_u), ptr(g_v), rows, stream()), "mflow_gauss_logprob_backward")
        cabi.count_launch()
        return g_u, g_v, (g if ctx.has[0] else None), (g if ctx.has[1] else None), None, None


def gauss_logprob(u, v=None, eps=1.0, clip=None, add0=None, add1=None):
    """logN(u; 0, 1) [+ logN(clamp(v); 0, eps^2)] [+ add0] [+ add1], per row."""
    return _GaussLogProb.apply(u, v, add0, add1, float(eps), float(clip) if clip else 0.0)


class _EvalOnly(torch.autograd.Function):
    """Identity whose backward raises: marks results that carry no derivative in this build (the full-Jacobian
    likelihood of mode="mf" would need double backward through every kernel - SURVEY.md section 8(f) rank 4)."""

    @staticmethod
    def forward(ctx, value, anchor, what):
        ctx.what = what
        return value.detach().clone()

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError(
            f"{ctx.what} is evaluation-only in the B200 build: its gradient needs second-order autograd through the fused "
            "kernels, which is not implemented.  Train with mode='mf-fixed-manifold' / 'projection' (the --sequential and "
            "--alternate schedules of experiments/train.py); train_manifold_flow / train_specified_manifold_flow "
            "(simultaneous M-flow training, train.py:145-156) and the SCANDAL loss are unsupported.")


def eval_only(value, anchor, what):
    """``value`` without a graph when nothing upstream needs one; otherwise a node that fails loudly in backward."""
    if value is None or not (torch.is_grad_enabled() and anchor is not None and anchor.requires_grad):
        return value
    return _EvalOnly.apply(value, anchor, what)


def jtj_logdet(jac):
    """log det(J^T J) per sample for J [B, D, n], n <= 32 (no gradient)."""
    require_cuda(jac)
    jac = _c(jac.detach())
    b, d, n = jac.shape
    out = torch.empty(b, dtype=torch.float32, device=jac.device)
    check(lib().mflow_jtj_logdet(ptr(jac), ptr(out), b, d, n, stream()), "mflow_jtj_logdet")
    cabi.count_launch()
    return out
