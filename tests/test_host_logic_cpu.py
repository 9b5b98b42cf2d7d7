"""CPU tests of host-side numerics that do not need the CUDA library (pure torch helpers of the product)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "manifold-flow_b200"))


@pytest.mark.parametrize("n,m,block", [(700, 33, 256), (512, 8, 256), (130, 3, 64), (64, 5, 256)])
def test_blocked_trisolve_math(n, m, block):
    """The blocked substitution behind ops.tri_solve (wide LULinear inverse, reference lu.py:75-95) solves op(T) X = B
    for lower-unit, upper, and their transposes (the backward pass uses the transposed systems)."""
    from manifold_flow.b200.ops import _blocked_trisolve

    torch.manual_seed(0)
    lower = torch.eye(n, dtype=torch.float64) + 0.05 * torch.randn(n, n, dtype=torch.float64).tril(-1)
    upper = torch.diag(torch.rand(n, dtype=torch.float64) + 0.5) + 0.05 * torch.randn(n, n, dtype=torch.float64).triu(1)
    rhs = torch.randn(n, m, dtype=torch.float64)
    for mat, is_upper, unit in ((lower, False, True), (upper, True, False), (lower.t(), True, True), (upper.t(), False, False)):
        x = _blocked_trisolve(mat, rhs, is_upper, unit, block)
        assert (mat @ x - rhs).abs().max().item() < 1e-10
    # a unit-triangular solve must ignore whatever sits on the stored diagonal
    junk = lower.clone()
    junk.diagonal().fill_(7.0)
    assert torch.allclose(_blocked_trisolve(junk, rhs, False, True, block), _blocked_trisolve(lower, rhs, False, True, block), atol=1e-12)


def test_lu_group_matches_per_layer_builds():
    """lu.LUGroup (batched dense L / U / W / W^-1 / logabsdet of same-width LU layers, reference lu.py:44-128 per layer)
    reproduces the per-layer construction, values and parameter gradients."""
    import copy

    from torch import nn

    from manifold_flow import transforms
    from manifold_flow.transforms.lu import assign_lu_groups

    torch.manual_seed(5)
    layers = nn.ModuleList([transforms.OneByOneConvolution(12) for _ in range(4)] + [transforms.LULinear(12) for _ in range(3)]
                           + [transforms.OneByOneConvolution(24)])
    for p in layers.parameters():
        p.data.normal_(0.0, 0.3)
    grouped = copy.deepcopy(layers)
    assign_lu_groups(grouped)
    assert all(getattr(m, "_lu_group", None) is not None for m in list(grouped)[:7])
    assert getattr(grouped[7], "_lu_group", None) is None          # a single layer of its width stays on its own
    assert grouped[0]._lu_group is not grouped[4]._lu_group        # different classes are not mixed

    def scalar(ms):
        total = 0.0
        for i, m in enumerate(ms):
            w, wi, lad = m.weight(), m.weight_inverse(), m.logabsdet()
            total = total + (w * (i + 1)).sin().sum() + (wi * 0.5).cos().sum() + lad * (i + 2)
        return total

    from manifold_flow.transforms import fusion

    fusion.begin_pass()
    ref, got = scalar(layers), scalar(grouped)
    assert torch.allclose(ref, got, rtol=1e-5, atol=1e-5)
    for a, b in zip(layers, grouped):
        assert torch.allclose(a.weight(), b.weight(), atol=1e-6)
        assert torch.allclose(a.weight_inverse(), b.weight_inverse(), atol=1e-5)
        assert torch.allclose(a.logabsdet(), b.logabsdet(), atol=1e-6)
    ref.backward()
    got.backward()
    for (n, pa), (_, pb) in zip(layers.named_parameters(), grouped.named_parameters()):
        if pa.grad is None:
            assert pb.grad is None, n
        else:
            assert torch.allclose(pa.grad, pb.grad, rtol=1e-4, atol=1e-5), n


def test_grouped_affine_maps_match_per_layer_maps():
    """The stacked (M, m, logabsdet) maps of lu.LUGroup / normalization.ActNormGroup equal the per-layer
    `_affine_map` results (conv.py:33-55, lu.py:56-95, normalization.py:99-131), values and gradients, both directions."""
    import copy

    from torch import nn

    from manifold_flow import transforms
    from manifold_flow.transforms.lu import assign_lu_groups

    torch.manual_seed(9)
    layers = nn.ModuleList([transforms.OneByOneConvolution(12) for _ in range(3)] + [transforms.LULinear(12) for _ in range(2)]
                           + [transforms.ActNorm(12) for _ in range(3)])
    for p in layers.parameters():
        p.data.normal_(0.0, 0.3)
    for m in layers:
        if isinstance(m, transforms.ActNorm):
            m.initialized = True
    grouped = copy.deepcopy(layers)
    assign_lu_groups(grouped)
    assert all(getattr(m, "_lu_group", None) is not None for m in list(grouped)[:5])
    assert all(getattr(m, "_an_group", None) is not None for m in list(grouped)[5:])
    from manifold_flow.transforms import fusion

    for inverse in (False, True):
        fusion.begin_pass()   # a model-level forward starts a new pass: graphs cached in the previous one are not reused
        for ms in (layers, grouped):
            for p in ms.parameters():
                p.grad = None
        totals = []
        for ms in (layers, grouped):
            total = 0.0
            for i, m in enumerate(ms):
                mat, bias, lad = m._affine_map(inverse)
                total = total + (mat * (i + 1)).sin().sum() + (bias * 0.7).cos().sum() + lad * (i + 2)
            totals.append(total)
        assert torch.allclose(totals[0], totals[1], rtol=1e-5, atol=1e-5)
        for a, b in zip(layers, grouped):
            for x, y in zip(a._affine_map(inverse), b._affine_map(inverse)):
                assert torch.allclose(x, y, atol=1e-5)
        totals[0].backward()
        totals[1].backward()
        for (n, pa), (_, pb) in zip(layers.named_parameters(), grouped.named_parameters()):
            assert (pa.grad is None) == (pb.grad is None), n
            if pa.grad is not None:
                assert torch.allclose(pa.grad, pb.grad, rtol=1e-4, atol=1e-5), n


def test_pair_group_composition_matches_sequential_composition():
    """fusion.PairGroup: conv(actnorm(x)) and its inverse composed for all pairs at once equal the per-pair
    composition of fusion.apply_affine (values and gradients)."""
    import copy

    from torch import nn

    from manifold_flow import transforms
    from manifold_flow.transforms import fusion
    from manifold_flow.transforms.lu import assign_lu_groups

    torch.manual_seed(21)
    steps = [transforms.CompositeTransform([transforms.ActNorm(12), transforms.OneByOneConvolution(12)]) for _ in range(3)]
    model = nn.ModuleList(steps)
    for p in model.parameters():
        p.data.normal_(0.0, 0.3)
    for m in model.modules():
        if isinstance(m, transforms.ActNorm):
            m.initialized = True
    grouped = copy.deepcopy(model)
    assign_lu_groups(grouped)
    assert all(getattr(s._transforms[0], "_pair_group", None) is not None for s in grouped)

    def compose(maps):
        mat, bias, lad = maps[0]
        for m2, b2, l2 in maps[1:]:
            bias, mat, lad = m2 @ bias + b2, m2 @ mat, lad + l2
        return mat, bias, lad

    for inverse in (False, True):
        fusion.begin_pass()
        ref_total, got_total = 0.0, 0.0
        for i, (ref_step, got_step) in enumerate(zip(model, grouped)):
            order = [1, 0] if inverse else [0, 1]
            ref = compose([ref_step._transforms[j]._affine_map(inverse) for j in order])
            act = got_step._transforms[0]
            stacked = act._pair_group.composed(inverse)
            got = tuple(t[act._pair_index] for t in stacked)
            for a, b in zip(ref, got):
                assert torch.allclose(a, b, atol=1e-5)
            ref_total = ref_total + (ref[0] * (i + 1)).sin().sum() + ref[1].cos().sum() + ref[2]
            got_total = got_total + (got[0] * (i + 1)).sin().sum() + got[1].cos().sum() + got[2]
        for ms in (model, grouped):
            for p in ms.parameters():
                p.grad = None
        ref_total.backward()
        got_total.backward()
        for (n, pa), (_, pb) in zip(model.named_parameters(), grouped.named_parameters()):
            assert torch.allclose(pa.grad, pb.grad, rtol=1e-4, atol=1e-5), n


def test_pair_group_permutation_lu_pairs():
    """(RandomPermutation -> LULinear) glue of the vector flows (vector_transforms.py:139-177): PairGroup composition
    equals the sequential composition, both directions, values and gradients."""
    import copy

    from torch import nn

    from manifold_flow import transforms
    from manifold_flow.transforms import fusion
    from manifold_flow.transforms.lu import assign_lu_groups

    torch.manual_seed(23)
    steps = [transforms.CompositeTransform([transforms.RandomPermutation(features=14), transforms.LULinear(14)]) for _ in range(4)]
    model = nn.ModuleList(steps)
    for p in model.parameters():
        p.data.normal_(0.0, 0.3)
    grouped = copy.deepcopy(model)
    assign_lu_groups(grouped)
    assert all(getattr(s._transforms[0], "_pair_group", None) is not None for s in grouped)
    for inverse in (False, True):
        fusion.begin_pass()
        ref_total, got_total = 0.0, 0.0
        for i, (ref_step, got_step) in enumerate(zip(model, grouped)):
            order = [1, 0] if inverse else [0, 1]
            maps = [ref_step._transforms[j]._affine_map(inverse) for j in order]
            mat, bias, lad = maps[0]
            mat, bias, lad = maps[1][0] @ mat, maps[1][0] @ bias + maps[1][1], lad + maps[1][2]
            first = got_step._transforms[0]
            got = tuple(t[first._pair_index] for t in first._pair_group.composed(inverse))
            for a, b in zip((mat, bias, lad), got):
                assert torch.allclose(a, b, atol=1e-5)
            ref_total = ref_total + (mat * (i + 1)).sin().sum() + bias.cos().sum() + lad
            got_total = got_total + (got[0] * (i + 1)).sin().sum() + got[1].cos().sum() + got[2]
        for ms in (model, grouped):
            for p in ms.parameters():
                p.grad = None
        ref_total.backward()
        got_total.backward()
        for (n, pa), (_, pb) in zip(model.named_parameters(), grouped.named_parameters()):
            assert torch.allclose(pa.grad, pb.grad, rtol=1e-4, atol=1e-5), n


def _glue_model(kind):
    from torch import nn

    from manifold_flow import transforms

    if kind == "actnorm":
        steps = [transforms.CompositeTransform([transforms.ActNorm(12), transforms.OneByOneConvolution(12)]) for _ in range(3)]
    else:
        steps = [transforms.CompositeTransform([transforms.RandomPermutation(features=14), transforms.LULinear(14)]) for _ in range(3)]
    model = nn.ModuleList(steps)
    for p in model.parameters():
        p.data.normal_(0.0, 0.3)
    return model


def _composed_maps(model, inverse):
    from manifold_flow.transforms import fusion

    fusion.begin_pass()
    out = []
    for step in model:
        first = step._transforms[0]
        out.append(tuple(t[first._pair_index].clone() for t in first._pair_group.composed(inverse)))
    return out


@pytest.mark.parametrize("kind", ["actnorm", "perm"])
def test_caches_follow_load_state_dict_with_other_permutations(kind):
    """the permutation buffers are random per construction; after load_state_dict of a checkpoint with other
    permutations every grouped 1x1 convolution / (permutation, LU) pair must use the NEW permutation."""
    from manifold_flow import transforms
    from manifold_flow.transforms.lu import assign_lu_groups

    torch.manual_seed(1)
    a = _glue_model(kind)
    torch.manual_seed(2)
    b = _glue_model(kind)           # other weights AND other permutations
    for m in list(a.modules()) + list(b.modules()):
        if isinstance(m, transforms.ActNorm):
            m.initialized = True
    a.eval(), b.eval()
    assign_lu_groups(a)
    assign_lu_groups(b)
    with torch.no_grad():
        for inverse in (False, True):
            _composed_maps(a, inverse)                 # fills every cache of `a` with its own permutations
        a.load_state_dict(b.state_dict())
        for inverse in (False, True):
            for got, want in zip(_composed_maps(a, inverse), _composed_maps(b, inverse)):
                for g, w in zip(got, want):
                    assert torch.equal(g, w)
        if kind == "actnorm":   # the un-paired grouped path (LUGroup.affine_maps) as well
            conv_a, conv_b = a[0]._transforms[1], b[0]._transforms[1]
            for inverse in (False, True):
                for g, w in zip(conv_a._affine_map(inverse), conv_b._affine_map(inverse)):
                    assert torch.equal(g, w)


def test_actnorm_data_init_invalidates_earlier_eval_caches():
    """eval forward (identity ActNorm maps cached) -> first training forward (data-dependent init writes the
    parameters) -> eval again WITHOUT an optimizer step must see the initialised statistics."""
    from manifold_flow import transforms
    from manifold_flow.transforms.lu import assign_lu_groups

    torch.manual_seed(4)
    model = _glue_model("actnorm")
    for m in model.modules():
        if isinstance(m, transforms.ActNorm):
            m.log_scale.data.zero_(), m.shift.data.zero_()
    assign_lu_groups(model)
    model.eval()
    with torch.no_grad():
        before = _composed_maps(model, False)
        x = torch.randn(64, 12, 3, 3) * 3.0 + 1.5
        for step in model:    # what the first training batch does (normalization.py:134-147)
            act = step._transforms[0]
            act.train()
            act._initialize(x)
            assert act.initialized and act.log_scale._version > 0
        model.eval()
        after = _composed_maps(model, False)
    for (m0, b0, l0), (m1, b1, l1), step in zip(before, after, model):
        assert not torch.allclose(m0, m1) and not torch.allclose(l0, l1)
        act, conv = step._transforms
        want = conv._affine_map(False)[0] @ torch.diag(torch.exp(act.log_scale))
        assert torch.allclose(m1, want, atol=1e-5)


def test_invalidate_caches_covers_data_writes():
    """Writes through .data do not bump Tensor._version: ddp.broadcast_parameters (and users) call invalidate_caches()."""
    from manifold_flow.b200.config import invalidate_caches
    from manifold_flow.transforms.lu import assign_lu_groups

    torch.manual_seed(6)
    model = _glue_model("perm")
    assign_lu_groups(model)
    with torch.no_grad():
        before = _composed_maps(model, False)
        for p in model.parameters():
            p.data.mul_(0.5)
        stale = _composed_maps(model, False)
        assert all(torch.equal(a[0], b[0]) for a, b in zip(before, stale))     # the hazard the epoch exists for
        invalidate_caches()
        fresh = _composed_maps(model, False)
        assert all(not torch.equal(a[0], b[0]) for a, b in zip(before, fresh))


def test_member_views_unbind_once_and_backpropagate_like_selects():
    """fusion.member_views hands out the members of a group's stacked maps through one unbind per tensor (memoised on
    the cached tuple): same values and the same gradients as per-member selects, rebuilt when the cache rebuilds."""
    from manifold_flow.transforms.fusion import member_views

    class Holder:
        pass

    torch.manual_seed(1)
    base = torch.randn(5, 3, 3, dtype=torch.float64, requires_grad=True)
    stacked = (base * 2.0, base.sum(dim=2), base.sum(dim=(1, 2)))
    holder = Holder()
    views = member_views(holder, "_members_fwd", stacked)
    assert member_views(holder, "_members_fwd", stacked) is views            # memoised on the tuple's identity
    weights = torch.randn(5, dtype=torch.float64)
    loss = sum(w * (views[0][i].sum() + views[1][i].sum() + views[2][i]) for i, w in enumerate(weights))
    (g_unbind,) = torch.autograd.grad(loss, base, retain_graph=True)
    loss_sel = sum(w * (stacked[0][i].sum() + stacked[1][i].sum() + stacked[2][i]) for i, w in enumerate(weights))
    (g_select,) = torch.autograd.grad(loss_sel, base)
    assert torch.equal(views[0][2], stacked[0][2]) and torch.allclose(g_unbind, g_select, atol=1e-14)
    rebuilt = (base * 3.0, base.sum(dim=2), base.sum(dim=(1, 2)))
    assert member_views(holder, "_members_fwd", rebuilt) is not views


def test_scoped_switches_restore_their_state():
    """ops.input_grads_only (full-Jacobian passes skip parameter gradients) and conditioner.planes_output (a coupling
    layer accepts its conditioner's NCHW result) are scoped: nested use and exceptions restore the previous state."""
    from manifold_flow.b200 import conditioner, ops

    assert not ops._GradScope.inputs_only and not conditioner._PlanesOut.enabled
    with ops.input_grads_only():
        assert ops._GradScope.inputs_only
        with ops.input_grads_only():
            assert ops._GradScope.inputs_only
        assert ops._GradScope.inputs_only
    assert not ops._GradScope.inputs_only
    with pytest.raises(RuntimeError):
        with conditioner.planes_output():
            assert conditioner._PlanesOut.enabled
            raise RuntimeError("boom")
    assert not conditioner._PlanesOut.enabled


def test_rows_x_planes_descriptor_fields():
    """The spline descriptor of a vector coupling whose parameters arrive as [rows / hw, C_t * P, h, w]: unshifted row
    pointers (offset 0), p_hw = h * w and t_off = the first transformed feature; the rows layout keeps p_hw = 0."""
    from manifold_flow.b200 import ops

    geom = ops.SplineGeometry(7, 1, 2, 7, 0, 2)
    x = torch.zeros(2048, 14)
    rows = torch.zeros(2048, 7 * 32)
    planes = torch.zeros(2, 7 * 32, 32, 32)
    d, off = ops._rqs_desc(x, rows, geom, 11, 10.0, 1e-3, 1e-3, 1e-3, 10.0, False)
    assert (d.p_hw, d.t_off, off) == (0, 0, 1) and d.n_outer == 2048 and d.n_chan == 7
    d, off = ops._rqs_desc(x, planes, geom, 11, 10.0, 1e-3, 1e-3, 1e-3, 10.0, False)
    assert (d.p_hw, d.t_off, off) == (1024, 1, 0) and d.x_so == 14 and d.x_sc == 2
