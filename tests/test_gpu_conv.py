"""GPU parity of kernel (b): tcgen05 / TMA 3xTF32 convolutions against fp64 and against cuDNN fp32,
forward and backward, for every feature-map size and both kernel sizes of ConvResidualNet."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CASES = [  # (batch, c_in, c_out, hw, kernel)
    (2, 100, 100, 32, 3), (3, 100, 100, 16, 3), (5, 100, 100, 8, 3), (9, 100, 100, 4, 3), (1, 100, 100, 4, 3),
    (2, 6, 100, 32, 1), (2, 12, 100, 16, 1), (3, 24, 100, 8, 1), (2, 48, 100, 4, 1),
    (2, 100, 192, 32, 1), (2, 100, 384, 16, 1), (2, 100, 768, 8, 1), (3, 100, 1536, 4, 1), (2, 16, 16, 8, 3),
]


def _rel(a, b):
    return float((a.double() - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("relu_in", [False, True])
def test_conv_forward_backward_vs_fp64(cuda_device, case, relu_in):
    from manifold_flow.b200 import ops

    b, ci, co, hw, k = case
    g = torch.Generator().manual_seed(b * 1000 + ci + co + hw)
    x = torch.randn(b, ci, hw, hw, generator=g)
    w = torch.randn(co, ci, k, k, generator=g) * (1.0 / (ci * k * k) ** 0.5)
    bias = torch.randn(co, generator=g)
    res = torch.randn(b, co, hw, hw, generator=g)
    gy = torch.randn(b, co, hw, hw, generator=g)

    def run(dtype, device, fn):
        leaves = [t.detach().clone().to(device=device, dtype=dtype).requires_grad_(True) for t in (x, w, bias, res)]
        y = fn(*leaves)
        y.backward(gy.to(device=device, dtype=dtype))
        return [y.detach()] + [t.grad for t in leaves]

    ref_fn = lambda xx, ww, bb, rr: F.conv2d(F.relu(xx) if relu_in else xx, ww, bb, padding=k // 2) + rr
    ref = run(torch.float64, "cpu", ref_fn)
    torch.backends.cudnn.allow_tf32 = False
    lib32 = run(torch.float32, cuda_device, ref_fn)
    got = run(torch.float32, cuda_device, lambda xx, ww, bb, rr: ops.conv_tc(xx, ww, bb, residual=rr, relu_in=relu_in))
    names = ["y", "grad_x", "grad_w", "grad_b", "grad_res"]
    for name, a, l, r in zip(names, got, lib32, ref):
        e_got, e_lib = _rel(a.cpu(), r), _rel(l.cpu(), r)
        # fp32-accurate: within 1e-5 of the fp64 result and no worse than 8x the fp32 library path
        assert e_got <= max(2e-5, 8 * e_lib), f"{name}: rel L2 error {e_got:.2e} (cuDNN fp32: {e_lib:.2e})"


def test_conv_resnet_matches_library_path(cuda_device):
    """ConvResidualNet through the tensor-core path == through cuDNN fp32, forward and gradients."""
    from manifold_flow import nn as nn_
    from manifold_flow.b200 import config

    torch.manual_seed(3)
    net = nn_.ConvResidualNet(6, 6 * 32, hidden_channels=100, num_blocks=2).to(cuda_device)
    x = torch.randn(4, 6, 32, 32, device=cuda_device)
    gy = torch.randn(4, 192, 32, 32, device=cuda_device)

    def run(flag):
        config.conv_tensor_cores = flag
        try:
            net.zero_grad()
            xx = x.clone().requires_grad_(True)
            y = net(xx)
            y.backward(gy)
            return [y.detach(), xx.grad] + [p.grad.clone() for p in net.parameters()]
        finally:
            config.conv_tensor_cores = True

    a, b = run(True), run(False)
    for u, v in zip(a, b):
        assert _rel(u, v.double()) <= 2e-3  # two fp32 paths; cuDNN picks non-deterministic wgrad algorithms


def test_unsupported_shapes_fall_back_loudly(cuda_device):
    from manifold_flow.b200 import ops

    x = torch.randn(2, 8, 12, 12, device=cuda_device)
    w = torch.randn(8, 8, 3, 3, device=cuda_device)
    assert not ops.conv_tc_supported(x, w)
    with pytest.raises(ValueError):
        ops.conv_tc(x, w)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(256, 100, 32, 32), (7, 100, 16, 16), (5, 192, 8, 8), (64, 48, 4, 4), (3, 5, 7), (1, 1, 1, 1), (0, 4, 2, 2),
                                   (300, 3, 33)])
def test_channel_sum_matches_fp64_sum(cuda_device, shape):
    """mflow_channel_sum (bias gradient of the conditioner convolutions) against an fp64 sum; deterministic."""
    from manifold_flow.b200 import ops

    torch.manual_seed(5)
    x = torch.randn(*shape, device=cuda_device) + 0.3
    got = ops.channel_sum(x)
    dims = [d for d in range(x.dim()) if d != 1]
    want = x.double().sum(dim=dims)
    scale = x.double().abs().sum(dim=dims).clamp_min(1.0)
    assert got.shape == want.shape
    assert ((got.double() - want).abs() / scale).max().item() < 1e-6 if got.numel() else True
    assert torch.equal(got, ops.channel_sum(x))


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(2, 100, 100, 32, 3, True), (3, 100, 100, 16, 3, True), (5, 100, 100, 8, 3, False), (2, 6, 100, 32, 1, False),
                                  (2, 100, 192, 32, 1, True), (3, 100, 384, 16, 1, True), (2, 24, 100, 8, 1, False), (1, 8, 8, 32, 3, False),
                                  (2, 130, 40, 16, 1, False), (7, 48, 100, 8, 1, True), (1, 1, 1, 8, 3, False), (6, 100, 100, 4, 3, True),
                                  (5, 48, 100, 4, 1, False), (3, 100, 1536, 4, 1, True), (1, 3, 5, 4, 3, False)])
def test_wgrad_tensor_core_matches_fp64_autograd(cuda_device, case):
    """mflow_conv2d_wgrad_tf32x3 against the fp64 autograd weight gradient of F.conv2d (reference: nn/resnet.py conv layers)."""
    from manifold_flow.b200 import ops

    b, ci, co, hw, k, relu = case
    torch.manual_seed(11)
    x = torch.randn(b, ci, hw, hw, device=cuda_device)
    g = torch.randn(b, co, hw, hw, device=cuda_device)
    w = torch.zeros(co, ci, k, k, device=cuda_device, dtype=torch.float64, requires_grad=True)
    xin = (F.relu(x) if relu else x).double()
    F.conv2d(xin, w, padding=k // 2).backward(g.double())
    got = ops.conv_wgrad_tc(g, x, k, relu)
    assert got.shape == w.grad.shape
    rel = ((got.double() - w.grad).norm() / w.grad.norm()).item()
    assert rel < 2e-6, rel                                   # fp32 library wgrad sits at 1e-7 .. 2e-6 on these
    assert torch.equal(got, ops.conv_wgrad_tc(g, x, k, relu))   # deterministic
    # the bias gradient rides along in the same kernel
    got2, gb = ops.conv_wgrad_tc(g, x, k, relu, with_bias=True)
    assert torch.equal(got, got2)
    want_b = g.double().sum(dim=(0, 2, 3))
    scale_b = g.double().abs().sum(dim=(0, 2, 3)).clamp_min(1.0)
    assert ((gb.double() - want_b).abs() / scale_b).max().item() < 1e-6


@pytest.mark.gpu
@pytest.mark.timeout(300, method="thread")
@pytest.mark.parametrize("case", [(1024, 100, 100, 32, 1, True, 600), (256, 100, 100, 32, 1, True, 1500), (256, 2, 100, 32, 1, False, 600),
                                  (256, 100, 192, 32, 1, False, 300), (256, 100, 100, 32, 3, True, 100), (512, 100, 100, 4, 1, True, 600)])
def test_wgrad_repeatable_under_load(cuda_device, case):
    """Back-to-back launches of the (deterministic) weight-gradient kernel on bench-sized inputs must all return the same
    bits.  Regression test for the raw-box ring of the 1x1 kernels: with three slots a transform set could pass an
    mbarrier wait on a stale phase parity (rare wrong results / hangs / launch failures in the large-batch vector
    configurations, where ResidualNet's Linear layers run as 1x1 convolutions with relu_in)."""
    from manifold_flow.b200 import ops

    b, ci, co, hw, k, relu, reps = case
    torch.manual_seed(3)
    x = torch.randn(b, ci, hw, hw, device=cuda_device)
    g = torch.randn(b, co, hw, hw, device=cuda_device)
    first = ops.conv_wgrad_tc(g, x, k, relu)
    bad = torch.zeros((), dtype=torch.int32, device=cuda_device)
    for _ in range(reps):   # no host synchronisation inside the loop: the launches queue up back to back
        bad += (ops.conv_wgrad_tc(g, x, k, relu) != first).any()
    assert int(bad) == 0
    # accuracy at this size: sums of up to 10^6 random-sign products cancel down to ~10^3, so the error is measured against
    # the sum of the products' magnitudes (backward error); the tensor core's fp32 accumulator truncates, which shows up as
    # a relative bias of ~3e-8 of that scale at 10^6 pixels
    xin = (F.relu(x) if relu else x).double()
    w = torch.zeros(co, ci, k, k, device=cuda_device, dtype=torch.float64, requires_grad=True)
    F.conv2d(xin, w, padding=k // 2).backward(g.double())
    w_abs = torch.zeros(co, ci, k, k, device=cuda_device, dtype=torch.float64, requires_grad=True)
    F.conv2d(xin.abs(), w_abs, padding=k // 2).backward(g.double().abs())
    assert ((first.double() - w.grad).abs() / w_abs.grad.clamp_min(1e-30)).max().item() < 5e-7


@pytest.mark.gpu
@pytest.mark.parametrize("rows,n_in,n_out,ctx", [(4096, 20, 640, None), (8192, 7, 224, 2), (1024, 33, 96, 1), (4112, 3, 23, None)])
def test_residual_net_tensor_core_path_matches_fp64(cuda_device, rows, n_in, n_out, ctx):
    """ResidualNet (reference nn/resnet.py:43-72) through the tcgen05 convolution kernels (batch rows as pixels) against
    an fp64 evaluation of the same formulas: outputs and all gradients.  Rows in which some pre-activation lies within
    1e-4 of the ReLU kink get a zero incoming gradient: there the fp32 and fp64 masks may legitimately differ (one such
    element among 400k changes the weight gradients by 1e-3)."""
    from manifold_flow import nn as nn_
    from manifold_flow.b200.config import config

    torch.manual_seed(17)
    net = nn_.ResidualNet(n_in, n_out, hidden_features=100, context_features=ctx, num_blocks=2).to(cuda_device)
    for p in net.parameters():   # the last layer of each block is initialised to ~0 (resnet.py:23-25): make every path matter
        p.data.normal_(0.0, 0.2)
    x = torch.randn(rows, n_in, device=cuda_device, requires_grad=True)
    c = None if ctx is None else torch.randn(rows, ctx, device=cuda_device)
    g = torch.randn(rows, n_out, device=cuda_device)

    P = {n: p.detach().double().requires_grad_() for n, p in net.named_parameters()}
    x64 = x.detach().double().requires_grad_()
    c64 = None if c is None else c.double()
    h = F.linear(x64 if c is None else torch.cat((x64, c64), 1), P["initial_layer.weight"], P["initial_layer.bias"])
    near_kink = torch.zeros(rows, dtype=torch.bool, device=cuda_device)
    for i in range(2):
        near_kink |= h.detach().abs().min(dim=1).values < 1e-4
        t = F.linear(F.relu(h), P[f"blocks.{i}.linear_layers.0.weight"], P[f"blocks.{i}.linear_layers.0.bias"])
        near_kink |= t.detach().abs().min(dim=1).values < 1e-4
        t = F.linear(F.relu(t), P[f"blocks.{i}.linear_layers.1.weight"], P[f"blocks.{i}.linear_layers.1.bias"])
        if c is not None:
            t = t * torch.sigmoid(F.linear(c64, P[f"blocks.{i}.context_layer.weight"], P[f"blocks.{i}.context_layer.bias"]))
        h = h + t
    y64 = F.linear(h, P["final_layer.weight"], P["final_layer.bias"])
    g = g * (~near_kink)[:, None]
    assert near_kink.float().mean().item() < 0.2
    y64.backward(g.double())

    old = config.linear_tc_min_rows
    try:
        config.linear_tc_min_rows = 1
        y = net(x, c)
        y.backward(g)
    finally:
        config.linear_tc_min_rows = old
    rel = lambda a, b: ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()
    assert rel(y, y64.detach()) < 1e-5
    assert rel(x.grad, x64.grad) < 2e-5
    for n, p in net.named_parameters():
        assert rel(p.grad, P[n].grad) < 2e-5, n


# ---- the sizes bench.py runs: per-GPU batch 256, multi-wave grids, split-K weight gradient ----
BENCH_CASES = [(256, 100, 100, 32, 3), (256, 100, 100, 4, 3), (256, 6, 100, 32, 1), (256, 100, 192, 32, 1), (256, 100, 1536, 4, 1),
               (256, 100, 100, 8, 3)]


@pytest.mark.parametrize("case", BENCH_CASES)
def test_conv_at_bench_batch_vs_fp64(cuda_device, case):
    """conv_tc forward, data gradient and weight / bias gradient at batch 256 against fp64 convolutions on the same GPU."""
    from manifold_flow.b200 import ops

    b, ci, co, hw, k = case
    g = torch.Generator(device="cuda").manual_seed(b + ci + co + hw)
    x = torch.randn(b, ci, hw, hw, generator=g, device=cuda_device)
    w = torch.randn(co, ci, k, k, generator=g, device=cuda_device) * (1.0 / (ci * k * k) ** 0.5)
    bias = torch.randn(co, generator=g, device=cuda_device)
    gy = torch.randn(b, co, hw, hw, generator=g, device=cuda_device)

    def run(dtype, fn):
        leaves = [t.detach().clone().to(dtype).requires_grad_(True) for t in (x, w, bias)]
        y = fn(*leaves)
        y.backward(gy.to(dtype))
        return [y.detach()] + [t.grad for t in leaves]

    ref = run(torch.float64, lambda xx, ww, bb: F.conv2d(F.relu(xx), ww, bb, padding=k // 2))
    got = run(torch.float32, lambda xx, ww, bb: ops.conv_tc(xx, ww, bb, relu_in=True))
    # y / grad_x reduce over <= 900 terms; grad_w / grad_b over up to 262144 pixels, ~5350 of them per split-K CTA in ONE
    # fp32 tensor-memory accumulator (the tensor core truncates when it accumulates, so the error grows with the length of
    # the chain: 3.7e-5 measured for the level-0 3x3 at batch 256; SURVEY appendix C asks 1e-3 for parameter gradients)
    for name, a, r in zip(["y", "grad_x", "grad_w", "grad_b"], got, ref):
        tol = 1e-4 if name in ("grad_w", "grad_b") else 2e-5
        assert _rel(a, r) <= tol, f"{name}: rel L2 error {_rel(a, r):.2e}"
        assert float((a.double() - r).abs().max()) <= 10 * tol * float(r.abs().max()), name


@pytest.mark.gpu
@pytest.mark.parametrize("rows,d_full,n_out,ctx,even", [(100, 40, 640, None, True), (256, 14, 224, 2, False), (256, 64, 1024, 1, True),
                                                         (1000, 3, 23, None, False), (5, 2, 29, 1, True), (2051, 40, 640, None, False)])
def test_fused_vector_conditioner_matches_fp64(cuda_device, rows, d_full, n_out, ctx, even):
    """mflow_mlp_forward / _backward (the whole ResidualNet of a coupling layer as one launch, reference
    nn/resnet.py:9-72) against an fp64 evaluation of the same formulas: outputs and every gradient, with the identity
    features read in place from a strided view of the coupling layer's input.  Rows with a pre-activation within 1e-4 of
    the ReLU kink get a zero incoming gradient (there the fp32 and fp64 masks may legitimately differ)."""
    from manifold_flow import nn as nn_
    from manifold_flow.b200 import cabi, ops

    torch.manual_seed(rows + n_out)
    full = torch.randn(rows, d_full, device=cuda_device, requires_grad=True)
    x = full[:, (1 if even else 0)::2]            # the identity features of an alternating mask: a strided view
    n_in = x.shape[1]
    net = nn_.ResidualNet(n_in, n_out, hidden_features=100, context_features=ctx, num_blocks=2).to(cuda_device)
    for p in net.parameters():
        p.data.normal_(0.0, 0.2)
    c = None if ctx is None else torch.randn(rows, ctx, device=cuda_device)
    g = torch.randn(rows, n_out, device=cuda_device)

    P = {n: p.detach().double().requires_grad_() for n, p in net.named_parameters()}
    x64 = x.detach().double().requires_grad_()
    c64 = None if c is None else c.double()
    h = F.linear(x64 if c is None else torch.cat((x64, c64), 1), P["initial_layer.weight"], P["initial_layer.bias"])
    near_kink = torch.zeros(rows, dtype=torch.bool, device=cuda_device)
    for i in range(2):
        near_kink |= h.detach().abs().min(dim=1).values < 1e-4
        t = F.linear(F.relu(h), P[f"blocks.{i}.linear_layers.0.weight"], P[f"blocks.{i}.linear_layers.0.bias"])
        near_kink |= t.detach().abs().min(dim=1).values < 1e-4
        t = F.linear(F.relu(t), P[f"blocks.{i}.linear_layers.1.weight"], P[f"blocks.{i}.linear_layers.1.bias"])
        if c is not None:
            t = t * torch.sigmoid(F.linear(c64, P[f"blocks.{i}.context_layer.weight"], P[f"blocks.{i}.context_layer.bias"]))
        h = h + t
    y64 = F.linear(h, P["final_layer.weight"], P["final_layer.bias"])
    g = g * (~near_kink)[:, None]
    y64.backward(g.double())

    n0 = cabi.launch_count
    y = net(x, c)
    assert cabi.launch_count - n0 == 1, "the conditioner must be ONE launch of libmflow_b200"
    y.backward(g)
    assert cabi.launch_count - n0 == 3
    rel = lambda a, b: ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()
    assert rel(y, y64.detach()) < 1e-5
    assert rel(full.grad[:, (1 if even else 0)::2], x64.grad) < 2e-5
    assert float(full.grad[:, (0 if even else 1)::2].abs().max()) == 0.0
    for n, p in net.named_parameters():
        assert rel(p.grad, P[n].grad) < 2e-5, n
    # deterministic: the row-ordered parameter-gradient reduction gives the same bits on a second pass
    first = {n: p.grad.clone() for n, p in net.named_parameters()}
    net.zero_grad(set_to_none=True)
    net(x, c).backward(g)
    assert all(torch.equal(first[n], p.grad) for n, p in net.named_parameters())
