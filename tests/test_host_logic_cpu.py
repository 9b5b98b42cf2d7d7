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
