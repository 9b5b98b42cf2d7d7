"""CPU, world_size 2, gloo: the host-side data-parallel logic (batch sharding, flat gradient buckets,
overlapped all-reduce, parameter broadcast) gives the same gradients as one process on the whole batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, overlap, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "manifold-flow_b200"))
    from manifold_flow.b200 import ddp

    r, w, _ = ddp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)  # ranks start from DIFFERENT weights; broadcast must fix that
    net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    ddp.broadcast_parameters(net, src=0)
    reducer = ddp.GradientAllReducer(net.parameters(), bucket_mb=0.0002, overlap=overlap)  # tiny buckets -> several of them
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(12, 5, generator=g), torch.randn(12, 3, generator=g)
    for _ in range(2):  # second step checks zero_grad / hook re-arming
        lo, hi = ddp.shard_batch(12, rank, world)
        reducer.zero_grad()
        loss = ((net(x_all[lo:hi]) - y_all[lo:hi]) ** 2).mean()
        loss.backward()
        reducer.finish()
    grads = [p.grad.clone() for p in net.parameters()]
    weights = [p.detach().clone() for p in net.parameters()]
    if rank == 0:
        torch.save(dict(grads=grads, weights=weights, n_buckets=len(reducer.buckets), payload=reducer.payload_bytes()), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap", [True, False])
def test_gradient_allreduce_matches_single_process(tmp_path, overlap):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), overlap, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["n_buckets"] >= 2
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    for w, p in zip(res["weights"], net.parameters()):
        assert torch.equal(w, p.detach())  # rank 0's weights everywhere
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(12, 5, generator=g), torch.randn(12, 3, generator=g)
    ((net(x_all) - y_all) ** 2).mean().backward()
    for got, p in zip(res["grads"], net.parameters()):
        torch.testing.assert_close(got, p.grad, rtol=1e-6, atol=1e-7)
    assert res["payload"] == sum(p.numel() for p in net.parameters()) * 4


def test_shard_batch_covers_everything():
    import sys

    from manifold_flow.b200 import ddp

    for n, w in ((25, 8), (100, 3), (7, 8), (0, 2)):
        spans = [ddp.shard_batch(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def _worker_robust(rank, world, port, out):
    """(1) zero_grad(set_to_none=True) between steps must not stop the averaging; (2) gradient accumulation
    with overlap=False; (3) a second backward() with overlap=True fails loudly instead of reducing a half-filled bucket."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "manifold-flow_b200"))
    from manifold_flow.b200 import ddp

    ddp.init_from_env("gloo")
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(12, 5, generator=g), torch.randn(12, 3, generator=g)
    lo, hi = ddp.shard_batch(12, rank, world)
    loss_fn = lambda a, b: ((net(x_all[a:b]) - y_all[a:b]) ** 2).sum() / 12.0
    res = {}
    for overlap in (True, False):
        reducer = ddp.GradientAllReducer(net.parameters(), bucket_mb=0.0002, overlap=overlap)
        reducer.zero_grad()
        loss_fn(lo, hi).backward()
        reducer.finish()
        net.zero_grad(set_to_none=True)        # what optimizer.zero_grad() does by default: the bucket views are gone
        loss_fn(lo, hi).backward()             # hooks must re-attach the views and still average
        reducer.finish()
        res[f"none_{overlap}"] = [p.grad.clone() * world for p in net.parameters()]   # sum-of-shard-means * 1 = full gradient
        assert all(reducer._aliases(p) for p in net.parameters())
        reducer.close()
    # accumulation over two micro-batches, all-reduce once
    reducer = ddp.GradientAllReducer(net.parameters(), bucket_mb=0.0002, overlap=False)
    reducer.zero_grad()
    mid = (lo + hi) // 2
    loss_fn(lo, mid).backward()
    loss_fn(mid, hi).backward()
    reducer.finish()
    res["accum"] = [p.grad.clone() * world for p in net.parameters()]
    reducer.close()
    # second backward with overlap=True -> loud failure
    reducer = ddp.GradientAllReducer(net.parameters(), bucket_mb=0.0002, overlap=True)
    reducer.zero_grad()
    loss_fn(lo, mid).backward()
    try:
        loss_fn(mid, hi).backward()
        res["raised"] = False
    except RuntimeError as exc:
        res["raised"] = "second backward" in str(exc)
    reducer.finish()
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


def test_reducer_survives_set_to_none_and_accumulation(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker_robust, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(5, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(12, 5, generator=g), torch.randn(12, 3, generator=g)
    (((net(x_all) - y_all) ** 2).sum() / 12.0).backward()
    for key in ("none_True", "none_False", "accum"):
        for got, p in zip(res[key], net.parameters()):
            torch.testing.assert_close(got, p.grad, rtol=1e-5, atol=1e-7)
    assert res["raised"] is True
