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
