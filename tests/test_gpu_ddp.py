"""GPU, 2 ranks over NCCL (needs two devices: `gpurun --gpus 2`): the data-parallel step of the reduced gan64d-shaped
M-flow gives the gradients of one rank on the whole batch (<= 1e-6 relative up to reduction
order), eagerly (all-reduces overlapped with backward) and as a replayed CUDA graph followed by the bucket all-reduces."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _grads(model):
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      TORCH_NCCL_ASYNC_ERROR_HANDLING="0")
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(repo, "manifold-flow_b200"), os.path.join(repo, "oracle"), os.path.join(repo, "tests")):
        sys.path.insert(0, p)
    import golden_util as gu
    import model_util as mu
    from manifold_flow.b200 import ddp

    ddp.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    model, _, _ = mu.seeded_product_model("cfg4s")
    model.to(dev).train()
    ddp.broadcast_parameters(model, src=0)
    n = 32
    x, ctx = gu.config_inputs("cfg4s", batch=n)
    lo, hi = ddp.shard_batch(n, rank, world)
    xs, cs = x[lo:hi].to(dev), ctx[lo:hi].to(dev)
    reducer = ddp.GradientAllReducer(model.parameters(), bucket_mb=1.0, overlap=True)

    def step(a, c, reduce=True):
        reducer.zero_grad()
        lp = model(a, mode="mf-fixed-manifold", context=c)[1]
        (-lp.sum() / n).backward()          # global mean: every rank contributes its shard's share ...
        if reduce:
            reducer.finish()                # ... averaged over ranks -> multiply by world below

    # every backward runs on the side stream (one on the default stream before a capture would tie the parameters'
    # gradient accumulators to the legacy stream, which a capturing stream may not wait for)
    sx, sc = xs.clone(), cs.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step(sx, sc)                    # eager: bucket all-reduces issued from the backward hooks
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    eager = {k: g * world for k, g in _grads(model).items()}
    # what bench.py does: this rank's forward + backward replayed as one CUDA graph, bucket all-reduces after the replay
    # (no hooks: an all-reduce issued from a hook during capture would become a graph node)
    reducer.close()
    reducer = ddp.GradientAllReducer(model.parameters(), bucket_mb=1.0, overlap=False)
    with torch.cuda.stream(side):
        step(sx, sc)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step(sx, sc, reduce=False)
    graph.replay()
    reducer.finish()
    torch.cuda.synchronize()
    replay = {k: g * world for k, g in _grads(model).items()}
    if rank == 0:
        torch.save({"eager": {k: v.cpu() for k, v in eager.items()}, "replay": {k: v.cpu() for k, v in replay.items()},
                    "buckets": len(reducer.buckets)}, out)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_gradients_match_one_rank(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import golden_util as gu
    import model_util as mu

    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["buckets"] >= 2
    dev = torch.device("cuda", 0)
    model, _, _ = mu.seeded_product_model("cfg4s")
    model.to(dev).train()
    n = 32
    x, ctx = gu.config_inputs("cfg4s", batch=n)
    lp = model(x.to(dev), mode="mf-fixed-manifold", context=ctx.to(dev))[1]
    (-lp.sum() / n).backward()
    one = {k: p.grad.cpu() for k, p in model.named_parameters() if p.grad is not None}
    top = max(float(g.double().norm()) for g in one.values())
    for key in ("eager", "replay"):
        assert set(res[key]) == set(one)
        for k, g in one.items():
            err = float((res[key][k].double() - g.double()).norm())
            assert err <= 1e-6 * float(g.double().norm()) + 1e-7 * top, f"{key} {k}: {err:.3e} vs |g| {float(g.norm()):.3e}"
