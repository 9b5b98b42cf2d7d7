#!/usr/bin/env python
"""Benchmark of the manifold-flow hot path on B200 (contract: see the task statement / DESIGN.md).

Headline metric (BASELINE.json): M-flow log_prob + backward, samples/s, on the gan64d-shaped 64x64x3
configuration (cfg4: 4-level multi-scale RQ-coupling outer flow with 20 couplings, LU post-processing on
3840 features, 64-D latent, 8-coupling conditional inner flow, 39.5 M parameters), mode "mf-fixed-manifold",
synthetic inputs x ~ U[0, 256), context ~ N(0, 1), seeded random-init weights with data-dependent ActNorm
initialisation.  One step = one batch through ``model(x, mode=..., context=c)`` -> ``(-log_prob.mean()).backward()``
(+ gradient all-reduce of the trained parameter group when N > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg1..cfg5] [--batch B]
                    [--impl b200|reference|reference-gpu]

``--config`` selects the other BASELINE.json shapes: cfg1 / cfg3 vector M-flows
(log_prob + backward), cfg2 the full-Jacobian ``mode="mf"`` evaluation (experiments/evaluate.py:219-224), cfg5 the
ambient image flow (``Flow.log_prob`` + backward, flows/flow.py:53-63).  Prints ONE JSON line.
``--impl reference`` times the reference algorithm's CPU restatement (``oracle/``) on the host cores on a bounded
sample (the reference's own batch size); ``--impl reference-gpu`` runs the same restatement through PyTorch eager
on the GPU (cuDNN / cuBLAS / ATen: "the kernel to beat"), with TF32 on (torch's cuDNN default)
and off.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(REPO, "manifold-flow_b200"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

# kind: "train" = log_prob + backward of a ManifoldFlow, "train_flow" = the same for Flow, "eval_mf" = mode="mf" forward
CONFIGS = {
    "cfg1": dict(workload="cfg1 spherical_gaussian-shaped 3->2 vector M-flow (5+5 RQ couplings, K=8), mode=mf-fixed-manifold, log_prob + backward",
                 metric="mflow_log_prob_backward_samples_per_sec", kind="train", mode="mf-fixed-manifold", dim=3, ctx=0, batch=1 << 20, ref_batch=100,
                 trained="inner_transform"),
    "cfg2": dict(workload="cfg2 power-shaped 3->2 vector M-flow (K=10, context 1), mode=mf full-Jacobian log-likelihood evaluation (forward only; the CPU port batches "
                          "the Jacobian as D reverse passes per layer - the reference itself loops over B*D autograd calls and is ~16x slower, BASELINE.md section 4)",
                 metric="mflow_mf_eval_samples_per_sec", kind="eval_mf", mode="mf", dim=3, ctx=1, batch=1 << 18, ref_batch=100, trained=None),
    "cfg3": dict(workload="cfg3 lhc40d-shaped 40->14 conditional vector M-flow (20+15 RQ couplings, LU linear, K=11), mode=mf-fixed-manifold, log_prob + backward",
                 metric="mflow_log_prob_backward_samples_per_sec", kind="train", mode="mf-fixed-manifold", dim=40, ctx=2, batch=1 << 16, ref_batch=100,
                 trained="inner_transform"),
    "cfg4": dict(workload="cfg4 gan64d-shaped 64x64x3 M-flow (39.5M params), mode=mf-fixed-manifold, log_prob + backward",
                 metric="mflow_log_prob_backward_samples_per_sec", kind="train", mode="mf-fixed-manifold", dim=None, ctx=1, batch=256, ref_batch=25,
                 trained="inner_transform"),
    "cfg5": dict(workload="cfg5 celeba-shaped 64x64x3 ambient flow (28 RQ couplings, 12.3M params), Flow.log_prob + backward",
                 metric="flow_log_prob_backward_samples_per_sec", kind="train_flow", mode=None, dim=None, ctx=0, batch=256, ref_batch=25, trained="all"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (weak scaling); default: the config's")
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--math", default=None, choices=[None, "fp32", "tf32"], help="conditioner GEMM math (default: fp32 parity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timer", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the replayed step's log_prob")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying a captured CUDA graph")
    ap.add_argument("--reduce", default="trained", choices=["trained", "all"],
                    help="N > 1: all-reduce the gradients of the parameter group the phase trains (experiments/train.py:287-297) or of all parameters")
    ap.add_argument("--allreduce", default="after-replay", choices=["after-replay", "in-graph"],
                    help="N > 1 with a captured step: all-reduce the gradient buckets right after the replay, on the same stream (default), or from "
                         "backward hooks inside the captured graph (experimental: NCCL kernels become graph nodes)")
    ap.add_argument("--detail", default=None, help="write the per-kernel breakdown (JSON) to this path")
    return ap.parse_args()


def synthetic_batch(cfg, batch, seed, pin=True):
    """Synthetic inputs of SURVEY.md 8(d): images x ~ U[0, 256), vectors x ~ N(0, 1); context ~ N(0, 1) (cfg2: U(-1, 1))."""
    c = CONFIGS[cfg]
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, 64, 64, generator=g) * 256.0 if c["dim"] is None else torch.randn(batch, c["dim"], generator=g)
    ctx = None
    if c["ctx"]:
        ctx = torch.rand(batch, c["ctx"], generator=g) * 2 - 1 if cfg == "cfg2" else torch.randn(batch, c["ctx"], generator=g)
    if pin and torch.cuda.is_available():
        x, ctx = x.pin_memory(), (None if ctx is None else ctx.pin_memory())
    return x, ctx


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# the reference algorithm's restatement (oracle/) - CPU baseline, reference arms, parity checker
# ---------------------------------------------------------------------------------------------------------
def _oracle_log_prob(oracle, cfg, model, sd, x, c):
    kind = CONFIGS[cfg]["kind"]
    if kind == "train_flow":
        return oracle.flow_forward(model, sd, x, c, decode=False)[1]
    if kind == "eval_mf":
        return oracle.mflow_log_prob_mf(model, sd, x, c)[1]
    return oracle.mflow_forward(model, sd, x, CONFIGS[cfg]["mode"], c)[1]


def oracle_step_fn(cfg, sd, batch, seed=11, device="cpu"):
    """The same step through oracle/mflow_oracle.py (PyTorch ATen ops in the reference's order) on ``device``."""
    import mflow_oracle as oracle

    model = oracle.config_model(cfg)
    sd = {k: v.to(device) for k, v in sd.items()}
    train = CONFIGS[cfg]["kind"] != "eval_mf"
    leaves = {k: (v.clone().requires_grad_(True) if train and torch.is_floating_point(v) and k.split(".")[-1] not in ("_shift", "_scale") else v)
              for k, v in sd.items()}
    x, c = synthetic_batch(cfg, batch, seed, pin=False)
    x, c = x.to(device), (None if c is None else c.to(device))

    def step():
        if not train:
            return float(_oracle_log_prob(oracle, cfg, model, leaves, x, c).mean())
        for v in leaves.values():
            if v.requires_grad:
                v.grad = None
        loss = -_oracle_log_prob(oracle, cfg, model, leaves, x, c).mean()
        loss.backward()
        return float(loss.detach())

    return step


def reference_state_dict(cfg):
    """Seeded weights + ActNorm statistics so the reference arms evaluate a sane (initialised) model."""
    import golden_util as gu
    import mflow_oracle as oracle
    from manifold_flow.b200 import factory

    torch.manual_seed(gu.SEED)
    model = factory.create_model(cfg)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    spec = oracle.config_model(cfg)
    if spec["kind"] in ("image_mf", "image_flow"):
        # data-dependent ActNorm initialisation, done with the oracle on the CPU (normalization.py:134-147)
        x, _ = synthetic_batch(cfg, 8, 5, pin=False)
        root, prefix = (spec["outer"], "outer_transform.") if spec["kind"] == "image_mf" else (spec["transform"], "transform.")
        _actnorm_init_walk(oracle, root, sd, prefix, x)
    return sd


def _actnorm_init_walk(oracle, spec, sd, p, x):
    """Forward through the transform, initialising every ActNorm on the fly."""
    kind = spec[0]
    if kind == "composite":
        for i, child in enumerate(spec[1]):
            x = _actnorm_init_walk(oracle, child, sd, f"{p}_transforms.{i}.", x)
        return x
    if kind == "multiscale":
        _, children, _ = spec
        hidden, outs = x, []
        for i, child in enumerate(children):
            y = _actnorm_init_walk(oracle, child, sd, f"{p}_transforms.{i}.", hidden)
            if i < len(children) - 1:
                out, hidden = torch.chunk(y, 2, dim=1)
            else:
                out = y
            outs.append(out.reshape(x.shape[0], -1))
        return torch.cat(outs, dim=-1)
    if kind == "actnorm":
        ls, sh = oracle.actnorm_init(x)
        sd[p + "log_scale"], sd[p + "shift"] = ls, sh
    with torch.no_grad():
        return oracle.apply(spec, sd, p, x, None, False)[0]


def _time_cpu(step, warmup, steps):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return time.perf_counter() - t0


def run_reference(args):
    """Reference arm: the oracle port on the box's host cores (the reference is pure Python and cannot travel)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg, c = args.config, CONFIGS[args.config]
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    b = c["ref_batch"]
    step = oracle_step_fn(cfg, reference_state_dict(cfg), b)
    dt = _time_cpu(step, args.warmup, args.steps)
    value = b * args.steps / dt
    sample = f"{args.steps} steps of batch {b} (the reference's batch size), {args.warmup} warm-up"
    print(json.dumps({
        "impl": "reference", "metric": c["metric"], "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": c["workload"], "global_batch": b, "where": "host CPU, PyTorch ATen fp32 (oracle port of the reference path)"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_reference_gpu(args):
    """The reference algorithm through PyTorch eager on ONE B200 (cuDNN convolutions, cuBLAS GEMMs, ~60 ATen launches per
    spline call): the "kernel to beat" of SURVEY.md 8(d).  Two settings: TF32 allowed (torch's default for cuDNN
    convolutions, i.e. what the reference really runs) and TF32 off (fp32 parity mode, what this repo's kernels match)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg, c = args.config, CONFIGS[args.config]
    dev = torch.device("cuda", 0)
    sd = reference_state_dict(cfg)
    b = args.batch or c["ref_batch"]
    out = {}
    for label, tf32 in (("tf32_on", True), ("tf32_off", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        step = oracle_step_fn(cfg, sd, b, device=dev)
        for _ in range(max(args.warmup, 2)):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        secs = e0.elapsed_time(e1) / 1e3
        out[label] = {"samples_per_s": b * args.steps / secs, "ms_per_step": 1e3 * secs / args.steps}
    print(json.dumps({
        "impl": "reference-gpu", "metric": c["metric"], "value": out["tf32_on"]["samples_per_s"], "unit": "samples/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 2), "ms_per_step": out["tf32_on"]["ms_per_step"], "higher_is_better": True,
        "dtype": "f32 (cuDNN/cuBLAS TF32 allowed)", "data": "synthetic",
        "config": {"workload": c["workload"], "global_batch": b,
                   "where": "one B200, PyTorch eager (oracle port of the reference path: cuDNN / cuBLAS / ATen), loss.item() per step"},
        "tf32_on": out["tf32_on"], "tf32_off": out["tf32_off"],
    }))


# ---------------------------------------------------------------------------------------------------------
# measured peaks
# ---------------------------------------------------------------------------------------------------------
def measure_matmul_peak(dev, tf32):
    """Dense 8192^3 torch.matmul (cuBLAS) the way MEASURED_PEAKS.json's bf16 figure was taken: best of 10, CUDA events.
    tf32=True: fp32 operands with TF32 tensor-core math; tf32=False: fp16 operands."""
    n = 8192
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        dt = torch.float32 if tf32 else torch.float16
        a, b = torch.randn(n, n, device=dev, dtype=dt), torch.randn(n, n, device=dev, dtype=dt)
        out = torch.empty(n, n, device=dev, dtype=dt)
        for _ in range(3):
            torch.matmul(a, b, out=out)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=out)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def load_json(path):
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


# ---------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------
def _trace(msg):
    if os.environ.get("MFLOW_BENCH_TRACE"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def run_b200(args):
    from manifold_flow.b200 import cabi, config, ddp, factory
    import golden_util as gu

    cfg, C = args.config, CONFIGS[args.config]
    if args.math:
        config.conditioner_math = args.math
    os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")  # collectives are captured into the step's CUDA graph
    rank, world, local = ddp.init_from_env("nccl")
    if world != args.gpus and rank == 0 and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cabi.check(cabi.lib().mflow_check_device(), "mflow_check_device")

    torch.manual_seed(gu.SEED)
    model = factory.create_model(cfg).to(dev)
    kind = C["kind"]
    train = kind != "eval_mf"
    model.train(train)
    b = args.batch or C["batch"]
    # a small pool of distinct pinned host batches (rank-specific seeds: every rank sees other data)
    pool = [synthetic_batch(cfg, b, 1000 * rank + i) for i in range(4)]
    to_dev = lambda t: None if t is None else t.to(dev)
    dev_pool = [(to_dev(x), to_dev(c)) for x, c in pool]

    def forward(x, c):
        if kind == "train_flow":
            return model.log_prob(x, context=c)
        return model(x, mode=C["mode"], context=c)[1]

    if train:
        with torch.no_grad():  # data-dependent ActNorm init (rank 0's statistics are broadcast)
            model.train()
            if kind == "train_flow":
                model.log_prob(dev_pool[0][0], context=dev_pool[0][1])
            else:
                model(dev_pool[0][0], mode="projection", context=dev_pool[0][1])
    _trace("actnorm init done")
    ddp.broadcast_parameters(model, src=0)
    _trace("parameters broadcast")
    use_graph = not args.no_graph
    # Multi-GPU: the flat gradient buckets of the TRAINED parameter group (experiments/train.py:287-297) are all-reduced
    # with NCCL, averaged in the reduction.  Eager steps and --allreduce in-graph: from the hooks of backward, overlapped
    # with the rest of backward.  Captured step (default): the graph holds this rank's forward + backward, the bucket
    # all-reduces follow the replay on the same stream.
    reducer = None
    reduced_bytes = 0
    in_graph = args.allreduce == "in-graph"
    if world > 1 and train:
        group = model.parameters() if (args.reduce == "all" or C["trained"] == "all") else getattr(model, C["trained"]).parameters()
        reducer = ddp.GradientAllReducer(group, bucket_mb=64.0, overlap=(not use_graph) or in_graph)
        reduced_bytes = reducer.payload_bytes()

    def step(x, c, reduce=True):
        """one whole step on this rank's shard: forward + backward (+ bucket all-reduces)"""
        if not train:
            with torch.no_grad():
                lp = forward(x, c)
            return lp.mean(), lp
        if reducer is not None:
            reducer.zero_grad()
            for p in model.parameters():   # parameters outside the reduced group
                if id(p) not in reducer._views:
                    p.grad = None
        else:
            model.zero_grad(set_to_none=True)
        lp = forward(x, c)
        loss = -lp.mean()
        loss.backward()
        if reducer is not None and reduce:
            reducer.finish()
        return loss, lp.detach()

    # ---- CUDA graph of the step: thousands of launches (and the NCCL kernels) become one replay ----
    graph = None
    static_x = static_c = static_loss = static_lp = None
    if use_graph:
        try:
            static_x = dev_pool[0][0].clone()
            static_c = None if dev_pool[0][1] is None else dev_pool[0][1].clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for it in range(3 if world == 1 else 11):
                    step(static_x, static_c)
                    _trace(f"warm-up step {it} issued")
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            _trace("warm-up done")
            if reducer is None:
                model.zero_grad(set_to_none=True)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                static_loss, static_lp = step(static_x, static_c, reduce=in_graph)
            _trace("captured")
            graph.replay()
            if reducer is not None and not in_graph:
                reducer.finish()
            torch.cuda.synchronize()
            _trace("first replay done")
        except Exception as exc:  # capture is an optimisation: report and fall back to eager launches
            print(f"[bench] CUDA graph capture failed, running eagerly: {type(exc).__name__}: {str(exc)[:300]}", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
        if world > 1:  # all ranks must take the same path
            ok = torch.tensor([1 if graph is not None else 0], device=dev)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
            if int(ok) == 0:
                graph = None

    def run_step(x, c):
        """One step on batch (x, c).  With a graph the batch is copied into the graph's static inputs (device->device
        for resident data, host->device for the end-to-end arm) and the graph replayed."""
        if graph is None:
            return step(x, c)
        static_x.copy_(x, non_blocking=True)
        if static_c is not None:
            static_c.copy_(c, non_blocking=True)
        graph.replay()
        if reducer is not None and not in_graph:
            reducer.finish()
        return static_loss, static_lp

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, host_io):
        h2d = d2h = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            if host_io:
                xh, ch = pool[i % len(pool)]
                h2d = xh.numel() * 4 + (0 if ch is None else ch.numel() * 4)
                if graph is None:
                    loss, _ = step(xh.to(dev, non_blocking=True), None if ch is None else ch.to(dev, non_blocking=True))
                else:
                    loss, _ = run_step(xh, ch)
                _ = loss.item()  # device -> host read of the step's result
                d2h = 4
            else:
                run_step(*dev_pool[i % len(dev_pool)])
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms) / 1e3, h2d, d2h

    _trace("graph phase over")
    n_before = cabi.launch_count
    step(*dev_pool[1 % len(dev_pool)])
    _trace("eager step done")
    launches_per_step_eager = cabi.launch_count - n_before  # libmflow_b200 kernels in one step (a graph replays the same)
    for i in range(max(args.warmup, 3)):
        run_step(*dev_pool[i % len(dev_pool)])
    # ---- device-resident timing (value) with clocks sampled during the region ----
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
    if rank == 0:
        sampler.start()
    launches0 = cabi.launch_count
    torch.cuda.nvtx.range_push("mflow_timed_region")  # ncu --nvtx --nvtx-include "mflow_timed_region/"
    secs, _, _ = timed(args.steps, host_io=False)
    torch.cuda.nvtx.range_pop()
    launches = cabi.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = world * b * args.steps / secs
    # ---- end-to-end timing through the public API with host buffers ----
    e2e_secs, h2d, d2h = timed(args.steps, host_io=True)
    e2e_value = world * b * args.steps / e2e_secs

    # ---- parity of the path that was just timed: log_prob of a replayed batch against the oracle on the CPU ----
    parity = None
    if not args.no_parity:
        _, lp_chk = run_step(dev_pool[2][0], dev_pool[2][1])   # every rank: the step may contain collectives
        torch.cuda.synchronize()
    if rank == 0 and not args.no_parity:
        import mflow_oracle as oracle

        n_chk = min(b, 8 if C["dim"] is None else (64 if kind == "eval_mf" else 256))
        xh, ch = pool[2]
        got = lp_chk[:n_chk].float().cpu()
        sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        sd64 = {k: (v.double() if torch.is_floating_point(v) else v) for k, v in sd_cpu.items()}
        om = oracle.config_model(cfg)
        xc, cc = xh[:n_chk].clone(), (None if ch is None else ch[:n_chk].clone())
        with torch.set_grad_enabled(kind == "eval_mf"):
            ref = _oracle_log_prob(oracle, cfg, om, sd_cpu, xc, cc).detach()
            truth = _oracle_log_prob(oracle, cfg, om, sd64, xc.double(), None if cc is None else cc.double()).detach()
        # Criterion of tests/test_gpu_models.py: within rtol / atol of the reference's fp32 result (_lp_tol), or - on samples
        # where the reference's own fp32 arithmetic is ill-conditioned (the PIE term divides by epsilon^2 = 1e-4, the
        # quadratic-root inverse) - no further from the fp64 result than 5x the reference's fp32 error
        rtol, atol = (1e-5, 5e-3) if C["dim"] is None else (2e-5, 2e-4)
        err = (got.double() - ref.double()).abs()
        bound = atol + rtol * ref.double().abs()
        ref_err = (ref.double() - truth).abs()
        noise = float(ref_err.kthvalue(max(1, int(0.9 * n_chk)))[0])   # batch-level fp32 noise floor of the reference's own arithmetic
        budget = 5 * ref_err + 3 * noise + 5e-6 * truth.abs() + atol
        good = (err <= bound) | ((got.double() - truth).abs() <= budget)
        parity = {"checked": True, "ok": bool(good.all()), "samples": n_chk, "outside_rtol_atol": int((err > bound).sum()),
                  "violations": int((~good).sum()), "max_abs_err": float(err.max()),
                  "max_err_over_bound": float((err / bound).max()), "max_abs_log_prob": float(ref.abs().max()),
                  "ref_fp32_vs_fp64_max_abs": float((ref.double() - truth).abs().max()), "rtol": rtol, "atol": atol,
                  "what": "log_prob of the first samples of a batch run through the timed (graph-replayed) step vs oracle/mflow_oracle.py on the CPU (fp32, and fp64 as the truth)"}

    # ---- per-kernel device time (roofline) ----
    peaks = load_json(os.path.join(REPO, "MEASURED_PEAKS.json"))
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    tc_peak, tc_src = ((peaks["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json bf16_tflops_sustained)")
                       if "bf16_tflops_sustained" in peaks else (1400.0, "fallback"))
    ncu = load_json(os.path.join(REPO, "profiles", "r2_ncu_metrics.json"))   # written by tools/ncu_summary.py from the committed captures
    roof = roof_hbm = detail = None
    step_ms = 1e3 * secs / args.steps
    if not args.no_kernel_timer:
        cabi.kernel_timer = cabi.KernelTimer()
        for i in range(2):
            # eager, instrumented steps: give the host a head start so that the launch queue never runs dry -
            # otherwise the event pairs would also time the gaps in which the GPU waits for the next launch
            torch.cuda._sleep(int(0.6 * 1.9e9))
            step(*dev_pool[i % len(dev_pool)])
        summ = cabi.kernel_timer.summary()
        cabi.kernel_timer = None

        def group(prefixes):
            sel = {k: v for k, v in summ.items() if k.startswith(prefixes)}
            return sel, sum(v["bytes"] for v in sel.values()), sum(v["ms"] for v in sel.values())

        def per_kernel(sel, unit, scale):
            return {k: {"launches_per_step": v["launches"] // 2, "ms_per_step": v["ms"] / 2,
                        unit: v["bytes"] / (v["ms"] * 1e-3) / scale if v["ms"] > 0 else None} for k, v in sel.items()}

        rqs, rb, rms = group(("rqs_",))
        conv, cf, cms = group(("conv",))
        wgr, wf, wms = group(("wgrad",))
        detail = {"spline": per_kernel(rqs, "algorithmic_GBps", 1e9), "conv": per_kernel(conv, "useful_TFLOPs", 1e12),
                  "wgrad": per_kernel(wgr, "useful_TFLOPs", 1e12)}
        if rms > 0:
            ach = rb / (rms * 1e-3) / 1e9
            best = max((v["bytes"] / (v["ms"] * 1e-3) / 1e9, k) for k, v in rqs.items() if v["ms"] > 0)
            roof_hbm = {"bound": "hbm", "kernel": "rqs_planes_* / rqs_rows_* (fused RQ-spline coupling fwd + inv + bwd; all launches of a step)",
                        "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "peak_source": hbm_src,
                        "traffic": ncu.get("rqs_planes_apply", {}).get("dram_bytes"), "share_of_step": (rms / 2) / step_ms,
                        "best_launch": {"kernel": best[1], "achieved": best[0], "frac": best[0] / hbm_peak}}
        if cms > 0:  # the dominant kernel of the image configs: tcgen05 implicit-GEMM convolution
            ach = cf / (cms * 1e-3) / 1e12
            fmt = config.conv_operand_format
            best = max((v["bytes"] / (v["ms"] * 1e-3) / 1e12, k) for k, v in conv.items() if v["ms"] > 0)
            roof = {"bound": "tensor",
                    "kernel": f"conv_split3_kernel<{fmt}> of libmflow_b200 (tcgen05 + TMA implicit-GEMM convolution, error-compensated {fmt} operand split: "
                              "3 tensor-core products per useful FLOP; all launches of a step)",
                    "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s", "frac": ach / tc_peak, "peak_source": tc_src,
                    "traffic": ncu.get("conv", {}).get("dram_bytes"),
                    "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE level-0 3x3 launch (B=256, 100->100, 32x32; algorithmic 209.7e6) "
                                    "from the ncu --set full capture summarised in profiles/r2_ncu_summary.md",
                    "share_of_step": (cms / 2) / step_ms, "passes_per_useful_flop": 3,
                    "tensor_work": {"achieved": 3 * ach, "unit": "TFLOP/s of tensor-core products", "frac_of_peak": 3 * ach / tc_peak,
                                    "note": "what the tensor pipe executes: 3 products per useful FLOP (padding of C 100 -> 112 / K to 16 not counted)"},
                    "best_launch": {"kernel": best[1], "achieved": best[0], "frac": best[0] / tc_peak, "tensor_work_frac": 3 * best[0] / tc_peak},
                    "tensor_pipe_active_ncu": ncu.get("conv", {}).get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    "note": "achieved counts useful fp32-equivalent FLOPs; tensor_pipe_active_ncu = sm__pipe_tensor_cycles_active (% of active cycles) of the level-0 "
                            "3x3 launch in the committed ncu capture"}
            if wms > 0:
                wach = wf / (wms * 1e-3) / 1e12
                roof["wgrad"] = {"achieved": wach, "frac": wach / tc_peak, "share_of_step": (wms / 2) / step_ms}
        else:
            roof = roof_hbm
    if rank == 0 and roof is not None and roof.get("bound") == "tensor":
        # the peaks of the tensor-core input types this kernel could use, measured the way MEASURED_PEAKS.json was
        try:
            tf32_peak, fp16_peak = measure_matmul_peak(dev, True), measure_matmul_peak(dev, False)
            roof["tf32_peak_measured"], roof["frac_vs_tf32_peak"] = tf32_peak, roof["achieved"] / tf32_peak
            roof["fp16_peak_measured"], roof["frac_vs_fp16_peak"] = fp16_peak, roof["achieved"] / fp16_peak
        except RuntimeError as exc:
            roof["tf32_peak_measured"] = f"unavailable: {str(exc)[:80]}"
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        rb_ = C["ref_batch"]
        cstep = oracle_step_fn(cfg, sd_cpu, rb_)
        cstep()
        t0 = time.perf_counter()
        n = 0
        while n < 2 or (time.perf_counter() - t0 < 10.0 and n < 50):
            cstep()
            n += 1
        dt = time.perf_counter() - t0
        cpu_base = {"value": rb_ * n / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": f"{n} steps of batch {rb_} of the same workload through oracle/mflow_oracle.py (PyTorch ATen CPU fp32), 1 warm-up"}
    if rank == 0:
        if detail is not None:
            print("[bench] per-kernel detail: " + json.dumps(detail), file=sys.stderr)
            if args.detail:
                with open(args.detail, "w") as f:
                    json.dump(detail, f, indent=1)
        working_set_note = ("per-step working set (activations + spline parameters, several GB) exceeds the 126 MB L2; 4 input batches rotate"
                            if b >= 4096 or C["dim"] is None else
                            "batch of the reference's size: the working set fits the L2 and the step is launch-latency bound; 4 input batches rotate, no L2 flush")
        line = {
            "metric": C["metric"], "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if config.conditioner_math == "fp32" else "f32 (tf32 conditioner GEMMs)", "data": "synthetic",
            "config": {"workload": C["workload"], "name": cfg, "global_batch": world * b, "per_gpu_batch": b, "parallelism": f"dp{world}",
                       "l2": working_set_note,
                       "allreduce": (f"{reduced_bytes} B per step: gradients of the trained group ({C['trained'] if args.reduce == 'trained' else 'all'}), NCCL AVG, "
                                     + ("issued from backward hooks inside the captured graph" if (in_graph and graph is not None) else
                                        "after the graph replay on the same stream" if graph is not None else "from backward hooks (eager)")
                                     if reducer is not None else None)},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_secs / args.steps},
            "gpu_launches": launches if graph is None else launches_per_step_eager * args.steps,
            "cuda_graph": graph is not None, "clocks": clocks, "roofline": roof, "roofline_hbm": roof_hbm, "cpu_baseline": cpu_base,
            "parity_checked": bool(parity and parity["checked"]), "parity": parity,
        }
        if parity is not None and not parity["ok"]:
            line["invalid"] = "parity check of the timed path failed"
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
        if graph is not None and in_graph:
            graph.reset()   # the captured NCCL kernels reference the communicator: release the graph before tearing it down
        torch.distributed.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        sys.exit(3)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "reference-gpu":
        run_reference_gpu(a)
    else:
        run_b200(a)
