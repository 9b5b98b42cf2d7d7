#!/usr/bin/env python
"""Benchmark of the manifold-flow hot path on B200 (contract: see the task statement / DESIGN.md).

Metric (BASELINE.json): M-flow log_prob + backward, samples/s, on the gan64d-shaped 64x64x3
configuration (cfg4: 4-level multi-scale RQ-coupling outer flow with 20 couplings, LU post-processing
on 3840 features, 64-D latent, 8-coupling conditional inner flow, 39.5 M parameters), mode
"mf-fixed-manifold", synthetic inputs x ~ U[0, 256), context ~ N(0, 1), seeded random-init weights
with data-dependent ActNorm initialisation.  One step = one batch through
``model(x, mode="mf-fixed-manifold", context=c)`` -> ``(-log_prob.mean()).backward()`` (+ gradient
all-reduce when N > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]

Prints ONE JSON line.  ``--impl reference`` times the reference algorithm's CPU restatement
(``oracle/``) on the host cores on a bounded sample (batch 25, the reference's own batch size).
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

WORKLOAD = "cfg4 gan64d-shaped 64x64x3 M-flow (39.5M params), mode=mf-fixed-manifold, log_prob + backward"
METRIC = "mflow_log_prob_backward_samples_per_sec"
REFERENCE_BATCH = 25  # experiments/configs/train_mf_gan64d_april.config: batchsize = 25


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (weak scaling)")
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default=None, choices=[None, "fp32", "tf32"], help="conditioner GEMM math (default: fp32 parity mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-timer", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying a captured CUDA graph (N=1)")
    return ap.parse_args()


def synthetic_batch(cfg, batch, seed, pin=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(batch, 3, 64, 64, generator=g) * 256.0 if cfg.startswith(("cfg4", "cfg5")) else torch.randn(batch, 40, generator=g)
    c = torch.randn(batch, 1, generator=g)
    if pin and torch.cuda.is_available():
        x, c = x.pin_memory(), c.pin_memory()
    return x, c


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


def oracle_step_fn(cfg, sd_cpu, batch, seed=11):
    """CPU restatement of the same step: (-log_prob.mean()).backward() through oracle/mflow_oracle.py."""
    import mflow_oracle as oracle

    model = oracle.config_model(cfg)
    leaves = {k: (v.clone().requires_grad_(True) if torch.is_floating_point(v) and k.split(".")[-1] not in ("_shift", "_scale") else v)
              for k, v in sd_cpu.items()}
    x, c = synthetic_batch(cfg, batch, seed, pin=False)

    def step():
        for v in leaves.values():
            if v.requires_grad:
                v.grad = None
        _, log_prob, _ = oracle.mflow_forward(model, leaves, x, "mf-fixed-manifold", c)
        loss = -log_prob.mean()
        loss.backward()
        return float(loss.detach())

    return step


def reference_state_dict(cfg):
    """Seeded weights + ActNorm statistics so the CPU arm evaluates a sane (initialised) model."""
    import golden_util as gu
    import mflow_oracle as oracle
    from manifold_flow.b200 import factory

    torch.manual_seed(gu.SEED)
    model = factory.create_model(cfg)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    # data-dependent ActNorm initialisation, done with the oracle on the CPU (normalization.py:134-147)
    x, c = synthetic_batch(cfg, 8, 5, pin=False)
    spec = oracle.config_model(cfg)
    _actnorm_init_walk(oracle, spec["outer"], sd, "outer_transform.", x)
    return sd


def _actnorm_init_walk(oracle, spec, sd, p, x):
    """Forward through the outer transform, initialising every ActNorm on the fly."""
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    sd = reference_state_dict(args.config)
    step = oracle_step_fn(args.config, sd, REFERENCE_BATCH)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = REFERENCE_BATCH * args.steps / dt
    sample = f"{args.steps} steps of batch {REFERENCE_BATCH} (the reference's batch size), {args.warmup} warm-up"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": REFERENCE_BATCH, "where": "host CPU, PyTorch ATen fp32 (oracle port of the reference path)"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_b200(args):
    from manifold_flow.b200 import cabi, config, ddp, factory
    import golden_util as gu

    if args.math:
        config.conditioner_math = args.math
    rank, world, local = ddp.init_from_env("nccl")
    if world != args.gpus and rank == 0 and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cabi.check(cabi.lib().mflow_check_device(), "mflow_check_device")

    torch.manual_seed(gu.SEED)
    model = factory.create_model(args.config).to(dev)
    model.train()
    b = args.batch
    # a small pool of distinct pinned host batches (rank-specific seeds: every rank sees other data)
    pool = [synthetic_batch(args.config, b, 1000 * rank + i) for i in range(4)]
    dev_pool = [(x.to(dev), c.to(dev)) for x, c in pool]
    with torch.no_grad():  # data-dependent ActNorm init (rank 0's statistics are broadcast)
        model(dev_pool[0][0], mode="projection", context=dev_pool[0][1])
    ddp.broadcast_parameters(model, src=0)
    use_graph = not args.no_graph
    # Multi-GPU: with a captured step the gradients are all-reduced after the replay (flat buckets, NCCL on the same
    # stream); the eager path overlaps the bucket all-reduces with backward through post-accumulate hooks.
    reducer = ddp.GradientAllReducer(model.parameters(), bucket_mb=64.0, overlap=not use_graph) if world > 1 else None

    def local_step(x, c):
        """forward + backward on this rank's shard (the part that is captured into a CUDA graph)"""
        if reducer is not None:
            reducer.zero_grad()
        else:
            model.zero_grad(set_to_none=True)
        log_prob = model(x, mode="mf-fixed-manifold", context=c)[1]
        loss = -log_prob.mean()
        loss.backward()
        return loss

    def step(x, c):
        loss = local_step(x, c)
        if reducer is not None:
            reducer.finish()
        return loss

    # ---- CUDA graph of the local step (forward + backward): ~4800 launches become one replay ----
    graph = None
    static_x = static_c = static_loss = None
    if use_graph:
        try:
            static_x, static_c = dev_pool[0][0].clone(), dev_pool[0][1].clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step(static_x, static_c)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if reducer is None:
                model.zero_grad(set_to_none=True)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = local_step(static_x, static_c)
            graph.replay()
            if reducer is not None:
                reducer.finish()
            torch.cuda.synchronize()
        except Exception as exc:  # capture is an optimisation: report and fall back to eager launches
            print(f"[bench] CUDA graph capture failed, running eagerly: {type(exc).__name__}: {str(exc)[:200]}", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
        if world > 1:  # all ranks must take the same path (the eager one needs the overlap hooks' bookkeeping off)
            ok = torch.tensor([1 if graph is not None else 0], device=dev)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
            if int(ok) == 0:
                graph = None

    def run_step(x, c, on_device):
        """One step on batch (x, c).  With a graph the batch is copied into the graph's static inputs
        (device->device for resident data, host->device for the end-to-end arm), the graph replayed and,
        on several GPUs, the gradient buckets all-reduced."""
        if graph is None:
            return step(x, c)
        static_x.copy_(x, non_blocking=True)
        static_c.copy_(c, non_blocking=True)
        graph.replay()
        if reducer is not None:
            reducer.finish()
        return static_loss

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, host_io):
        h2d = d2h = 0
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(n_steps):
            if host_io:
                xh, ch = pool[i % len(pool)]
                h2d = xh.numel() * 4 + ch.numel() * 4
                if graph is None:
                    loss = step(xh.to(dev, non_blocking=True), ch.to(dev, non_blocking=True))
                else:
                    loss = run_step(xh, ch, False)
                _ = loss.item()  # device -> host read of the step's result
                d2h = 4
            else:
                x, c = dev_pool[i % len(dev_pool)]
                run_step(x, c, True)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms) / 1e3, wall, h2d, d2h

    n_before = cabi.launch_count
    step(*dev_pool[1 % len(dev_pool)])
    launches_per_step_eager = cabi.launch_count - n_before  # libmflow_b200 kernels in one step (a graph replays the same)
    for i in range(max(args.warmup, 3)):
        run_step(*dev_pool[i % len(dev_pool)], True)
    # ---- device-resident timing (value) with clocks sampled during the region ----
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
    if rank == 0:
        sampler.start()
    launches0 = cabi.launch_count
    torch.cuda.nvtx.range_push("mflow_timed_region")  # ncu --nvtx --nvtx-include "mflow_timed_region/"
    secs, _, _, _ = timed(args.steps, host_io=False)
    torch.cuda.nvtx.range_pop()
    launches = cabi.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = world * b * args.steps / secs
    # ---- end-to-end timing through the public API with host buffers ----
    e2e_secs, _, h2d, d2h = timed(args.steps, host_io=True)
    e2e_value = world * b * args.steps / e2e_secs
    # ---- per-kernel device time of the spline launches (roofline) ----
    roof = None
    if not args.no_kernel_timer:
        cabi.kernel_timer = cabi.KernelTimer()
        for i in range(2):
            # eager, instrumented steps: give the host a head start so that the launch queue never runs dry -
            # otherwise the event pairs would also time the gaps in which the GPU waits for the next launch
            torch.cuda._sleep(int(0.2 * 1.9e9))
            step(*dev_pool[i % len(dev_pool)])
        summ = cabi.kernel_timer.summary()
        cabi.kernel_timer = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        tc_peak, tc_src = ((peaks["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json bf16_tflops_sustained: dense bf16 cuBLAS, kernel timed inside a long step)")
                           if "bf16_tflops_sustained" in peaks else (1400.0, "fallback"))
        step_ms = 1e3 * secs / args.steps
        rqs = {k: v for k, v in summ.items() if k.startswith("rqs_")}
        conv = {k: v for k, v in summ.items() if k.startswith("conv")}
        wgr = {k: v for k, v in summ.items() if k.startswith("wgrad")}

        def agg(group):
            return sum(v["bytes"] for v in group.values()), sum(v["ms"] for v in group.values())

        rb, rms = agg(rqs)
        spline = {"bound": "hbm", "kernel": "rqs_planes_* / rqs_rows_* (fused RQ-spline coupling fwd+inv+bwd, all launches of a step)",
                  "achieved": rb / (rms * 1e-3) / 1e9 if rms > 0 else None, "peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s",
                  "traffic": None, "share_of_step": (rms / 2) / step_ms,
                  "per_kernel": {k: {"launches_per_step": v["launches"] // 2, "ms_per_step": v["ms"] / 2,
                                     "algorithmic_GBps": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None} for k, v in rqs.items()}}
        spline["frac"] = spline["achieved"] / hbm_peak if spline["achieved"] else None
        if conv:  # the dominant kernel of the step: tcgen05 3xTF32 implicit-GEMM convolution (tensor-bound)
            cf, cms = agg(conv)  # "bytes" holds useful fp32-equivalent FLOPs for these tags
            ach = cf / (cms * 1e-3) / 1e12 if cms > 0 else None
            roof = {"bound": "tensor", "kernel": "conv_tf32x3_kernel (tcgen05 + TMA implicit-GEMM conv, 3xTF32 = 3 tensor passes per useful FLOP; all launches of a step)",
                    "achieved": ach, "peak": tc_peak, "peak_source": tc_src, "unit": "TFLOP/s", "frac": ach / tc_peak if ach else None,
                    "traffic": 167.44e6, "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one level-0 3x3 launch (B=256, 100->100, 32x32) from the ncu --set full capture in profiles/r1_conv_tf32x3_ncu_full.md; algorithmic 209.7e6 (part of the output is still in L2 when the kernel ends)",
                    "share_of_step": (cms / 2) / step_ms,
                    "tensor_pipe_active_ncu": 0.697, "tf32_passes_per_useful_flop": 3,
                    "note": "achieved counts useful fp32-equivalent FLOPs; the tensor pipe executes 3x that in TF32 (nominal TF32 peak is half the bf16 peak); tensor_pipe_active_ncu = sm__pipe_tensor_cycles_active of the level-0 3x3 launch (profiles/r1_conv_tf32x3_ncu_full.md)",
                    "per_kernel": {k: {"launches_per_step": v["launches"] // 2, "ms_per_step": v["ms"] / 2,
                                       "useful_TFLOPs": v["bytes"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else None} for k, v in conv.items()},
                    "secondary": spline}
            if wgr:  # the weight-gradient kernel of the same convolutions (pixels-as-K GEMM, same 3xTF32 scheme)
                wf, wms = agg(wgr)
                wach = wf / (wms * 1e-3) / 1e12 if wms > 0 else None
                roof["wgrad"] = {"bound": "tensor", "kernel": "wgrad_tf32x3_kernel (+ split-K reduce; all launches of a step)", "achieved": wach,
                                 "peak": tc_peak, "unit": "TFLOP/s", "frac": wach / tc_peak if wach else None, "traffic": None,
                                 "share_of_step": (wms / 2) / step_ms,
                                 "per_kernel": {k: {"launches_per_step": v["launches"] // 2, "ms_per_step": v["ms"] / 2,
                                                    "useful_TFLOPs": v["bytes"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else None}
                                                for k, v in wgr.items()}}
        else:
            roof = spline
    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        cstep = oracle_step_fn(args.config, sd_cpu, REFERENCE_BATCH)
        cstep()
        t0 = time.perf_counter()
        n = 0
        while n < 2 or (time.perf_counter() - t0 < 10.0 and n < 6):
            cstep()
            n += 1
        dt = time.perf_counter() - t0
        cpu_base = {"value": REFERENCE_BATCH * n / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": f"{n} steps of batch {REFERENCE_BATCH} of the same workload through oracle/mflow_oracle.py (PyTorch ATen CPU fp32), 1 warm-up"}
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if config.conditioner_math == "fp32" else "f32 (tf32 conditioner GEMMs)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * b, "per_gpu_batch": b, "parallelism": f"dp{world}",
                       "l2": "per-step working set (activations + spline parameters, several GB) exceeds the 126 MB L2; 4 input batches rotate",
                       "conditioner": ("image convolutions forward, data gradient and weight gradient: tcgen05 3xTF32 kernels of libmflow_b200; vector-conditioner GEMMs (256 rows: below the tensor-core threshold) and the 3840-wide LU layers: cuBLAS fp32 GEMMs (LU inverse: blocked substitution)"
                                       if config.conv_tensor_cores else "cuBLAS/cuDNN fp32 library calls")},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_secs / args.steps},
            "gpu_launches": launches if graph is None else launches_per_step_eager * args.steps,
            "cuda_graph": graph is not None, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
        }))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
