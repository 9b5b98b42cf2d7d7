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
    reducer = ddp.GradientAllReducer(model.parameters(), bucket_mb=64.0) if world > 1 else None

    def step(x, c):
        if reducer is not None:
            reducer.zero_grad()
        else:
            model.zero_grad(set_to_none=True)
        log_prob = model(x, mode="mf-fixed-manifold", context=c)[1]
        loss = -log_prob.mean()
        loss.backward()
        if reducer is not None:
            reducer.finish()
        return loss

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
                x, c = xh.to(dev, non_blocking=True), ch.to(dev, non_blocking=True)
                h2d = xh.numel() * 4 + ch.numel() * 4
                loss = step(x, c)
                _ = loss.item()  # device -> host read of the step's result
                d2h = 4
            else:
                x, c = dev_pool[i % len(dev_pool)]
                step(x, c)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms) / 1e3, wall, h2d, d2h

    for i in range(max(args.warmup, 3)):
        step(*dev_pool[i % len(dev_pool)])
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
            step(*dev_pool[i % len(dev_pool)])
        summ = cabi.kernel_timer.summary()
        cabi.kernel_timer = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak, which = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback")
        tot_b = sum(v["bytes"] for v in summ.values())
        tot_ms = sum(v["ms"] for v in summ.values())
        achieved = tot_b / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else None
        roof = {"bound": "hbm", "kernel": "rqs_planes_* / rqs_rows_* (fused RQ-spline coupling fwd+inv+bwd, all launches of a step)",
                "achieved": achieved, "peak": peak, "peak_source": which, "unit": "GB/s", "frac": (achieved / peak if achieved else None),
                "traffic": None,
                "per_kernel": {k: {"launches_per_step": v["launches"] // 2, "ms_per_step": v["ms"] / 2,
                                   "algorithmic_GBps": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None} for k, v in summ.items()},
                "share_of_step": (tot_ms / 2) / (1e3 * secs / args.steps)}
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
                       "conditioner": "cuBLAS/cuDNN fp32 library calls (round 1); spline/glue/epilogue kernels: libmflow_b200"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_secs / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
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
