"""Synchronisation-free training step.

The reference's ``Trainer.batch_train`` (experiments/training/trainer.py:512-521) runs, per batch: the forward pass with
three or four ``torch.isnan(...).any()`` host synchronisations (:108-122, :586-588), ``optimizer.zero_grad()``,
``loss.backward()``, ``clip_grad_norm_``, ``optimizer.step()`` (:177-184) and one ``.item()`` per loss term (:518-520),
plus D2H copies of the whole batch (:589-595).  With the fused kernels of this package a step of the vector models
is a few hundred launches of a few microseconds each, so those synchronisations and the launch latency are what is left.

``GraphedTrainStep`` captures forward + loss + backward + gradient clipping + optimizer step (+ the gradient
all-reduce hand-off of ``ddp.GradientAllReducer``) into ONE CUDA graph and replays it: no host synchronisation inside a
step, NaN accounting on the device, loss values read back only when the caller asks.  The first ``warmup`` calls run
eagerly (they are real training steps: ActNorm's data-dependent initialisation happens there), the next call captures.
``install_graphed_batch_train(trainer)`` swaps it in behind the reference trainer's own ``batch_train`` signature.
"""
import torch
from torch.nn.utils import clip_grad_norm_


class GraphedTrainStep:
    """One optimisation step as a replayed CUDA graph.

    model            : ManifoldFlow / EncoderManifoldFlow / Flow of this package, on a CUDA device
    optimizer        : a ``capturable=True`` torch optimizer (``torch.optim.AdamW(..., capturable=True)``); its ``lr`` may be
                       changed between steps by a scheduler when it was created as a tensor (``lr=torch.tensor(3e-4)``)
    loss_functions   : callables ``f(x_reco, x, log_prob, hidden=None) -> scalar`` (experiments/training/losses.py)
    loss_weights     : one weight per loss function (default 1)
    forward_kwargs   : e.g. ``{"mode": "mf-fixed-manifold"}``
    clip_gradient    : max gradient norm or None (trainer.py:180-181)
    parameters       : the parameters the norm is clipped over (default: the optimizer's)
    reducer          : optional ``ddp.GradientAllReducer`` over the trained parameters; its bucket all-reduces run after the
                       replay (multi-GPU) - the optimizer step is then captured as a second graph
    """

    def __init__(self, model, optimizer, loss_functions, loss_weights=None, forward_kwargs=None, clip_gradient=None, parameters=None,
                 warmup=3, reducer=None):
        if not all(group.get("capturable", False) for group in optimizer.param_groups):
            raise ValueError("GraphedTrainStep needs a capturable optimizer: torch.optim.AdamW(params, lr=..., capturable=True)")
        self.model, self.optimizer = model, optimizer
        self.loss_functions = list(loss_functions)
        self.loss_weights = [1.0] * len(self.loss_functions) if loss_weights is None else list(loss_weights)
        self.forward_kwargs = dict(forward_kwargs or {})
        self.clip_gradient = clip_gradient
        self.parameters = [p for g in optimizer.param_groups for p in g["params"]] if parameters is None else list(parameters)
        self.reducer = reducer
        # at least one eager step: it creates the optimizer state (allocated and zeroed inside a capture it would be reset by
        # every replay) and runs ActNorm's data-dependent initialisation
        self.warmup = max(1, int(warmup))
        self._calls = 0
        self._graphs = {}          # input signature -> (graph, optimizer graph | None, static inputs, static outputs)
        self._side = None
        dev = self.parameters[0].device
        self.nan_counts = torch.zeros(3, dtype=torch.int64, device=dev)   # NaNs seen in x_reco, log_prob, loss (trainer.py:586-598)
        self.steps_done = 0

    # ---- the step body (identical eager and captured) ---------------------------------------------------------------
    def _forward_backward(self, x, context):
        kwargs = dict(self.forward_kwargs)
        if context is not None:
            kwargs["context"] = context
        results = self.model(x, **kwargs)
        x_reco, log_prob = results[0], results[1]
        hidden = results[3] if len(results) == 4 else None
        losses = [fn(x_reco, x, log_prob, hidden=hidden) for fn in self.loss_functions]
        loss = self.loss_weights[0] * losses[0]
        for w, term in zip(self.loss_weights[1:], losses[1:]):
            loss = loss + w * term
        with torch.no_grad():   # NaN accounting without a host round trip
            self.nan_counts[0] += torch.isnan(x_reco).sum()
            if log_prob is not None:
                self.nan_counts[1] += torch.isnan(log_prob).sum()
            self.nan_counts[2] += torch.isnan(loss).sum()
        if self.reducer is not None:
            self.reducer.zero_grad()
            for p in self.model.parameters():
                if id(p) not in self.reducer._views:
                    p.grad = None
        else:
            self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        return loss.detach(), torch.stack([term.detach() for term in losses])

    def _update(self):
        if self.clip_gradient is not None:
            clip_grad_norm_(self.parameters, self.clip_gradient)
        self.optimizer.step()

    def _eager(self, x, context):
        loss, terms = self._forward_backward(x, context)
        if self.reducer is not None:
            self.reducer.finish()
        self._update()
        return loss, terms

    # ---- public -----------------------------------------------------------------------------------------------------
    def __call__(self, x, context=None):
        """Run one step on (x, context) (CUDA tensors, or pinned host tensors: they are copied into the graph's static
        inputs).  Returns (loss, loss_terms) as device tensors that are overwritten by the next call."""
        self._calls += 1
        self.steps_done += 1
        if self._side is None:
            self._side = torch.cuda.Stream()
        key = (tuple(x.shape), None if context is None else tuple(context.shape))
        if self._calls <= self.warmup or (key not in self._graphs and len(self._graphs) >= 4):
            # eager steps run on the side stream too: a backward on the default stream would tie the parameters' gradient
            # accumulators to the legacy stream, which a later capture may not wait for
            dev = self.parameters[0].device
            self._side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                out = self._eager(x.to(dev, non_blocking=True), None if context is None else context.to(dev, non_blocking=True))
            torch.cuda.current_stream().wait_stream(self._side)
            return out
        if key not in self._graphs:
            self._graphs[key] = self._capture(x, context)
        graph, opt_graph, (sx, sc), outputs = self._graphs[key]
        sx.copy_(x, non_blocking=True)
        if sc is not None:
            sc.copy_(context, non_blocking=True)
        graph.replay()
        if opt_graph is not None:   # multi-GPU: the bucket all-reduces sit between backward and the update
            self.reducer.finish()
            opt_graph.replay()
        return outputs

    def _capture(self, x, context):
        dev = self.parameters[0].device
        sx = torch.empty(x.shape, dtype=torch.float32, device=dev)
        sc = None if context is None else torch.empty(context.shape, dtype=torch.float32, device=dev)
        sx.copy_(x)
        if sc is not None:
            sc.copy_(context)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        opt_graph = None
        # this capture is itself a real step: the step it records is replayed once right away
        self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(graph, stream=self._side):
            outputs = self._forward_backward(sx, sc)
            if self.reducer is None:
                self._update()
        if self.reducer is not None:
            opt_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(opt_graph, stream=self._side):
                self._update()
        return graph, opt_graph, (sx, sc), outputs

    def loss_values(self, outputs):
        """One device -> host read of (loss, terms) as Python floats."""
        loss, terms = outputs
        flat = torch.cat([loss.reshape(1), terms.reshape(-1)]).tolist()
        return flat[0], flat[1:]

    def check_nans(self):
        """Host check of the device-side NaN counters (call once per epoch, not per batch)."""
        counts = self.nan_counts.tolist()
        self.nan_counts.zero_()
        return dict(x_reco=counts[0], log_prob=counts[1], loss=counts[2])


def install_graphed_batch_train(trainer, capturable_optimizer, warmup=3, reducer=None):
    """Replace ``trainer.batch_train`` of a reference ``ForwardTrainer`` / ``ConditionalForwardTrainer``
    (experiments/training/trainer.py:512-521) by a graphed step with the same signature and return value
    ``(loss: float, loss_contributions: list of float)`` - one device -> host read per batch instead of one per loss term
    plus three NaN checks plus the ``last_batch`` copies.  ``capturable_optimizer`` replaces the optimizer the trainer
    built (it must be constructed with ``capturable=True`` over the same parameters)."""
    steps = {}

    def batch_train(batch_data, loss_functions, loss_weights, optimizer, clip_gradient, parameters, forward_kwargs=None, custom_kwargs=None):
        key = (tuple(id(f) for f in loss_functions), tuple(loss_weights), tuple(sorted((forward_kwargs or {}).items())), clip_gradient)
        if key not in steps:
            steps[key] = GraphedTrainStep(trainer.model, capturable_optimizer, loss_functions, loss_weights, forward_kwargs, clip_gradient,
                                          list(parameters) if parameters is not None else None, warmup=warmup, reducer=reducer)
        step = steps[key]
        x = batch_data[0]
        if x.dim() < 2:
            x = x.view(x.size(0), -1)
        context = batch_data[1] if getattr(trainer, "conditional", False) or type(trainer).__name__.startswith("Conditional") else None
        if context is not None and context.dim() < 2:
            context = context.view(context.size(0), -1)
        return step.loss_values(step(x.float(), None if context is None else context.float()))

    trainer.batch_train = batch_train
    trainer._graphed_steps = steps
    return trainer
