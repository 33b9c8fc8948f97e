"""Whole-step CUDA graphs for the distillation step.

One CRD step is ~65 kernel launches (Embed heads, the fused gather kernel + 2 finishers, the row update, the heads'
backward, fused Adam); at the reference's real sizes (batch 16-64, n_data ~1e3, `options.py:136,87-91`) and even at
BASELINE config 2 with a per-step host sync (`loss.item()`, `train_test_MT.py:224-236`) the step is bound by the host
issuing those launches, not by the kernels.  `GraphedTrainStep` captures

    loss = loss_fn(*inputs);  loss.backward();  optimizer.step()

once into a `torch.cuda.CUDAGraph` over static input buffers and replays it: one launch per step.  The reference has
no equivalent (eager PyTorch 1.10); this is the B200-side answer to its five `.item()` syncs and dozens of tiny kernels.

Rules (checked where possible):
  * shapes are fixed at construction; inputs are copied into the static buffers (device or pinned-host sources);
  * drop every reference to losses / outputs of earlier eager steps before constructing (a live autograd graph pins
    the parameters' AccumulateGrad nodes to the stream those steps ran on, which breaks the capture);
  * run >= 1 eager step first (the constructor does `warmup` of them): the first CRD call sets Z with a host sync
    (`CRD_criterion.py:52-59`), the optimizer allocates its state, the Kronecker weights get packed;
  * the optimizer must be capturable (`torch.optim.Adam(..., capturable=True)`; plain SGD is as it is);
  * randomness must come from device state: torch's own dropout is graph-safe, and the Kronecker dropout of the fusion
    modules draws its per-call seed word on the device (`fusion.kron_linear`), so every replay gets a fresh mask;
  * replays change the parameters without bumping their autograd version counters; `replay` bumps them so that
    version-keyed caches (the packed Kronecker weight) stay honest if eager calls are mixed in.
"""
from __future__ import annotations

import torch

_bump = getattr(torch.autograd.graph, "increment_version", None)


_END_OF_STEP = []


def defer_to_end_of_step(fn):
    """Run `fn()` after the optimizer step of the GraphedTrainStep that is executing / being captured (modules use it to
    join side streams they forked for work that may overlap the backward pass: a capture has to see every forked stream
    rejoin before it ends).  Outside a GraphedTrainStep the callbacks simply run at the next `run_deferred()`."""
    _END_OF_STEP.append(fn)


def run_deferred():
    while _END_OF_STEP:
        _END_OF_STEP.pop(0)()


class GraphedTrainStep:
    """step = GraphedTrainStep(loss_fn, params, optimizer, example_inputs); loss = step(*inputs)

    loss_fn(*inputs) -> scalar/[1] loss, built from modules whose parameters are `params`.
    `example_inputs`: CUDA tensors fixing shapes/dtypes (floating-point ones listed in `grad_inputs` get .grad).
    `n_buffers=2` keeps two captured graphs over two static input sets so that the upload of step i+1 (on another
    stream) can overlap the replay of step i."""

    def __init__(self, loss_fn, params, optimizer, example_inputs, grad_inputs=(), warmup=3, n_buffers=1,
                 before_capture=None, pass_next_inputs=False):
        """pass_next_inputs: call `loss_fn(*inputs, next_inputs=<static inputs of the slot replayed after this one>)` --
        for losses that can start work on the next batch early (ShardedCRDLoss routes the next contrast_idx under this
        step's tail).  Only meaningful with `replay()` over pre-filled buffers in round-robin order (an even n_buffers)."""
        self.pass_next_inputs = bool(pass_next_inputs)
        self.params = [p for p in params]
        self.optimizer = optimizer
        self.loss_fn = loss_fn
        dev = example_inputs[0].device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs CUDA tensors (no CPU path)")
        for g in optimizer.param_groups:       # Adam-family optimizers keep `step` on the host unless capturable=True
            if "capturable" in g and not g["capturable"]:
                raise RuntimeError("optimizer must be constructed with capturable=True to be replayed from a CUDA graph")
        self.static_in, self.static_loss, self.graphs = [], [], []
        side = torch.cuda.Stream(dev, priority=-1)     # the step itself outranks side work its modules fork (e.g. routing of the next batch)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(n_buffers):
                ins = [t.detach().clone() for t in example_inputs]
                for i in grad_inputs:
                    ins[i].requires_grad_(True)
                self.static_in.append(ins)
            for _ in range(max(1, warmup)):      # with pass_next_inputs: the LAST slot, so that slot 0 finds itself prefetched
                self._eager(self.static_in[-1 if self.pass_next_inputs else 0])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if before_capture is not None:
            before_capture()
        pool = None
        for ins in self.static_in:
            g = torch.cuda.CUDAGraph()
            for p in self.params:
                p.grad = None
            for i in grad_inputs:
                ins[i].grad = None
            # thread_local: a NCCL watchdog thread polling events must not invalidate the capture
            # ... and capture on the stream the warm-up ran on: AccumulateGrad nodes that outlive the warm-up keep its
            # stream, and autograd may not make another (let alone the legacy) stream wait on a capturing one
            with torch.cuda.graph(g, pool=pool, stream=side, capture_error_mode="thread_local"):
                loss = self._loss(ins)
                loss.backward()
                self.optimizer.step()
                run_deferred()
            pool = g.pool()
            self.graphs.append(g)
            self.static_loss.append(loss.detach())
        self._next = 0

    def _eager(self, ins):
        for p in self.params:
            p.grad = None
        loss = self._loss(ins)
        loss.backward()
        self.optimizer.step()
        run_deferred()
        return loss

    def _loss(self, ins):
        if not self.pass_next_inputs:
            return self.loss_fn(*ins)
        k = next(i for i, s in enumerate(self.static_in) if s is ins)
        return self.loss_fn(*ins, next_inputs=self.static_in[(k + 1) % len(self.static_in)])

    def buffers(self, slot=None):
        """Static input tensors of a slot (default: the one the next call replays) -- copy into them yourself to skip
        the per-call copy, e.g. `buf.copy_(pinned_host, non_blocking=True)` on an upload stream."""
        return self.static_in[self._next if slot is None else slot]

    def replay(self, slot=None):
        """Replay one captured step on the current stream over whatever the slot's static buffers hold."""
        s = self._next if slot is None else slot
        self.graphs[s].replay()
        if _bump is not None:
            for p in self.params:
                _bump(p)
        self._next = (s + 1) % len(self.graphs)
        return self.static_loss[s]

    def __call__(self, *inputs):
        ins = self.static_in[self._next]
        for dst, src in zip(ins, inputs):
            dst.detach().copy_(src, non_blocking=True)
        return self.replay()
