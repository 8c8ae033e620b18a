"""Gradient exchange of the hot path's own parameters under data parallelism.

The path shards by image: the kernels own no parameters and need no
collective.  The only parameters inside the module boundary are the two ``nn.Linear`` of every
``MSDeformableAttention`` (reference src/d_fine/arch/dfine_decoder.py:79-80); in the reference
their gradients are averaged over the ranks by ``DistributedDataParallel``
(reference src/dl/train.py:161-166).  ``GradBucket`` is that step for this path alone: one flat
bucket, ONE all-reduce (NCCL on GPUs, NVLink / NVSwitch), the averaged gradients written back
into ``param.grad``.  Plain ``torch.distributed`` plumbing -- no kernels of ours.
"""
from __future__ import annotations

import os

from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def path_parameters(modules: Iterable[torch.nn.Module]) -> List[torch.nn.Parameter]:
    """Trainable parameters of the given MSDeformableAttention modules in a fixed order
    (module order, then ``named_parameters`` order) -- the same on every rank."""
    out: List[torch.nn.Parameter] = []
    for m in modules:
        out.extend(p for _, p in m.named_parameters() if p.requires_grad)
    return out


class GradBucket:
    """Flat all-reduce bucket over a fixed parameter list."""

    def __init__(self, params: Sequence[torch.nn.Parameter], group=None):
        self.params = list(params)
        if not self.params:
            raise ValueError("GradBucket: empty parameter list")
        self.group = group
        self.numels = [p.numel() for p in self.params]
        self.nbytes = sum(p.numel() * p.element_size() for p in self.params)

    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def reduce(self, grads: Optional[Sequence[torch.Tensor]] = None) -> torch.Tensor:
        """Average the gradients over the ranks of the group, in place.  ``grads`` defaults to
        the parameters' ``.grad``; every rank must pass gradients for every parameter
        (the reference wraps with ``find_unused_parameters=False``).  Returns the flat bucket.
        Enqueued on the current stream; does not synchronise."""
        if grads is None:
            grads = [p.grad for p in self.params]
        if len(grads) != len(self.params) or any(g is None for g in grads):
            missing = [i for i, g in enumerate(grads) if g is None]
            raise RuntimeError(f"GradBucket.reduce: parameters {missing} have no gradient on this rank")
        for g, n in zip(grads, self.numels):
            if g.numel() != n:
                raise RuntimeError("GradBucket.reduce: gradient shape does not match its parameter")
        flat = torch.cat([g.reshape(-1) for g in grads])
        world = self.world()
        if world > 1:
            if flat.is_cuda:
                dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)   # NCCL averages in the collective
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)   # gloo: no AVG
                flat.mul_(1.0 / world)
            with torch.no_grad():
                torch._foreach_copy_(list(grads), [c.view_as(g) for c, g in zip(flat.split(self.numels), grads)])
        return flat


class LayerwiseGradSync:
    """DDP-style overlap for the path's parameters: one bucket per MSDeformableAttention module, reduced on
    a side stream as soon as the module's last parameter gradient has been accumulated (the backward of
    decoder layer i+1 finishes before that of layer i starts, reference dfine_decoder.py:470-515), so the
    all-reduce of layer i+1 runs under the backward kernels of layers i, i-1, ...; only the first layer's
    exchange is exposed.  Works under CUDA-graph capture (the side stream forks from and re-joins the
    capturing stream) and with CPU tensors / gloo (no streams: the reduce runs in the hook).

    Usage:  sync = LayerwiseGradSync(modules); ...; loss.backward(); sync.finish()
    """

    def __init__(self, modules: Iterable[torch.nn.Module], group=None):
        self.buckets: List[GradBucket] = []
        self._pending: List[int] = []
        self._handles = []
        self._side = None
        self.group = group
        for m in modules:
            params = path_parameters([m])
            if not params:
                continue
            k = len(self.buckets)
            self.buckets.append(GradBucket(params, group))
            self._pending.append(len(params))
            for p in params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._make_hook(k)))
        if not self.buckets:
            raise ValueError("LayerwiseGradSync: no trainable parameters")
        self.nbytes = sum(b.nbytes for b in self.buckets)
        self.launched = 0

    def _make_hook(self, k: int):
        def hook(param):
            self._pending[k] -= 1
            if self._pending[k] == 0:
                self._launch(k)
        return hook

    def _launch(self, k: int) -> None:
        bucket = self.buckets[k]
        self.launched += 1
        g0 = bucket.params[0].grad
        if g0 is None or not g0.is_cuda:
            bucket.reduce()
            return
        dev = g0.device
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
        cur = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._side.wait_event(ev)          # the gradients of this bucket are complete on `cur`
        with torch.cuda.stream(self._side):
            bucket.reduce()

    def finish(self) -> None:
        """Join: the current stream waits for every exchange started by this backward pass; a bucket
        whose gradients never arrived is an error (every rank must reduce the same buckets)."""
        missing = [k for k, n in enumerate(self._pending) if n != 0]
        for k, b in enumerate(self.buckets):
            self._pending[k] = len(b.params)
        if missing:
            raise RuntimeError(f"LayerwiseGradSync.finish: buckets {missing} did not receive all their gradients")
        if self._side is not None:
            torch.cuda.current_stream(self._side.device).wait_stream(self._side)

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


class NvlsGradExchange:
    """The gradient all-reduce of the path's Linears without a collective launch (B200 / NVSwitch).

    One symmetric float32 buffer (torch.distributed._symmetric_memory: a replica on every rank plus an NVLS
    multicast mapping over all replicas) holds the [dW; db] block of every layer.  While the exchange is
    active, `ops.linear_wgrad` writes each layer's gradient into a local block and returns views of the
    REPLICA as the gradients; `end_step` launches `dfine_multicast_add`: every rank adds its blocks, scaled
    by 1 / world, into ALL replicas with multimem.red through the switch (the sum over the ranks is formed
    in the NVSwitch), followed by a device-side barrier.  Per step:

        begin_step():  zero the local replica; barrier A   (on a forked side stream: hidden under the step)
        ... backward: dfine_linear_wgrad per layer into the local blocks ...
        end_step():    dfine_multicast_add (1.2 MB, one launch; layer-wise variant: block_ready); barrier B ->
                       every replica holds the rank average (DDP's result, reference src/dl/train.py:161-166)

    All of it is capturable in a CUDA graph.  Raises if the device / fabric has no multicast support."""

    def __init__(self, slots: Sequence[int], device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("NvlsGradExchange needs an initialised process group")
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.slots = [int(n) for n in slots]                     # floats per layer block (N*K + N)
        self.offsets, off = [], 0
        for n in self.slots:
            self.offsets.append(off)
            off += (n + 63) & ~63                                 # 256-byte aligned blocks
        self.total = off
        self.buf = symm.empty(self.total, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.buf, self.group)
        self.mc_ptr = int(self.handle.multicast_ptr)
        if self.mc_ptr == 0:
            raise RuntimeError("NvlsGradExchange: no NVLS multicast mapping on this device / fabric")
        self.scale = 1.0 / self.world
        self.nbytes = self.total * 4
        self.local = torch.zeros(self.total, dtype=torch.float32, device=device)   # this rank's own gradients
        self._next = 0
        self._side = None
        # layer-wise (DFINE_NVLS_LAYERWISE=1, off by default): a layer's block is added into the replicas on the
        # side stream as soon as its weight-gradient kernel is done, under the backward of the remaining layers.
        # Measured at N = 2 inside the step (bench.py, DFINE_NVLS_PROBE): exchange off 0.951 ms; coupling only
        # (barrier A + join) 0.958; + the 1.18-MB add 0.982 (the add alone takes 3.9 us on an idle pair of GPUs);
        # + barrier B 0.988.  Layer-wise adds: 0.991 vs 0.995 in the same run -- the 24 us are not the add's own
        # duration on the critical path but what the reductions cost the kernels running beside them.
        self.layerwise = os.environ.get("DFINE_NVLS_LAYERWISE", "0") == "1"
        self._added = 0
        self.buf.zero_()
        self.handle.barrier(channel=0)

    # -- protocol -----------------------------------------------------------------------------
    def begin_step(self) -> None:
        """Zero the local replica and meet the other ranks (barrier A) on a side stream forked from the
        current one: both run under the step's forward / backward kernels; end_step() joins."""
        self._next = 0
        self._added = 0
        dev = self.local.device
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
        self._side.wait_stream(torch.cuda.current_stream(dev))    # after the last reader of the replica
        with torch.cuda.stream(self._side):
            self.buf.zero_()
            if "begin" not in os.environ.get("DFINE_NVLS_PROBE", ""):
                self.handle.barrier(channel=0)

    def end_step(self) -> None:
        from . import _lib
        if self._next != len(self.slots):
            raise RuntimeError(f"NvlsGradExchange.end_step: {self._next} of {len(self.slots)} gradient blocks "
                               "were produced (every rank must produce all of them)")
        torch.cuda.current_stream(self.local.device).wait_stream(self._side)   # every replica is zero
        skip = os.environ.get("DFINE_NVLS_PROBE", "")     # timing probes only (results are wrong with them)
        if "add" not in skip and self._added < len(self.slots):
            # whatever was not added layer by layer (everything when layerwise is off)
            o = self.offsets[self._added]
            with torch.cuda.device_of(self.local):
                rc = _lib.lib().dfine_multicast_add(self.local.data_ptr() + 4 * o, self.mc_ptr + 4 * o, self.total - o,
                                                    self.scale, torch.cuda.current_stream(self.local.device).cuda_stream)
            _lib.check(rc, "dfine_multicast_add")
            self._added = len(self.slots)
        if "barrier" not in skip:
            self.handle.barrier(channel=1)

    # -- called by ops.linear_wgrad --------------------------------------------------------------
    def next_block(self, n_floats: int):
        """(local block, replica view) of the next layer's [dW; db] gradient, in backward order."""
        k = self._next
        if k >= len(self.slots) or self.slots[k] != n_floats:
            raise RuntimeError(f"NvlsGradExchange: unexpected gradient block {k} of {n_floats} floats "
                               f"(configured: {self.slots})")
        self._next += 1
        o = self.offsets[k]
        return self.local[o:o + n_floats], self.buf[o:o + n_floats]

    def block_ready(self) -> None:
        """Called by ops.linear_wgrad right after the weight-gradient kernel of the block handed out last was
        enqueued on the current stream: (layer-wise mode) fork its multicast add onto the side stream."""
        from . import _lib
        k = self._next - 1
        if not self.layerwise or k != self._added or "add" in os.environ.get("DFINE_NVLS_PROBE", ""):
            return
        dev = self.local.device
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        o = self.offsets[k]
        n = (self.slots[k] + 3) & ~3
        with torch.cuda.stream(self._side):          # after barrier A (same stream): every replica is zero
            self._side.wait_event(ev)
            with torch.cuda.device_of(self.local):
                rc = _lib.lib().dfine_multicast_add(self.local.data_ptr() + 4 * o, self.mc_ptr + 4 * o, n, self.scale,
                                                    self._side.cuda_stream)
            _lib.check(rc, "dfine_multicast_add")
        self._added = k + 1

    def __enter__(self):
        from . import ops
        ops._WGRAD_EXCHANGE = self
        return self

    def __exit__(self, *exc):
        from . import ops
        ops._WGRAD_EXCHANGE = None
        return False
