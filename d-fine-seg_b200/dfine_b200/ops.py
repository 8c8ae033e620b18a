"""torch.autograd front ends of the C-ABI kernels (host plumbing only: shapes, dtypes,
pointers, streams).  All compute happens in libdfine_b200.so; CPU tensors are rejected.
"""
from __future__ import annotations

import os

import functools
import weakref
from typing import List, Optional, Sequence, Tuple, Union

import torch

from . import _lib
from ._lib import BF16, F32, MSDA_FUSED_INPUTS, check

_DT = {torch.float32: F32, torch.bfloat16: BF16}

# Instrumentation used by bench.py: number of kernels launched through the C-ABI, and
# optional CUDA-event brackets around each launch (on the launching stream).
LAUNCHES = {"count": 0}
_WGRAD_EXCHANGE = None     # grad_sync.NvlsGradExchange while a data-parallel step is running
_TIMERS = None  # None or dict: kernel name -> list of (start_event, end_event)


def enable_kernel_timers(enable: bool = True) -> None:
    global _TIMERS
    _TIMERS = {} if enable else None


def kernel_timers():
    return _TIMERS


class _timed:
    """Counts the launch and, when timers are enabled, brackets it with CUDA events."""

    def __init__(self, name: str, t: torch.Tensor, kernels: int = 1):
        self.name, self.t, self.kernels = name, t, kernels

    def __enter__(self):
        LAUNCHES["count"] += self.kernels
        if _TIMERS is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record(torch.cuda.current_stream(self.t.device))
        return self

    def __exit__(self, *exc):
        if _TIMERS is not None:
            self.e.record(torch.cuda.current_stream(self.t.device))
            _TIMERS.setdefault(self.name, []).append((self.s, self.e))
        return False


def _dt(t: torch.Tensor, what: str) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"{what}: dtype {t.dtype} not supported (float32 or bfloat16)") from None


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "dfine_b200 kernels run on CUDA tensors only (there is no CPU fallback); "
                f"got a tensor on {t.device}")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class LevelSpec:
    """Per-level tables in the form the C-ABI wants (host int32 arrays)."""

    def __init__(self, spatial_shapes: Sequence[Sequence[int]], num_points: Sequence[int]):
        shapes = [(int(h), int(w)) for h, w in spatial_shapes]
        npts = [int(n) for n in num_points]
        assert len(npts) == len(shapes), "len(num_points) must equal the number of levels"
        self.shapes, self.npts = shapes, npts
        self.n_lvl = len(shapes)
        self.sizes = [h * w for h, w in shapes]
        self.starts = [sum(self.sizes[:i]) for i in range(self.n_lvl)]
        self.L = sum(self.sizes)
        self.P = sum(npts)
        self.hw_c = _lib.i32_array([v for hw in shapes for v in hw])
        self.start_c = _lib.i32_array(self.starts)
        self.npts_c = _lib.i32_array(npts)


@functools.lru_cache(maxsize=64)
def _level_spec(shapes: Tuple[Tuple[int, int], ...], npts: Tuple[int, ...]) -> LevelSpec:
    return LevelSpec(shapes, npts)


def level_spec(spatial_shapes, num_points) -> LevelSpec:
    return _level_spec(tuple((int(h), int(w)) for h, w in spatial_shapes),
                       tuple(int(n) for n in num_points))


# --------------------------------------------------------------------------------------
# value layout: recover `memory [B, L, H*c]` behind TransformerDecoder.value_op's views
# --------------------------------------------------------------------------------------
def memory_from_value(value: Union[torch.Tensor, Sequence[torch.Tensor]], spec: LevelSpec
                      ) -> Tuple[torch.Tensor, int, int, bool]:
    """Returns (memory [B, L, H*c], H, c, zero_copy).

    `value` is either the `memory` tensor itself ([B, L, C] with `H` unknown is not
    accepted here -- pass [B, L, H, c]) or the tuple of per-level views
    [B, H, c, h_l*w_l] that TransformerDecoder.value_op builds
    (reference dfine_decoder.py:416-426: reshape -> permute(0,2,3,1) -> split).  The views
    alias one [B, L, H, c] buffer; when that pattern is recognised the kernel reads the
    buffer in place and autograd is routed to the base tensor (one [B, L, C] gradient, no
    cat / permute / reshape copies).  Anything else is packed with one copy.
    """
    if isinstance(value, torch.Tensor):
        if value.dim() != 4:
            raise ValueError("value tensor must be [B, L, H, c]")
        B, L, H, c = value.shape
        if L != spec.L:
            raise ValueError(f"value length {L} != sum(h*w) = {spec.L}")
        if value.dtype not in _DT:
            value = _supported_memory(value)
        return value.reshape(B, L, H * c), H, c, True
    views = list(value)
    if len(views) != spec.n_lvl:
        raise ValueError(f"expected {spec.n_lvl} value levels, got {len(views)}")
    B, H, c, _ = views[0].shape
    C = H * c
    for v, n in zip(views, spec.sizes):
        if tuple(v.shape) != (B, H, c, n):
            raise ValueError(f"value level shape {tuple(v.shape)} != {(B, H, c, n)}")
    base = views[0]._base
    want_strides = (spec.L * C, c, 1, C)
    ok = (base is not None and base.dim() == 3 and tuple(base.shape) == (B, spec.L, C)
          and base.is_contiguous() and base.dtype == views[0].dtype
          and base.requires_grad == views[0].requires_grad)
    if ok:
        esz = base.element_size()
        for v, st in zip(views, spec.starts):
            vb = v._base
            if (vb is None or vb.data_ptr() != base.data_ptr() or vb.shape != base.shape
                    or tuple(v.stride()) != want_strides
                    or v.data_ptr() != base.data_ptr() + st * C * esz):
                ok = False
                break
    if ok:
        return _supported_memory(base), H, c, True
    # generic (slow) path: one packing copy, autograd flows through the views
    mem = torch.cat([v.permute(0, 3, 1, 2) for v in views], dim=1)  # [B, L, H, c]
    mem = mem.reshape(B, spec.L, C)
    return (mem if mem.dtype in _DT else mem.float()), H, c, False


# float16 route (the reference trainer's default AMP dtype: `autocast(device)` without a dtype is
# float16 on CUDA, src/dl/train.py:545-551).  The kernels read float32 / bfloat16 values; a float16
# `memory` is widened to float32 ONCE per forward -- exactly the cast grid_sample's autocast-to-fp32
# rule makes the reference do per layer and level -- and every decoder layer of that forward shares
# the widened tensor (and therefore one gradient hub).  Single entry: it is replaced by the next
# forward and dropped when the hub has delivered the gradient.
_WIDENED = {}


def _supported_memory(base: torch.Tensor) -> torch.Tensor:
    if base.dtype in _DT:
        return base
    key = (id(base), base._version, torch.is_grad_enabled(), base.data_ptr())
    ent = _WIDENED.get("entry")
    if ent is not None and ent[0] == key and ent[1]() is base:
        return ent[2]
    wide = base.float()
    _WIDENED["entry"] = (key, weakref.ref(base), wide)
    return wide


# --------------------------------------------------------------------------------------
# K1 / K2
# --------------------------------------------------------------------------------------
def _mem_strides(memory: torch.Tensor) -> Tuple[int, int]:
    if memory.stride(2) != 1:
        raise ValueError("memory must have unit stride along the channel dimension")
    return memory.stride(0), memory.stride(1)


def msda_forward_raw(memory, spec: LevelSpec, H: int, samp, attn, ref, pts_scale,
                     offset_scale: float, fused: bool, out_dtype: torch.dtype,
                     want_idx: bool = False, samp_rs: int = 0, attn_rs: int = 0,
                     records: Optional[torch.Tensor] = None):
    """Direct call of dfine_msda_fwd (no autograd).  Returns out [B, Lq, C] (, idx).
    samp_rs / attn_rs: row strides (elements) when samp / attn alias a wider tensor."""
    _require_cuda(memory, samp, attn, ref, pts_scale)
    B, L, C = memory.shape
    c = C // H
    Lq = samp.shape[1]
    sb, sl = _mem_strides(memory)
    out = torch.empty((B, Lq, C), dtype=out_dtype, device=memory.device)
    idx = (torch.empty((B, Lq, H, spec.P, 4), dtype=torch.int32, device=memory.device)
           if want_idx else None)
    with torch.cuda.device_of(memory), _timed("msda_fwd", memory):
        rc = _lib.lib().dfine_msda_fwd(
            memory.data_ptr(), sb, sl, spec.hw_c, spec.start_c, spec.npts_c, spec.n_lvl,
            samp.data_ptr(), attn.data_ptr(), _ptr(ref), _ptr(pts_scale), float(offset_scale),
            out.data_ptr(), _ptr(idx), B, Lq, H, c, _dt(memory, "value"), _dt(samp, "samp"),
            _dt(out, "out"), (MSDA_FUSED_INPUTS if fused else 0),
            samp_rs, attn_rs, _ptr(records),
            _stream(memory))
    check(rc, "dfine_msda_fwd")
    return (out, idx) if want_idx else out


def msda_backward_raw(memory, spec: LevelSpec, H: int, samp, attn, ref, pts_scale,
                      offset_scale: float, fused: bool, grad_out,
                      gv_dtype: torch.dtype = torch.float32, force_atomic: bool = False,
                      samp_rs: int = 0, attn_rs: int = 0, grad_raw: Optional[torch.Tensor] = None,
                      records: Optional[torch.Tensor] = None,
                      accumulate_into: Optional[torch.Tensor] = None):
    """Direct call of dfine_msda_bwd.  Returns (grad_memory [B,L,C] in `gv_dtype`, fp32
    grad_samp, fp32 grad_attn).  The library normally produces grad_memory with its
    atomic-free gather path directly in `gv_dtype`; shapes it cannot take (or
    force_atomic=True) use fp32 vector reductions, followed by a cast if bf16 was asked.
    accumulate_into: a running [B,L,C] gradient of `memory` (float32 or bfloat16) that this
    layer's grad_value is ADDED to in place (DFINE_MSDA_GRAD_VALUE_ACCUMULATE); it is returned
    as grad_memory."""
    _require_cuda(memory, samp, attn, grad_out)
    B, L, C = memory.shape
    c = C // H
    Lq = samp.shape[1]
    sb, sl = _mem_strides(memory)
    dev = memory.device
    base_flags = MSDA_FUSED_INPUTS if fused else 0
    if grad_raw is not None:
        # both gradients go into one [B, Lq, 3HP] buffer (concatenated Linear), in its dtype
        g_samp, g_attn = grad_raw, grad_raw.reshape(-1)[2 * H * spec.P:]
        gs_rs = ga_rs = grad_raw.shape[-1]
        if grad_raw.dtype == torch.bfloat16:
            base_flags |= _lib.MSDA_GRAD_SAMP_BF16
    else:
        g_samp = torch.empty(samp.shape, dtype=torch.float32, device=dev)
        g_attn = torch.empty(attn.shape, dtype=torch.float32, device=dev)
        gs_rs = ga_rs = 0
    ws, ws_bytes = None, 0
    if not force_atomic:
        if records is not None:   # written by the forward: the backward skips its phase 1
            ws, ws_bytes = records, records.numel()
            base_flags |= _lib.MSDA_RECORDS_VALID
        else:
            ws_bytes = _lib.lib().dfine_msda_bwd_workspace_bytes(B, Lq, H, spec.P)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

    def call(buf, flags):
        with torch.cuda.device_of(memory), _timed("msda_bwd", memory, 2):
            return _lib.lib().dfine_msda_bwd(
                memory.data_ptr(), sb, sl, spec.hw_c, spec.start_c, spec.npts_c, spec.n_lvl,
                samp.data_ptr(), attn.data_ptr(), _ptr(ref), _ptr(pts_scale), float(offset_scale),
                grad_out.data_ptr(), buf.data_ptr(), g_samp.data_ptr(), g_attn.data_ptr(),
                B, Lq, H, c, _dt(memory, "value"), _dt(samp, "samp"), _dt(grad_out, "grad_out"),
                flags, samp_rs, attn_rs, gs_rs, ga_rs, _ptr(ws), ws_bytes, _stream(memory))

    if accumulate_into is not None:
        g_mem = accumulate_into
        if tuple(g_mem.shape) != (B, spec.L, C) or not g_mem.is_contiguous() or g_mem.dtype not in _DT:
            raise ValueError("accumulate_into must be a contiguous [B, L, C] float32 / bfloat16 tensor")
        flags = base_flags | _lib.MSDA_GRAD_VALUE_ACCUMULATE
        if g_mem.dtype == torch.bfloat16:
            flags |= _lib.MSDA_GRAD_VALUE_BF16
        if force_atomic:
            flags |= _lib.MSDA_FORCE_ATOMIC
        rc = call(g_mem, flags)
        if rc == _lib.E_UNSUPPORTED and g_mem.dtype == torch.bfloat16:
            # shapes the gather path cannot take (sample lists larger than shared memory): this
            # layer's gradient through the fp32 vector-reduction path, then one in-place add
            part = torch.empty((B, spec.L, C), dtype=torch.float32, device=dev)
            check(call(part, base_flags | _lib.MSDA_FORCE_ATOMIC), "dfine_msda_bwd")
            g_mem.add_(part)
            return g_mem, g_samp, g_attn
        check(rc, "dfine_msda_bwd")
        return g_mem, g_samp, g_attn
    if gv_dtype == torch.bfloat16 and not force_atomic:
        g_mem = torch.empty((B, spec.L, C), dtype=torch.bfloat16, device=dev)
        rc = call(g_mem, base_flags | _lib.MSDA_GRAD_VALUE_BF16)
        if rc == 0:
            return g_mem, g_samp, g_attn
        if rc != _lib.E_UNSUPPORTED:
            check(rc, "dfine_msda_bwd")
    g_mem = torch.empty((B, spec.L, C), dtype=torch.float32, device=dev)
    check(call(g_mem, base_flags | (_lib.MSDA_FORCE_ATOMIC if force_atomic else 0)), "dfine_msda_bwd")
    if gv_dtype == torch.bfloat16:
        g_mem = cast_f32_to_bf16(g_mem)
    return g_mem, g_samp, g_attn


_SIDE_STREAMS = {}
_OVERLAP_GRAD_VALUE = False


def overlap_grad_value(enabled: bool = True) -> bool:
    """Switch the two-stream backward on/off (default OFF); returns the old value.  With the
    shared `memory` gradient (share_memory_grad) the grad_value kernel of a layer has no consumer
    until the hub node runs, so it can run on a side stream and overlap the dots kernel, the
    Linear backward GEMMs and the next layers' backward on the main stream.  Measured at config 3:
    1.087 ms / step (1.145 ms with a high-priority side stream) against 1.051 ms serial -- the
    two gather kernels compete for the same LSU / L2 resources -- hence off by default."""
    global _OVERLAP_GRAD_VALUE
    old, _OVERLAP_GRAD_VALUE = _OVERLAP_GRAD_VALUE, bool(enabled)
    return old


def _side_stream(dev: torch.device) -> torch.cuda.Stream:
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev, priority=0)
    return st


def msda_backward_split(memory, spec: LevelSpec, H: int, samp, attn, ref, pts_scale, offset_scale: float,
                        fused: bool, grad_out, sess: dict, records: torch.Tensor, samp_rs: int = 0,
                        attn_rs: int = 0, grad_raw: Optional[torch.Tensor] = None):
    """The backward of one layer as two independent launches of dfine_msda_bwd:
    DFINE_MSDA_BWD_VALUE_ONLY (grad_value into the hub session's shared buffer) on the side
    stream, DFINE_MSDA_BWD_DOTS_ONLY (grad_samp / grad_attn) on the current stream.  Returns
    (grad_samp, grad_attn), or None when the shape needs the fallback path (nothing launched)."""
    B, L, C = memory.shape
    c = C // H
    Lq = samp.shape[1]
    sb, sl = _mem_strides(memory)
    dev = memory.device
    base = (MSDA_FUSED_INPUTS if fused else 0) | _lib.MSDA_RECORDS_VALID
    if grad_raw is not None:
        g_samp, g_attn = grad_raw, grad_raw.reshape(-1)[2 * H * spec.P:]
        gs_rs = ga_rs = grad_raw.shape[-1]
        if grad_raw.dtype == torch.bfloat16:
            base |= _lib.MSDA_GRAD_SAMP_BF16
    else:
        g_samp = torch.empty(samp.shape, dtype=torch.float32, device=dev)
        g_attn = torch.empty(attn.shape, dtype=torch.float32, device=dev)
        gs_rs = ga_rs = 0

    def call(buf, flags, what):
        with torch.cuda.device_of(memory), _timed(what, memory):
            return _lib.lib().dfine_msda_bwd(
                memory.data_ptr(), sb, sl, spec.hw_c, spec.start_c, spec.npts_c, spec.n_lvl,
                samp.data_ptr(), attn.data_ptr(), _ptr(ref), _ptr(pts_scale), float(offset_scale),
                grad_out.data_ptr(), _ptr(buf), g_samp.data_ptr(), g_attn.data_ptr(),
                B, Lq, H, c, _dt(memory, "value"), _dt(samp, "samp"), _dt(grad_out, "grad_out"),
                flags, samp_rs, attn_rs, gs_rs, ga_rs, records.data_ptr(), records.numel(), _stream(memory))

    first = sess["buf"] is None
    buf = torch.empty((B, spec.L, C), dtype=memory.dtype, device=dev) if first else sess["buf"]
    vflags = base | _lib.MSDA_BWD_VALUE_ONLY
    if buf.dtype == torch.bfloat16:
        vflags |= _lib.MSDA_GRAD_VALUE_BF16
    if not first:
        vflags |= _lib.MSDA_GRAD_VALUE_ACCUMULATE
    main, side = torch.cuda.current_stream(dev), _side_stream(dev)
    ready = torch.cuda.Event()
    ready.record(main)                       # grad_out (and the buffer's previous life) are complete
    with torch.cuda.stream(side):
        side.wait_event(ready)
        rc = call(buf, vflags, "msda_bwd_value")
        if rc == _lib.E_UNSUPPORTED:
            return None
        check(rc, "dfine_msda_bwd(value)")
        done = torch.cuda.Event()
        done.record(side)
    # what the side-stream kernel reads stays alive until the hub has joined the two streams (the
    # caching allocator -- and a CUDA-graph capture's private pool -- would otherwise hand the
    # blocks to later main-stream work while the kernel is still running)
    sess.setdefault("keep", []).extend((grad_out, records))
    sess["buf"], sess["event"] = buf, done
    check(call(None, base | _lib.MSDA_BWD_DOTS_ONLY, "msda_bwd_dots"), "dfine_msda_bwd(dots)")
    return g_samp, g_attn


def new_records(memory: torch.Tensor, spec: LevelSpec, H: int, Lq: int) -> torch.Tensor:
    """Workspace for the per-sample geometry records (16 bytes per sampling point).  Allocated by a forward
    that will be differentiated: shapes the BACKWARD kernels do not take are refused here, before any
    work is done (the forward alone accepts up to DFINE_MAX_POINTS = 32 points per head)."""
    if spec.P > 16:
        raise ValueError(f"dfine_msda_bwd takes at most 16 sampling points per head (got {spec.P}); the forward "
                         "alone (no gradients) accepts up to 32")
    n = _lib.lib().dfine_msda_bwd_workspace_bytes(memory.shape[0], Lq, H, spec.P)
    return torch.empty(n, dtype=torch.uint8, device=memory.device)


def cast_f32_to_bf16(src: torch.Tensor) -> torch.Tensor:
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    with torch.cuda.device_of(src), _timed("cast_bf16", src):
        rc = _lib.lib().dfine_cast_f32_to_bf16(src.data_ptr(), dst.data_ptr(), src.numel(),
                                               _stream(src))
    check(rc, "dfine_cast_f32_to_bf16")
    return dst


# --------------------------------------------------------------------------------------
# one gradient buffer for `memory` across the decoder layers
# --------------------------------------------------------------------------------------
# All decoder layers sample the SAME `memory` (TransformerDecoder.value_op is called once,
# reference dfine_decoder.py:454, and its views feed every layer, :483).  Autograd would
# receive one [B, L, C] gradient per layer and add them pairwise (N_l - 1 full-size add
# kernels).  Instead the layers' backward kernels accumulate into ONE buffer
# (DFINE_MSDA_GRAD_VALUE_ACCUMULATE) and a hub node, which autograd runs after every layer that
# took part in the graph, hands that buffer to `memory` once.
_SHARE_MEMORY_GRAD = True


def share_memory_grad(enabled: bool = True) -> bool:
    """Switch the shared `memory` gradient buffer on/off (default on); returns the old value."""
    global _SHARE_MEMORY_GRAD
    old, _SHARE_MEMORY_GRAD = _SHARE_MEMORY_GRAD, bool(enabled)
    return old


# id(anchor tensor) -> (weakref(anchor), token, session).  Kept OFF the tensor objects: a token
# stored on a leaf `memory` would close a reference cycle through its AccumulateGrad node that
# Python's GC cannot see.  Entries leave when their hub runs backward, when the anchor dies,
# or when more than _MAX_HUBS forward-only graphs are pending.
_HUBS = {}
_MAX_HUBS = 2


class _MemoryHubFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, memory, sess, key):
        ctx.sess, ctx.key = sess, key
        ctx.set_materialize_grads(False)
        return memory.new_empty(0, dtype=torch.float32)   # token: carries the graph edge only

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, _g_token):
        ent = _HUBS.get(ctx.key)
        if ent is not None and ent[2] is ctx.sess:
            del _HUBS[ctx.key]
        tid = torch._C._current_graph_task_id() if hasattr(torch._C, "_current_graph_task_id") else None
        if ctx.sess.get("task", tid) != tid:      # what is in the buffer belongs to an earlier, unfinished pass
            ctx.sess["buf"] = None
        ctx.sess.pop("task", None)
        buf, ctx.sess["buf"] = ctx.sess.get("buf"), None
        ev = ctx.sess.pop("event", None)
        if ev is not None and buf is not None:   # grad_value kernels ran on the side stream
            torch.cuda.current_stream(buf.device).wait_event(ev)
        ctx.sess.pop("keep", None)
        _WIDENED.pop("entry", None)
        return buf, None, None


def _memory_token(memory: torch.Tensor, anchor: Optional[torch.Tensor] = None):
    """(token, session) of the gradient hub of `memory`, shared by every layer that samples
    the same tensor object `anchor` (default: `memory` itself) in one forward pass;
    (None, None) when sharing is off or memory needs no gradient."""
    if not (_SHARE_MEMORY_GRAD and memory.requires_grad and torch.is_grad_enabled()):
        return None, None
    anchor = memory if anchor is None else anchor
    key = id(anchor)
    ent = _HUBS.get(key)
    if ent is not None and ent[0]() is anchor and ent[1].requires_grad:
        return ent[1], ent[2]
    for k in [k for k, e in _HUBS.items() if e[0]() is None]:
        del _HUBS[k]
    while len(_HUBS) >= _MAX_HUBS:
        del _HUBS[next(iter(_HUBS))]
    sess = {"buf": None}
    token = _MemoryHubFn.apply(memory, sess, key)
    _HUBS[key] = (weakref.ref(anchor), token, sess)
    return token, sess


def _new_backward_pass(sess: dict) -> None:
    """The shared buffer belongs to ONE backward pass.  A pass that ran layer backwards without reaching the hub
    (torch.autograd.grad(..., inputs=[query], retain_graph=True), an exception mid-backward) leaves a stale
    sum behind: the first layer of the next pass -- recognised by the autograd engine's graph-task id -- starts a
    fresh buffer instead of accumulating onto it."""
    tid = torch._C._current_graph_task_id() if hasattr(torch._C, "_current_graph_task_id") else None
    if sess.get("task") != tid or tid is None or tid < 0:
        if sess.get("task", tid) != tid:
            ev = sess.pop("event", None)
            if ev is not None and sess.get("buf") is not None:
                torch.cuda.current_stream(sess["buf"].device).wait_event(ev)
            sess["buf"] = None
            sess.pop("keep", None)
        sess["task"] = tid


def _grad_memory(ctx, memory, *args, **kw):
    """Runs dfine_msda_bwd for one layer.  With a hub session the layer's grad_value lands in
    the shared buffer (first layer writes it, later ones accumulate) and None is returned for
    the per-layer gradient; otherwise the layer's own gradient is returned."""
    sess = ctx.sess
    if sess is None:
        return msda_backward_raw(memory, *args, gv_dtype=memory.dtype, **kw)
    _new_backward_pass(sess)
    if _OVERLAP_GRAD_VALUE and kw.get("records") is not None:
        got = msda_backward_split(memory, *args, sess=sess, **kw)
        if got is not None:
            return None, got[0], got[1]
        ev = sess.pop("event", None)
        if ev is not None:   # the fallback below touches the shared buffer on this stream
            torch.cuda.current_stream(memory.device).wait_event(ev)
    if sess["buf"] is None:
        g_mem, g_samp, g_attn = msda_backward_raw(memory, *args, gv_dtype=memory.dtype, **kw)
        sess["buf"] = g_mem
    else:
        _, g_samp, g_attn = msda_backward_raw(memory, *args, accumulate_into=sess["buf"], **kw)
    return None, g_samp, g_attn


class _MsdaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, memory, samp, attn, ref, pts_scale, spec, H, offset_scale, fused, out_dtype,
                token=None, sess=None):
        # token / sess: hub of the shared `memory` gradient (then `memory` arrives detached)
        samp = samp.contiguous()
        attn = attn.contiguous()
        need = any(ctx.needs_input_grad[:3]) or ctx.needs_input_grad[10]
        rec = new_records(memory, spec, H, samp.shape[1]) if need else None
        out = msda_forward_raw(memory, spec, H, samp, attn, ref, pts_scale, offset_scale, fused,
                               out_dtype, records=rec)
        ctx.rec = rec
        ctx.sess = sess if ctx.needs_input_grad[10] else None
        ctx.save_for_backward(memory, samp, attn, ref, pts_scale)
        ctx.spec, ctx.H, ctx.offset_scale, ctx.fused = spec, H, offset_scale, fused
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        memory, samp, attn, ref, pts_scale = ctx.saved_tensors
        if grad_out.dtype not in _DT:
            grad_out = grad_out.float()
        g_mem, g_samp, g_attn = _grad_memory(ctx, memory, ctx.spec, ctx.H, samp, attn, ref,
                                             pts_scale, ctx.offset_scale, ctx.fused,
                                             grad_out.contiguous(), records=ctx.rec)
        if g_samp.dtype != samp.dtype:
            g_samp = g_samp.to(samp.dtype)
            g_attn = g_attn.to(attn.dtype)
        return g_mem, g_samp, g_attn, None, None, None, None, None, None, None, None, None


class _MsdaPackedFn(torch.autograd.Function):
    """Fused-input kernels fed by ONE concatenated Linear output raw [B, Lq, 3HP]
    ([..., :2HP] sampling offsets, [..., 2HP:] attention logits): no slicing copies, one
    gradient tensor in raw's dtype written directly by the backward kernel."""

    @staticmethod
    def forward(ctx, memory, raw, ref, pts_scale, spec, H, offset_scale, out_dtype, token=None,
                sess=None):
        rs = raw.shape[-1]
        attn_view = raw.reshape(-1)[2 * H * spec.P:]
        need = any(ctx.needs_input_grad[:2]) or ctx.needs_input_grad[8]
        rec = new_records(memory, spec, H, raw.shape[1]) if need else None
        out = msda_forward_raw(memory, spec, H, raw, attn_view, ref, pts_scale, offset_scale, True,
                               out_dtype, samp_rs=rs, attn_rs=rs, records=rec)
        ctx.rec = rec
        ctx.sess = sess if ctx.needs_input_grad[8] else None
        ctx.save_for_backward(memory, raw, ref, pts_scale)
        ctx.spec, ctx.H, ctx.offset_scale = spec, H, offset_scale
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        memory, raw, ref, pts_scale = ctx.saved_tensors
        if grad_out.dtype not in _DT:
            grad_out = grad_out.float()
        rs = raw.shape[-1]
        g_raw = torch.empty_like(raw)
        attn_view = raw.reshape(-1)[2 * ctx.H * ctx.spec.P:]
        g_mem, _, _ = _grad_memory(ctx, memory, ctx.spec, ctx.H, raw, attn_view, ref, pts_scale,
                                   ctx.offset_scale, True, grad_out.contiguous(), samp_rs=rs,
                                   attn_rs=rs, grad_raw=g_raw, records=ctx.rec)
        return g_mem, g_raw, None, None, None, None, None, None, None, None


_ONES = {}


def _ones_row(n: int, dtype, device):
    key = (n, dtype, device)
    t = _ONES.get(key)
    if t is None:
        t = _ONES[key] = torch.ones(1, n, dtype=dtype, device=device)
    return t


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b as library GEMMs only (cuBLAS): under autocast the operands are cast to
    the autocast dtype like F.linear would; the backward computes grad_bias as a
    ones-row GEMV instead of a column reduction kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        cdt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
        x2 = x.reshape(-1, x.shape[-1]).to(cdt)
        w = weight.to(cdt)
        y = torch.nn.functional.linear(x2, w, bias.to(cdt))
        ctx.save_for_backward(x2, w)
        ctx.x_shape, ctx.x_dtype, ctx.w_dtype, ctx.b_dtype = x.shape, x.dtype, weight.dtype, bias.dtype
        return y.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x2, w = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        if g2.dtype != w.dtype:
            g2 = g2.to(w.dtype)
        gx = _mm(g2, w, ctx.x_dtype).reshape(ctx.x_shape)
        gw = _mm(g2.t(), x2, ctx.w_dtype)
        if colsum_supported(g2):
            gb = colsum(g2)
            if gb.dtype != ctx.b_dtype:
                gb = gb.to(ctx.b_dtype)
        else:   # odd widths: library GEMM with a row of ones
            gb = _mm(_ones_row(g2.shape[0], g2.dtype, g2.device), g2, ctx.b_dtype).reshape(-1)
        return gx, gw, gb


def linear_fwd_supported(x: torch.Tensor, N: int, x_add: Optional[torch.Tensor] = None) -> bool:
    """Shapes / dtypes dfine_linear_fwd takes (include/dfine_b200.h): x float32 (optionally + float32 / bfloat16 x_add) or
    bfloat16, rows of K = x.shape[-1] elements with K % 64 == 0, N splitting into equal tiles of <= 512 columns."""
    if os.environ.get("DFINE_LINEAR_FWD", "1") == "0":     # A/B switch: cast kernel + cuBLAS GEMM
        return False
    K = x.shape[-1]
    tiles = (N + 511) // 512
    if N % tiles or (N // tiles) % 32 or K % 64 or x.numel() == 0:
        return False
    if not (x.is_cuda and x.dtype in _DT and x.is_contiguous() and x.data_ptr() % 16 == 0):
        return False
    if x_add is not None and not (x.dtype == torch.float32 and x_add.dtype in _DT and x_add.is_cuda
                                  and x_add.shape == x.shape and x_add.is_contiguous()
                                  and x_add.data_ptr() % 16 == 0):
        return False
    return True


def linear_fwd(x: torch.Tensor, w_bf16: torch.Tensor, bias: torch.Tensor, x_add: Optional[torch.Tensor] = None,
               relu: bool = False, out_dtype: torch.dtype = torch.bfloat16, save_input: bool = False):
    """act(bf16(x [+ x_add]) @ w^T + bias) in one tcgen05 launch (dfine_linear_fwd): the arithmetic of
    F.linear under torch.autocast(bfloat16).  x [..., K] float32 / bfloat16, w_bf16 [N, K] bfloat16, bias [N]
    float32 / bfloat16.  Returns y [..., N] (out_dtype), and with save_input also the bfloat16 operand rows
    [M, K] (what dfine_linear_wgrad needs in the backward)."""
    _require_cuda(x, w_bf16, bias)
    N, K = w_bf16.shape
    if w_bf16.dtype != torch.bfloat16 or not w_bf16.is_contiguous() or x.shape[-1] != K or bias.numel() != N:
        raise ValueError("linear_fwd: need a contiguous bfloat16 weight [N, K], x [..., K] and bias [N]")
    if not linear_fwd_supported(x, N, x_add):
        raise ValueError(f"linear_fwd: unsupported input (x {tuple(x.shape)} {x.dtype}, N = {N}); "
                         "see linear_fwd_supported")
    M = x.numel() // K
    y = torch.empty((*x.shape[:-1], N), dtype=out_dtype, device=x.device)
    xs = torch.empty((M, K), dtype=torch.bfloat16, device=x.device) if save_input and x.dtype == torch.float32 else None
    with torch.cuda.device_of(x), _timed("linear_fwd", x):
        rc = _lib.lib().dfine_linear_fwd(x.data_ptr(), _DT[x.dtype], K, _ptr(x_add),
                                         _DT[x_add.dtype] if x_add is not None else F32, K, w_bf16.data_ptr(),
                                         bias.data_ptr(), _dt(bias, "bias"), y.data_ptr(), _DT[out_dtype], N,
                                         _ptr(xs), M, N, K, int(bool(relu)), _stream(x))
    check(rc, "dfine_linear_fwd")
    if save_input:
        return y, (xs if xs is not None else x.reshape(M, K))
    return y


def gate_fwd(x1: torch.Tensor, x2: torch.Tensor, w_bf16: torch.Tensor, bias: torch.Tensor, ln_weight: torch.Tensor,
             ln_bias: torch.Tensor, eps: float) -> torch.Tensor:
    """Gate.forward of the reference (dfine_decoder.py:258-271) under bf16 autocast as ONE launch
    (dfine_gate_fwd): LayerNorm(g1 * x1 + g2 * x2) with [g1 | g2] = sigmoid(cat(x1, x2) @ w^T + bias).
    x1, x2 float32 [..., C]; w_bf16 [2C, 2C] bfloat16; bias [2C]; LayerNorm parameters float32 [C]."""
    _require_cuda(x1, x2, w_bf16, bias, ln_weight, ln_bias)
    C = x1.shape[-1]
    if (x1.dtype != torch.float32 or x2.dtype != torch.float32 or x1.shape != x2.shape or not x1.is_contiguous()
            or not x2.is_contiguous() or w_bf16.dtype != torch.bfloat16 or tuple(w_bf16.shape) != (2 * C, 2 * C)
            or not w_bf16.is_contiguous() or ln_weight.dtype != torch.float32 or ln_bias.dtype != torch.float32):
        raise ValueError("gate_fwd: need contiguous float32 x1, x2 [..., C], bfloat16 w [2C, 2C], float32 LayerNorm "
                         "parameters")
    M = x1.numel() // C
    out = torch.empty_like(x1)
    with torch.cuda.device_of(x1), _timed("gate_fwd", x1):
        rc = _lib.lib().dfine_gate_fwd(x1.data_ptr(), C, x2.data_ptr(), C, w_bf16.data_ptr(), bias.data_ptr(),
                                       _dt(bias, "bias"), ln_weight.data_ptr(), ln_bias.data_ptr(), float(eps),
                                       out.data_ptr(), C, M, C, _stream(x1))
    check(rc, "dfine_gate_fwd")
    return out


def ffn_out_fwd(h: torch.Tensor, w_bf16: torch.Tensor, bias: torch.Tensor, residual: torch.Tensor,
                ln_weight: torch.Tensor, ln_bias: torch.Tensor, eps: float) -> torch.Tensor:
    """linear2 + residual + clamp + norm3 of TransformerDecoderLayer.forward (dfine_decoder.py:251-253) under
    bf16 autocast as ONE launch (dfine_ffn_out_fwd).  h bfloat16 [..., F]; w_bf16 [C, F]; residual float32 [..., C]."""
    _require_cuda(h, w_bf16, bias, residual, ln_weight, ln_bias)
    C, F = w_bf16.shape
    if (h.dtype != torch.bfloat16 or not h.is_contiguous() or h.shape[-1] != F or residual.dtype != torch.float32
            or not residual.is_contiguous() or residual.shape[-1] != C or w_bf16.dtype != torch.bfloat16
            or not w_bf16.is_contiguous() or residual.numel() // C != h.numel() // F):
        raise ValueError("ffn_out_fwd: need contiguous bfloat16 h [..., F], bfloat16 w [C, F], float32 residual [..., C]")
    M = h.numel() // F
    out = torch.empty_like(residual)
    with torch.cuda.device_of(h), _timed("ffn_out_fwd", h):
        rc = _lib.lib().dfine_ffn_out_fwd(h.data_ptr(), F, w_bf16.data_ptr(), bias.data_ptr(), _dt(bias, "bias"),
                                          residual.data_ptr(), C, ln_weight.data_ptr(), ln_bias.data_ptr(), float(eps),
                                          out.data_ptr(), C, M, C, F, _stream(h))
    check(rc, "dfine_ffn_out_fwd")
    return out


def ffn_fwd_supported(x: torch.Tensor, F: int) -> bool:
    if os.environ.get("DFINE_FFN_FUSED", "1") == "0":      # A/B switch: dfine_linear_fwd(relu) + dfine_ffn_out_fwd
        return False
    return (x.is_cuda and x.dtype == torch.float32 and x.shape[-1] in (128, 256) and F % 128 == 0 and F > 0
            and x.numel() > 0)


def ffn_fwd(x: torch.Tensor, w1_bf16: torch.Tensor, b1: torch.Tensor, w2_bf16: torch.Tensor, b2: torch.Tensor,
            ln_weight: torch.Tensor, ln_bias: torch.Tensor, eps: float) -> torch.Tensor:
    """LayerNorm(clamp(x + linear2(relu(linear1(x))), -65504, 65504)) under bf16 autocast as ONE launch
    (dfine_ffn_fwd: the hidden rows stay on the SM).  x float32 [..., C]; w1_bf16 [F, C], w2_bf16 [C, F] bfloat16;
    b1 [F], b2 [C] in one dtype (float32 / bfloat16)."""
    _require_cuda(x, w1_bf16, b1, w2_bf16, b2, ln_weight, ln_bias)
    C = x.shape[-1]
    F = w1_bf16.shape[0]
    if (x.dtype != torch.float32 or w1_bf16.dtype != torch.bfloat16 or w2_bf16.dtype != torch.bfloat16
            or tuple(w1_bf16.shape) != (F, C) or tuple(w2_bf16.shape) != (C, F) or b1.dtype != b2.dtype
            or not w1_bf16.is_contiguous() or not w2_bf16.is_contiguous() or b1.numel() != F or b2.numel() != C):
        raise ValueError("ffn_fwd: need float32 x [..., C], bfloat16 w1 [F, C] and w2 [C, F], biases of one dtype")
    if not ffn_fwd_supported(x, F):
        raise ValueError(f"ffn_fwd: unsupported shape C = {C}, F = {F} (C in (128, 256), F a multiple of 128)")
    xc = x.contiguous()
    M = xc.numel() // C
    out = torch.empty_like(xc)
    with torch.cuda.device_of(xc), _timed("ffn_fwd", xc):
        rc = _lib.lib().dfine_ffn_fwd(xc.data_ptr(), C, w1_bf16.data_ptr(), b1.contiguous().data_ptr(),
                                      w2_bf16.data_ptr(), b2.contiguous().data_ptr(), _dt(b1, "bias"),
                                      ln_weight.data_ptr(), ln_bias.data_ptr(), float(eps), out.data_ptr(), C, M, C, F,
                                      _stream(xc))
    check(rc, "dfine_ffn_fwd")
    return out


def lqe_fwd(scores: torch.Tensor, pred_corners: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
            b2: torch.Tensor, k: int = 4, reg_max: int = 32, emulate_bf16: bool = False) -> torch.Tensor:
    """LQE.forward of the reference (dfine_decoder.py:307-313) as ONE launch (dfine_lqe_fwd, inference):
    scores [..., num_classes] + MLP(cat(topk(softmax(pred_corners [..., 4*(reg_max+1)]), k), mean)).
    w1 [hidden, 4(k+1)], b1 [hidden], w2 [1, hidden], b2 [1]: float32 parameters of reg_conf."""
    _require_cuda(scores, pred_corners, w1, b1, w2, b2)
    nc = scores.shape[-1]
    N = scores.numel() // nc
    if pred_corners.numel() != N * 4 * (reg_max + 1):
        raise ValueError("lqe_fwd: pred_corners must hold 4*(reg_max+1) logits per row of scores")
    if any(t.dtype != torch.float32 for t in (w1, b1, w2, b2)):
        raise TypeError("lqe_fwd: the MLP's parameters must be float32")
    sc, pc = scores.contiguous(), pred_corners.contiguous()
    out = torch.empty_like(sc)
    with torch.cuda.device_of(sc), _timed("lqe_fwd", sc):
        rc = _lib.lib().dfine_lqe_fwd(pc.data_ptr(), _dt(pc, "pred_corners"), sc.data_ptr(), _dt(sc, "scores"),
                                      w1.contiguous().data_ptr(), b1.contiguous().data_ptr(),
                                      w2.contiguous().data_ptr(), b2.contiguous().data_ptr(), out.data_ptr(), N, nc,
                                      int(k), int(w1.shape[0]), int(reg_max), int(bool(emulate_bf16)), _stream(sc))
    check(rc, "dfine_lqe_fwd")
    return out


def bf16_param(p: torch.Tensor) -> torch.Tensor:
    """A contiguous bfloat16 copy of a parameter, cached ON the parameter object until it is modified in place
    (inference: autocast casts the weights of every Linear on every call; here once).  The cache lives and dies
    with the tensor object -- an address-keyed table would hand a new model the weights of a freed one."""
    if p.dtype == torch.bfloat16 and p.is_contiguous():
        return p
    hit = p.__dict__.get("_dfine_bf16")
    if hit is not None and hit[0] == p._version and hit[1].device == p.device:
        return hit[1]
    t = p.detach().to(torch.bfloat16).contiguous()
    p.__dict__["_dfine_bf16"] = (p._version, t)
    return t


def _packed_params(w0, b0, w1, b1, cdt, x, cache: bool):
    """[w0; w1], [b0; b1] in the compute dtype (dfine_pack_linear).  cache (inference: no parameter needs a
    gradient): the result is kept (on the w0 object, checked against the identity and version of all four
    parameters) until one of them is modified in place.  Training packs on every call, so that a CUDA graph
    of the step contains the pack and replays see the optimizer's updates."""
    ver = (w0._version, b0._version, w1._version, b1._version, cdt, x.device)
    if cache:
        hit = w0.__dict__.get("_dfine_pack")
        if hit is not None and hit[0] == ver and hit[1]() is b0 and hit[2]() is w1 and hit[3]() is b1:
            return hit[4], hit[5]
    n0, n1, K = w0.shape[0], w1.shape[0], w0.shape[1]
    w = torch.empty((n0 + n1, K), dtype=cdt, device=x.device)
    b = torch.empty((n0 + n1,), dtype=cdt, device=x.device)
    with torch.cuda.device_of(x), _timed("pack_linear", x):
        rc = _lib.lib().dfine_pack_linear(w0.data_ptr(), b0.data_ptr(), n0, w1.data_ptr(), b1.data_ptr(),
                                          n1, K, w.data_ptr(), b.data_ptr(), _DT[cdt], _stream(x))
    check(rc, "dfine_pack_linear")
    if cache:
        w0.__dict__["_dfine_pack"] = (ver, weakref.ref(b0), weakref.ref(w1), weakref.ref(b1), w, b)
    return w, b


class _PackedLinearFn(torch.autograd.Function):
    """The two Linears of MSDeformableAttention as ONE GEMM: y = (x [+ x_add]) [W0; W1]^T + [b0; b1].  The
    parameters stay separate tensors (state-dict compatible); they are concatenated and cast to the compute
    dtype by one kernel (dfine_pack_linear).  Under bf16 autocast the forward is dfine_linear_fwd (tcgen05: the
    positional add, the fp32 -> bf16 cast of the query, the GEMM and the bias in one launch; it also emits the
    bf16 operand rows for the backward); otherwise a cuBLAS GEMM.  Backward: cuBLAS for the input gradient,
    dfine_linear_wgrad for dW and db, returned as views of one [N0+N1, K] result."""

    @staticmethod
    def forward(ctx, x, x_add, w0, b0, w1, b1):
        cdt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
        if cdt not in _DT:
            raise TypeError(f"packed_linear: compute dtype {cdt} not supported (see packed_linear_supported)")
        n0, n1 = w0.shape[0], w1.shape[0]
        need_x = any(ctx.needs_input_grad[2:])
        w, b = _packed_params(w0, b0, w1, b1, cdt, x, cache=not need_x)
        xc = x.contiguous()
        ac = x_add.contiguous() if x_add is not None else None
        # dfine_linear_fwd wins where it removes elementwise passes: with the positional rows (measured at
        # config 3, device time: 12.7 us against 17.9 us for add + cast + cuBLAS).  Without them the bf16 copy
        # the cast kernel leaves in L2 makes cast + cuBLAS 2.6 us faster per layer: kept unless
        # DFINE_LINEAR_FWD=always.
        want = ac is not None or os.environ.get("DFINE_LINEAR_FWD") == "always"
        if want and cdt == torch.bfloat16 and xc.dtype == torch.float32 and linear_fwd_supported(xc, n0 + n1, ac):
            if need_x:
                y, x2 = linear_fwd(xc, w, b, x_add=ac, save_input=True)
            else:
                y, x2 = linear_fwd(xc, w, b, x_add=ac), None
        else:
            xs = x if x_add is None else x + x_add
            x2 = xs.reshape(-1, xs.shape[-1]).to(cdt)
            y = torch.nn.functional.linear(x2, w, b).reshape(*x.shape[:-1], n0 + n1)
        ctx.save_for_backward(x2, w)
        ctx.x_shape, ctx.x_dtype, ctx.n0, ctx.has_add = x.shape, x.dtype, n0, x_add is not None
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x2, w = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        if g2.dtype != w.dtype:
            g2 = g2.to(w.dtype)
        gx = _mm(g2, w, ctx.x_dtype).reshape(ctx.x_shape) if (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) \
            else None
        gw = gb = None
        if any(ctx.needs_input_grad[2:]):
            if linear_wgrad_supported(g2, x2):
                gw, gb = linear_wgrad(g2, x2)      # dW and db in one tensor-core launch
            else:
                gw = _mm(g2.t(), x2, torch.float32)
                gb = colsum(g2) if colsum_supported(g2) else g2.float().sum(0)
        n0 = ctx.n0
        if gw is None:
            return gx, (gx if ctx.has_add else None), None, None, None, None
        return gx, (gx if ctx.has_add else None), gw[:n0], gb[:n0], gw[n0:], gb[n0:]


def packed_linear_supported(x, w0, b0, w1, b1) -> bool:
    cdt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else x.dtype
    if cdt not in _DT:      # float16 autocast: the Linear runs in float16 like the reference's (fused_linear)
        return False
    return all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for t in (w0, b0, w1, b1)) \
        and x.is_cuda and w0.shape[1] == w1.shape[1]


def packed_linear(x, w0, b0, w1, b1, x_add=None):
    """(x [+ x_add]) @ [w0; w1]^T + [b0; b1]  (x_add: the decoder layer's query_pos_embed, dfine_decoder.py:245)."""
    return _PackedLinearFn.apply(x, x_add, w0, b0, w1, b1)


def linear_wgrad_supported(gy: torch.Tensor, x: torch.Tensor) -> bool:
    if os.environ.get("DFINE_LINEAR_WGRAD", "1") == "0":     # A/B switch: cuBLAS split-K GEMM + dfine_colsum
        return False
    return (gy.is_cuda and x.is_cuda and gy.dim() == 2 and x.dim() == 2 and gy.shape[0] == x.shape[0]
            and gy.shape[0] > 0 and gy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16
            and gy.stride(1) == 1 and x.stride(1) == 1 and gy.shape[1] % 8 == 0 and x.shape[1] % 8 == 0
            and x.shape[1] <= 256 and gy.stride(0) % 8 == 0 and x.stride(0) % 8 == 0
            and gy.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0)


def linear_wgrad(gy: torch.Tensor, x: torch.Tensor):
    """(gy.t() @ x, gy.sum(0)) in float32 for bf16 gy [M, N], x [M, K <= 256] (dfine_linear_wgrad:
    one tcgen05 launch; the two results are views of one buffer)."""
    _require_cuda(gy, x)
    if not linear_wgrad_supported(gy, x):
        raise ValueError("linear_wgrad: need row-major bf16 gy [M, N], x [M, K] with N, K, strides multiples "
                         "of 8 and K <= 256")
    M, N = gy.shape
    K = x.shape[1]
    ex = _WGRAD_EXCHANGE
    if ex is not None:
        # data parallel (grad_sync.NvlsGradExchange): the gradient is produced in the exchange's local block;
        # end_step() adds all blocks into every rank's replica through the NVLS multicast mapping, and the
        # views returned here (of the replica) hold the rank average after its barrier
        buf, out = ex.next_block(N * K + N)
        with torch.cuda.device_of(gy), _timed("linear_wgrad", gy):
            rc = _lib.lib().dfine_linear_wgrad(gy.data_ptr(), gy.stride(0), x.data_ptr(), x.stride(0), M, N, K,
                                               buf.data_ptr(), _stream(gy))
        check(rc, "dfine_linear_wgrad")
        ex.block_ready()
        return out[:N * K].view(N, K), out[N * K:]
    buf = torch.empty(N * K + N, dtype=torch.float32, device=gy.device)
    with torch.cuda.device_of(gy), _timed("linear_wgrad", gy):
        rc = _lib.lib().dfine_linear_wgrad(gy.data_ptr(), gy.stride(0), x.data_ptr(), x.stride(0), M, N, K,
                                           buf.data_ptr(), _stream(gy))
    check(rc, "dfine_linear_wgrad")
    return buf[:N * K].view(N, K), buf[N * K:]


def colsum_supported(x: torch.Tensor) -> bool:
    return (x.dim() == 2 and x.stride(1) == 1 and x.dtype in _DT and x.shape[1] % 2 == 0
            and x.shape[1] <= 1024 and x.stride(0) % 2 == 0 and x.data_ptr() % 8 == 0)


def colsum(x: torch.Tensor) -> torch.Tensor:
    """Column sums of a row-major 2-D float32 / bfloat16 matrix -> float32 [N] (dfine_colsum)."""
    _require_cuda(x)
    if not colsum_supported(x):
        raise ValueError("colsum: need a row-major 2-D float32/bfloat16 matrix with an even width <= 1024")
    out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    with torch.cuda.device_of(x), _timed("colsum", x):
        rc = _lib.lib().dfine_colsum(x.data_ptr(), _dt(x, "x"), x.shape[0], x.shape[1], x.stride(0),
                                     out.data_ptr(), _stream(x))
    check(rc, "dfine_colsum")
    return out


def _mm(a, b, out_dtype):
    """cuBLAS GEMM with the result written directly in out_dtype (no cast kernel)."""
    if out_dtype == a.dtype:
        return torch.mm(a, b)
    try:
        return torch.mm(a, b, out_dtype=out_dtype)
    except (RuntimeError, TypeError, NotImplementedError):
        return torch.mm(a, b).to(out_dtype)


def fused_linear(x, weight, bias):
    return _LinearFn.apply(x, weight, bias)


def msda_fused_packed(value, value_spatial_shapes, raw, ref_boxes, pts_scale, num_points_list,
                      offset_scale: float = 0.5) -> torch.Tensor:
    """Like msda_fused, but both raw Linear outputs live in one tensor raw [B, Lq, 3HP]
    (offsets first, then logits): the output of a single concatenated Linear."""
    spec = level_spec(value_spatial_shapes, num_points_list)
    memory, H, c, _ = memory_from_value(value, spec)
    _require_cuda(memory, raw, ref_boxes, pts_scale)
    B, Lq = raw.shape[:2]
    if raw.shape[-1] != 3 * H * spec.P:
        raise ValueError(f"raw last dim {raw.shape[-1]} != 3*H*P = {3 * H * spec.P}")
    if raw.dtype not in _DT:
        raw = raw.float()
    ref = ref_boxes.reshape(B, Lq, 4).float().contiguous()
    out_dtype = torch.promote_types(memory.dtype, raw.dtype)
    if torch.is_autocast_enabled():
        out_dtype = torch.float32
    token, sess = _memory_token(memory, value if isinstance(value, torch.Tensor) else None)
    return _MsdaPackedFn.apply(memory if token is None else memory.detach(), raw.contiguous(), ref,
                               pts_scale.float().contiguous(), spec, H, float(offset_scale),
                               out_dtype, token, sess)


def msda_core(value, value_spatial_shapes, sampling_locations, attention_weights,
              num_points_list, method: str = "default") -> torch.Tensor:
    """Drop-in for deformable_attention_core_func_v2 (reference arch/utils.py:191-264).

    value: tuple of [B, H, c, h_l*w_l] views (TransformerDecoder.value_op) or [B, L, H, c];
    sampling_locations [B, Lq, H, P, 2]; attention_weights [B, Lq, H, P] -> [B, Lq, H*c].
    """
    if method != "default":
        raise NotImplementedError(
            "dfine_b200 implements the bilinear ('default') sampler only; "
            "cross_attn_method='discrete' is not used by any shipped config")
    spec = level_spec(value_spatial_shapes, num_points_list)
    memory, H, c, _ = memory_from_value(value, spec)
    _require_cuda(memory, sampling_locations, attention_weights)
    out_dtype = torch.promote_types(memory.dtype, attention_weights.dtype)
    if torch.is_autocast_enabled():
        out_dtype = torch.float32  # grid_sampler is an autocast-to-fp32 op
    loc = sampling_locations.float()
    attn = attention_weights.float()
    token, sess = _memory_token(memory, value if isinstance(value, torch.Tensor) else None)
    return _MsdaFn.apply(memory if token is None else memory.detach(), loc, attn, None, None, spec,
                         H, 0.5, False, out_dtype, token, sess)


def msda_fused(value, value_spatial_shapes, raw_offsets, raw_logits, ref_boxes, pts_scale,
               num_points_list, offset_scale: float = 0.5) -> torch.Tensor:
    """MSDeformableAttention.forward minus the two Linears (reference
    dfine_decoder.py:144-176, reference_points last-dim 4): softmax over the P points,
    sampling-location arithmetic, bilinear gather and weighted sum in one kernel.

    raw_offsets [B, Lq, H*P*2] (or [B,Lq,H,P,2]); raw_logits [B, Lq, H*P];
    ref_boxes [B, Lq, 1, 4] or [B, Lq, 4] (cx, cy, w, h); pts_scale [P] float32.
    """
    spec = level_spec(value_spatial_shapes, num_points_list)
    memory, H, c, _ = memory_from_value(value, spec)
    _require_cuda(memory, raw_offsets, raw_logits, ref_boxes, pts_scale)
    B, Lq = raw_offsets.shape[:2]
    if raw_offsets.dtype != raw_logits.dtype:
        raw_logits = raw_logits.to(raw_offsets.dtype)
    if raw_offsets.dtype not in _DT:
        raw_offsets, raw_logits = raw_offsets.float(), raw_logits.float()
    samp = raw_offsets.reshape(B, Lq, H, spec.P, 2)
    attn = raw_logits.reshape(B, Lq, H, spec.P)
    ref = ref_boxes.reshape(B, Lq, 4).float().contiguous()
    out_dtype = torch.promote_types(memory.dtype, raw_offsets.dtype)
    if torch.is_autocast_enabled():
        out_dtype = torch.float32
    token, sess = _memory_token(memory, value if isinstance(value, torch.Tensor) else None)
    return _MsdaFn.apply(memory if token is None else memory.detach(), samp, attn, ref,
                         pts_scale.float().contiguous(), spec, H, float(offset_scale), True,
                         out_dtype, token, sess)


# --------------------------------------------------------------------------------------
# K3
# --------------------------------------------------------------------------------------
def fdr_project(up: torch.Tensor, reg_scale: torch.Tensor, reg_max: int = 32) -> torch.Tensor:
    """weighting_function (reference arch/utils.py:145-188) in one launch."""
    _require_cuda(up, reg_scale)
    up = up.detach().float().contiguous()
    rs = reg_scale.detach().float().contiguous()
    out = torch.empty(reg_max + 1, dtype=torch.float32, device=up.device)
    with torch.cuda.device_of(up), _timed("fdr_project", up):
        rc = _lib.lib().dfine_fdr_project(up.data_ptr(), rs.data_ptr(), out.data_ptr(), reg_max,
                                          _stream(up))
    check(rc, "dfine_fdr_project")
    return out


_DEV_SCALARS = {}


def _as_dev_scalar(v, like: torch.Tensor) -> torch.Tensor:
    if isinstance(v, torch.Tensor):
        return v.detach().to(device=like.device, dtype=torch.float32).reshape(-1)[:1].contiguous()
    # Python scalars: one device tensor per (value, device), created once -- a fresh torch.tensor(...) would
    # be a pageable host -> device copy on every call (a sync, and illegal under CUDA-graph capture)
    key = (float(v), like.device)
    t = _DEV_SCALARS.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("dfine_b200: scalar constant requested for the first time during CUDA-graph "
                               "capture; run one eager warm-up call first")
        t = _DEV_SCALARS[key] = torch.tensor([float(v)], dtype=torch.float32, device=like.device)
    return t


class _FdrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, corners, ref_init, project, reg_scale, reg_max, want_dist, want_boxes):
        lead = corners.shape[:-1]
        N = corners.numel() // (4 * (reg_max + 1))
        dev = corners.device
        dist = torch.empty((*lead, 4), dtype=torch.float32, device=dev) if want_dist else None
        boxes = torch.empty((*lead, 4), dtype=torch.float32, device=dev) if want_boxes else None
        with torch.cuda.device_of(corners), _timed("fdr_fwd", corners):
            rc = _lib.lib().dfine_fdr_fwd(corners.data_ptr(), _dt(corners, "corners"),
                                          _ptr(ref_init), project.data_ptr(), reg_scale.data_ptr(),
                                          _ptr(dist), _ptr(boxes), N, reg_max, _stream(corners))
        check(rc, "dfine_fdr_fwd")
        ctx.save_for_backward(corners, ref_init, project, reg_scale)
        ctx.reg_max = reg_max
        return dist, boxes

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_dist, g_boxes):
        corners, ref_init, project, reg_scale = ctx.saved_tensors
        N = corners.numel() // (4 * (ctx.reg_max + 1))
        gd = g_dist.float().contiguous() if g_dist is not None else None
        gb = g_boxes.float().contiguous() if g_boxes is not None else None
        gc = torch.empty_like(corners)    # written directly in the corners' dtype
        if gd is None and gb is None:
            return gc.zero_(), None, None, None, None, None, None
        with torch.cuda.device_of(corners), _timed("fdr_bwd", corners):
            rc = _lib.lib().dfine_fdr_bwd(corners.data_ptr(), _dt(corners, "corners"),
                                          _ptr(ref_init), project.data_ptr(), reg_scale.data_ptr(),
                                          _ptr(gb), _ptr(gd), gc.data_ptr(), _dt(gc, "grad_corners"), N,
                                          ctx.reg_max, _stream(corners))
        check(rc, "dfine_fdr_bwd")
        return gc, None, None, None, None, None, None


def _fdr_prepare(corners, project, reg_scale, reg_max):
    _require_cuda(corners, project)
    if corners.dtype not in _DT:
        corners = corners.float()
    if corners.shape[-1] != 4 * (reg_max + 1):
        raise ValueError(f"pred_corners last dim {corners.shape[-1]} != 4*(reg_max+1)")
    return (corners.contiguous(), project.detach().float().contiguous(),
            _as_dev_scalar(reg_scale, corners))


def fdr_integral(corners: torch.Tensor, project: torch.Tensor, reg_max: int = 32) -> torch.Tensor:
    """Drop-in for Integral.forward (reference dfine_decoder.py:291-295): [..., 4*(reg_max+1)]
    -> [..., 4] distances, float32."""
    corners, project, rs = _fdr_prepare(corners, project, 1.0, reg_max)
    dist, _ = _FdrFn.apply(corners, None, project, rs, reg_max, True, False)
    return dist


def fdr_decode(corners: torch.Tensor, ref_init: torch.Tensor, project: torch.Tensor, reg_scale,
               reg_max: int = 32, return_dist: bool = False):
    """distance2bbox(ref_init, Integral(corners, project), reg_scale) in one kernel (reference
    dfine_decoder.py:497-499, arch/utils.py:119-142): -> cxcywh boxes [..., 4] float32."""
    corners, project, rs = _fdr_prepare(corners, project, reg_scale, reg_max)
    _require_cuda(ref_init)
    ref = ref_init.detach().float().contiguous()
    dist, boxes = _FdrFn.apply(corners, ref, project, rs, reg_max, return_dist, True)
    return (boxes, dist) if return_dist else boxes


# --------------------------------------------------------------------------------------
# K4
# --------------------------------------------------------------------------------------
def mask_gemm_raw(coef: torch.Tensor, proto: torch.Tensor, out_dtype: torch.dtype,
                  apply_sigmoid: bool) -> torch.Tensor:
    """coef bf16 [B, M, K] x proto bf16 [B, K, N] -> [B, M, N] on tcgen05 tensor cores."""
    _require_cuda(coef, proto)
    B, M, K = coef.shape
    N = proto.shape[2]
    out = torch.empty((B, M, N), dtype=out_dtype, device=coef.device)
    with torch.cuda.device_of(coef), _timed("mask_gemm", coef):
        rc = _lib.lib().dfine_mask_gemm_fwd(coef.data_ptr(), proto.data_ptr(), out.data_ptr(),
                                            B, M, K, N, _dt(out, "out"), int(apply_sigmoid),
                                            _stream(coef))
    check(rc, "dfine_mask_gemm_fwd")
    return out


def mask_gemm_bwd_supported(K: int, N: int) -> bool:
    return K % 128 == 0 and K <= 256 and N % 8 == 0


def mask_gemm_bwd_raw(coef: torch.Tensor, proto: torch.Tensor, go: torch.Tensor, want_coef: bool = True,
                      want_proto: bool = True, proto_dtype: Optional[torch.dtype] = None):
    """Backward of the mask contraction on tcgen05 tensor cores (autograd of reference
    dfine_decoder.py:940): grad_coef [B, M, K] (float32) = go x proto^T, grad_proto [B, K, N] = coef^T x go.
    coef bf16 [B, M, K], proto bf16 [B, K, N], go bf16 [B, M, N]; returns (grad_coef | None, grad_proto | None)."""
    _require_cuda(coef, proto, go)
    B, M, K = coef.shape
    N = proto.shape[2]
    if go.dtype != torch.bfloat16 or coef.dtype != torch.bfloat16 or proto.dtype != torch.bfloat16:
        raise TypeError("mask_gemm_bwd: coef, proto and grad_out must be bfloat16")
    go = go.contiguous()
    g_coef = torch.empty((B, M, K), dtype=torch.float32, device=coef.device) if want_coef else None
    g_proto = torch.empty((B, K, N), dtype=proto_dtype or torch.bfloat16, device=coef.device) if want_proto else None
    with torch.cuda.device_of(coef), _timed("mask_gemm_bwd", coef, kernels=int(want_coef) + int(want_proto)):
        rc = _lib.lib().dfine_mask_gemm_bwd(coef.data_ptr(), proto.data_ptr(), go.data_ptr(), _ptr(g_coef),
                                            _ptr(g_proto), B, M, K, N,
                                            _dt(g_proto, "grad_proto") if want_proto else _lib.BF16, _stream(coef))
    check(rc, "dfine_mask_gemm_bwd")
    return g_coef, g_proto


class _MaskFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coef, proto, out_dtype, apply_sigmoid):
        out = mask_gemm_raw(coef, proto, out_dtype, apply_sigmoid)
        ctx.save_for_backward(coef, proto, out if apply_sigmoid else None)
        ctx.apply_sigmoid = apply_sigmoid
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        coef, proto, out = ctx.saved_tensors
        if ctx.apply_sigmoid:
            go = go * out * (1 - out)
        go = go.to(torch.bfloat16)          # the dtype autocast's bmm backward would compute in
        want_coef, want_proto = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        g_coef, g_proto = mask_gemm_bwd_raw(coef, proto, go, want_coef, want_proto, proto.dtype)
        if g_coef is not None:
            g_coef = g_coef.to(coef.dtype)
        return g_coef, g_proto, None, None


def mask_logits(coef: torch.Tensor, mask_feat: torch.Tensor, apply_sigmoid: bool = False,
                out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Drop-in for einsum("bqc,bchw->bqhw") (reference dfine_decoder.py:937-940).

    Computes in bf16 with fp32 accumulation on the tensor cores, like the reference under
    torch.autocast.  coef [B, Q, C], mask_feat [B, C, h, w] -> [B, Q, h, w].
    """
    B, Q, C = coef.shape
    hh, ww = mask_feat.shape[-2:]
    a = coef.to(torch.bfloat16).contiguous()
    b = mask_feat.to(torch.bfloat16).reshape(B, C, hh * ww).contiguous()
    out = _MaskFn.apply(a, b, out_dtype or torch.bfloat16, apply_sigmoid)
    return out.reshape(B, Q, hh, ww)


# --------------------------------------------------------------------------------------
# K5  Hungarian matching on the device
# --------------------------------------------------------------------------------------
def lsap(cost: torch.Tensor, n_targets: Sequence[int]):
    """Batched rectangular assignment, index-identical to scipy.optimize.linear_sum_assignment applied
    to torch.nan_to_num(cost[b, :, :n_targets[b]], nan=1.0) (reference src/d_fine/matcher.py:112-116).

    cost: float32 CUDA tensor [B, Q, T] (any strides), n_targets: per-image target counts (host ints).
    Returns (q_idx, t_idx): int64 CUDA tensors [B, K], K = max_b min(Q, n_targets[b]); pair k of image b is
    (q_idx[b, k], t_idx[b, k]) in ascending query order, -1 beyond min(Q, n_targets[b])."""
    _require_cuda(cost)
    if cost.dtype != torch.float32 or cost.dim() != 3:
        raise TypeError("lsap: cost must be a float32 tensor [B, Q, T]")
    B, Q, T = cost.shape
    n = [int(v) for v in n_targets]
    if len(n) != B or any(v < 0 or v > T for v in n):
        raise ValueError(f"lsap: n_targets must hold {B} counts in [0, {T}]")
    K = max(1, max(min(Q, v) for v in n))
    out = torch.empty((2, B, K), dtype=torch.int64, device=cost.device)
    with torch.cuda.device_of(cost), _timed("lsap", cost):
        rc = _lib.lib().dfine_lsap(cost.data_ptr(), cost.stride(0), cost.stride(1), cost.stride(2),
                                   _lib.i32_array(n), B, Q, out[0].data_ptr(), out[1].data_ptr(), K, _stream(cost))
    check(rc, "dfine_lsap")
    return out[0], out[1]


# --------------------------------------------------------------------------------------
# K6  mask loss over the matched rows (fused focal-BCE + dice statistics)
# --------------------------------------------------------------------------------------
class _MaskLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, tgt):
        M, N = logits.shape
        stats = torch.empty((M, 4), dtype=torch.float32, device=logits.device)
        with torch.cuda.device_of(logits), _timed("mask_loss_fwd", logits):
            rc = _lib.lib().dfine_mask_loss_fwd(logits.data_ptr(), _dt(logits, "logits"), logits.stride(0),
                                                tgt.data_ptr(), M, N, stats.data_ptr(), _stream(logits))
        check(rc, "dfine_mask_loss_fwd")
        ctx.save_for_backward(logits, tgt, stats)
        return stats

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_stats):
        logits, tgt, stats = ctx.saved_tensors
        M, N = logits.shape
        g_stats = g_stats.to(torch.float32).contiguous()
        grad = torch.empty((M, N), dtype=logits.dtype, device=logits.device)
        with torch.cuda.device_of(logits), _timed("mask_loss_bwd", logits):
            rc = _lib.lib().dfine_mask_loss_bwd(logits.data_ptr(), _dt(logits, "logits"), logits.stride(0),
                                                tgt.data_ptr(), M, N, stats.data_ptr(), g_stats.data_ptr(),
                                                grad.data_ptr(), _dt(grad, "grad"), _stream(logits))
        check(rc, "dfine_mask_loss_bwd")
        return grad, None


def mask_loss_stats(logits: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """Per-row statistics of the mask loss (reference dfine_criterion.py:273-312) for logits [M, N]
    (float32 / bfloat16) and targets [M, N] (float32): [M, 4] = {sum focal, sum p t, sum p, sum t}."""
    _require_cuda(logits, tgt)
    if logits.dim() != 2 or tgt.shape != logits.shape:
        raise ValueError("mask_loss_stats: logits and tgt must be [M, N] tensors of the same shape")
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    tgt = tgt.to(torch.float32).contiguous()
    return _MaskLossFn.apply(logits, tgt)


def mask_losses(logits: torch.Tensor, tgt: torch.Tensor, eps: float = 1e-6):
    """(loss_mask_bce, loss_mask_dice) of DFINECriterion._focal_loss_mask / _dice_loss
    (dfine_criterion.py:273-312) for pred_sel [M, h, w] and tgt_sel [M, h, w]."""
    M = logits.shape[0]
    st = mask_loss_stats(logits.reshape(M, -1), tgt.reshape(M, -1))
    n = logits[0].numel()
    bce = (st[:, 0] / n).mean()
    dice = (1.0 - (2.0 * st[:, 1] + eps) / (st[:, 2] + st[:, 3] + eps)).mean()
    return bce, dice
