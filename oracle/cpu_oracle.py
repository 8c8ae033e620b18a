"""ctypes/numpy front end of oracle/dfine_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package (d-fine-seg_b200/dfine_b200) never
does.  See the header of dfine_oracle.c for what is restated and how it is pinned.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libdfine_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    """Compile dfine_oracle.c with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "dfine_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _f(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def _i(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_i32p)


def _c32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def level_tables(spatial_shapes: Sequence[Sequence[int]], num_points: Sequence[int]):
    """(lvl_hw [n,2], lvl_start [n], lvl_npts [n]) as int32 arrays."""
    hw = np.ascontiguousarray(np.asarray(spatial_shapes, dtype=np.int32).reshape(-1, 2))
    sizes = hw[:, 0].astype(np.int64) * hw[:, 1]
    start = np.ascontiguousarray(np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32))
    npts = np.ascontiguousarray(np.asarray(num_points, dtype=np.int32))
    assert len(npts) == len(hw)
    return hw, start, npts


def msda_fwd(value, spatial_shapes, num_points, loc, attn, want_idx: bool = False):
    """value [B,L,H,c]; loc [B,Lq,H,P,2]; attn [B,Lq,H,P] -> out [B,Lq,H*c] (, idx, wts)."""
    value, loc, attn = _c32(value), _c32(loc), _c32(attn)
    B, L, H, c = value.shape
    Lq, P = loc.shape[1], loc.shape[3]
    hw, start, npts = level_tables(spatial_shapes, num_points)
    assert int(npts.sum()) == P and int((hw[:, 0] * hw[:, 1]).sum()) == L
    out = np.empty((B, Lq, H * c), np.float32)
    idx = np.empty((B, Lq, H, P, 4), np.int32) if want_idx else None
    wts = np.empty((B, Lq, H, P, 4), np.float32) if want_idx else None
    rc = lib().oracle_msda_fwd(_f(value), B, L, H, c, len(npts), _i(hw), _i(start), _i(npts),
                               _f(loc), _f(attn), Lq, _f(out),
                               _i(idx) if want_idx else None, _f(wts) if want_idx else None)
    assert rc == 0, rc
    return (out, idx, wts) if want_idx else out


def msda_bwd(value, spatial_shapes, num_points, loc, attn, grad_out):
    """-> grad_value [B,L,H,c], grad_loc [B,Lq,H,P,2], grad_attn [B,Lq,H,P]."""
    value, loc, attn, grad_out = _c32(value), _c32(loc), _c32(attn), _c32(grad_out)
    B, L, H, c = value.shape
    Lq, P = loc.shape[1], loc.shape[3]
    hw, start, npts = level_tables(spatial_shapes, num_points)
    gv = np.empty_like(value)
    gl = np.empty_like(loc)
    ga = np.empty_like(attn)
    rc = lib().oracle_msda_bwd(_f(value), B, L, H, c, len(npts), _i(hw), _i(start), _i(npts),
                               _f(loc), _f(attn), Lq, _f(grad_out), _f(gv), _f(gl), _f(ga))
    assert rc == 0, rc
    return gv, gl, ga


def msda_locations(raw, ref_boxes, pts_scale, offset_scale: float = 0.5):
    raw, ref_boxes, pts_scale = _c32(raw), _c32(ref_boxes), _c32(pts_scale)
    B, Lq, H, P, _ = raw.shape
    out = np.empty_like(raw)
    lib().oracle_msda_locations(_f(raw), _f(ref_boxes.reshape(B, Lq, 4)), _f(pts_scale),
                                ctypes.c_float(offset_scale), B, Lq, H, P, _f(out))
    return out


def msda_locations_bwd(grad_loc, ref_boxes, pts_scale, offset_scale: float = 0.5):
    grad_loc, ref_boxes, pts_scale = _c32(grad_loc), _c32(ref_boxes), _c32(pts_scale)
    B, Lq, H, P, _ = grad_loc.shape
    out = np.empty_like(grad_loc)
    lib().oracle_msda_locations_bwd(_f(grad_loc), _f(ref_boxes.reshape(B, Lq, 4)), _f(pts_scale),
                                    ctypes.c_float(offset_scale), B, Lq, H, P, _f(out))
    return out


def softmax(x):
    x = _c32(x)
    n = x.shape[-1]
    y = np.empty_like(x)
    lib().oracle_softmax(_f(x), ctypes.c_size_t(x.size // n), n, _f(y))
    return y


def softmax_bwd(y, gy):
    y, gy = _c32(y), _c32(gy)
    n = y.shape[-1]
    gx = np.empty_like(y)
    lib().oracle_softmax_bwd(_f(y), _f(gy), ctypes.c_size_t(y.size // n), n, _f(gx))
    return gx


def msda_fused_fwd(value, spatial_shapes, num_points, raw_offsets, raw_logits, ref_boxes,
                   pts_scale, offset_scale: float = 0.5):
    """MSDeformableAttention.forward minus the two Linears (dfine_decoder.py:144-176)."""
    loc = msda_locations(raw_offsets, ref_boxes, pts_scale, offset_scale)
    attn = softmax(raw_logits)
    return msda_fwd(value, spatial_shapes, num_points, loc, attn)


def msda_fused_bwd(value, spatial_shapes, num_points, raw_offsets, raw_logits, ref_boxes,
                   pts_scale, grad_out, offset_scale: float = 0.5):
    loc = msda_locations(raw_offsets, ref_boxes, pts_scale, offset_scale)
    attn = softmax(raw_logits)
    gv, gl, ga = msda_bwd(value, spatial_shapes, num_points, loc, attn, grad_out)
    return gv, msda_locations_bwd(gl, ref_boxes, pts_scale, offset_scale), softmax_bwd(attn, ga)


def fdr_project(up: float, reg_scale: float, reg_max: int = 32) -> np.ndarray:
    out = np.empty(reg_max + 1, np.float32)
    rc = lib().oracle_fdr_project(ctypes.c_float(up), ctypes.c_float(reg_scale), reg_max, _f(out))
    assert rc == 0, rc
    return out


def fdr_fwd(corners, ref_init, project, reg_scale: float, reg_max: int = 32
            ) -> Tuple[np.ndarray, np.ndarray]:
    """-> (dist [...,4], boxes [...,4])."""
    corners, ref_init, project = _c32(corners), _c32(ref_init), _c32(project)
    lead = corners.shape[:-1]
    N = int(np.prod(lead)) if lead else 1
    dist = np.empty((N, 4), np.float32)
    boxes = np.empty((N, 4), np.float32)
    rc = lib().oracle_fdr_fwd(_f(corners), _f(ref_init), _f(project), ctypes.c_float(reg_scale),
                              _f(dist), _f(boxes), ctypes.c_size_t(N), reg_max)
    assert rc == 0, rc
    return dist.reshape(*lead, 4), boxes.reshape(*lead, 4)


def fdr_bwd(corners, ref_init, project, reg_scale: float, grad_boxes: Optional[np.ndarray],
            grad_dist: Optional[np.ndarray] = None, reg_max: int = 32) -> np.ndarray:
    corners, ref_init, project = _c32(corners), _c32(ref_init), _c32(project)
    N = corners.size // (4 * (reg_max + 1))
    gb = _c32(grad_boxes) if grad_boxes is not None else None
    gd = _c32(grad_dist) if grad_dist is not None else None
    gc = np.empty_like(corners)
    rc = lib().oracle_fdr_bwd(_f(corners), _f(ref_init), _f(project), ctypes.c_float(reg_scale),
                              _f(gb) if gb is not None else None,
                              _f(gd) if gd is not None else None, _f(gc),
                              ctypes.c_size_t(N), reg_max)
    assert rc == 0, rc
    return gc


def mask_gemm(coef, proto, apply_sigmoid: bool = False) -> np.ndarray:
    """coef [B,M,K], proto [B,K,N] (or [B,K,h,w]) -> [B,M,N] (or [B,M,h,w])."""
    coef, proto = _c32(coef), _c32(proto)
    B, M, K = coef.shape
    tail = proto.shape[2:]
    N = int(np.prod(tail))
    out = np.empty((B, M, N), np.float32)
    rc = lib().oracle_mask_gemm(_f(coef), _f(proto.reshape(B, K, N)), _f(out), B, M, K, N,
                                int(apply_sigmoid))
    assert rc == 0, rc
    return out.reshape(B, M, *tail)


def mask_gemm_bwd(coef, proto, grad_out):
    """Gradients of out[b,m,n] = sum_k coef[b,m,k] proto[b,k,n] (autograd of the einsum at reference
    src/d_fine/arch/dfine_decoder.py:940): grad_coef = grad_out x proto^T, grad_proto = coef^T x grad_out.
    Plain float64 numpy (test infrastructure: the checker of dfine_mask_gemm_bwd)."""
    coef, proto, go = (np.asarray(a, dtype=np.float64) for a in (coef, proto, grad_out))
    B, M, K = coef.shape
    tail = proto.shape[2:]
    p = proto.reshape(B, K, -1)
    g = go.reshape(B, M, -1)
    return np.einsum("bmn,bkn->bmk", g, p), np.einsum("bmk,bmn->bkn", coef, g).reshape(B, K, *tail)


def mask_loss(pred, tgt, eps: float = 1e-6):
    """Focal-BCE and dice loss over matched mask rows, and their gradients w.r.t. the logits (reference
    src/d_fine/dfine_criterion.py:273-312), float64 numpy: returns (loss_bce, loss_dice, grad_bce, grad_dice,
    stats [M, 4] = {sum focal, sum p t, sum p, sum t}).  Test infrastructure: the checker of
    dfine_mask_loss_fwd / _bwd."""
    x = np.asarray(pred, np.float64).reshape(pred.shape[0], -1)
    t = np.asarray(tgt, np.float64).reshape(tgt.shape[0], -1)
    M, N = x.shape
    fg = t.mean(1, keepdims=True)
    alpha = 0.5 + 0.25 * np.clip(1 - 2 * fg, -1, 1)
    p = 1 / (1 + np.exp(-x))
    bce = np.maximum(x, 0) - x * t + np.log1p(np.exp(-np.abs(x)))
    pt = p * t + (1 - p) * (1 - t)
    at = alpha * t + (1 - alpha) * (1 - t)
    focal = at * (1 - pt) ** 2 * bce
    stats = np.stack([focal.sum(1), (p * t).sum(1), p.sum(1), t.sum(1)], 1)
    loss_bce = (stats[:, 0] / N).mean()
    den = stats[:, 2] + stats[:, 3] + eps
    loss_dice = (1 - (2 * stats[:, 1] + eps) / den).mean()
    pq = p * (1 - p)
    dfocal = at * ((1 - pt) ** 2 * (p - t) - 2 * (1 - pt) * (2 * t - 1) * pq * bce)
    grad_bce = dfocal / (N * M)
    g_inter = (-2 / den / M)[:, None]
    g_psum = ((2 * stats[:, 1] + eps) / den ** 2 / M)[:, None]
    grad_dice = (g_inter * t + g_psum) * pq
    return loss_bce, loss_dice, grad_bce.reshape(pred.shape), grad_dice.reshape(pred.shape), stats


def lsap(cost) -> Tuple[np.ndarray, np.ndarray]:
    """Rectangular linear-sum assignment of one float32 cost matrix [queries, targets] exactly as
    scipy.optimize.linear_sum_assignment solves it (reference src/d_fine/matcher.py:115): returns
    (query indices ascending, target indices), int64."""
    c = _c32(cost)
    nq, nt = c.shape
    k = min(nq, nt)
    oq, ot = np.empty(k, np.int64), np.empty(k, np.int64)
    fn = lib().oracle_lsap
    fn.restype = ctypes.c_int
    fn.argtypes = [_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    rc = fn(_f(c), nq, nt, nt, oq.ctypes.data, ot.ctypes.data)
    assert rc == k, rc
    return oq, ot


def linear_wgrad(grad_y: np.ndarray, x: np.ndarray):
    """Weight and bias gradient of y = x W^T + b (autograd of the two nn.Linear of MSDeformableAttention,
    reference src/d_fine/arch/dfine_decoder.py:87-88, :139-147): dW = grad_y^T x, db = sum over the rows of
    grad_y.  Plain float64 numpy (test infrastructure: the checker of dfine_linear_wgrad)."""
    g = np.asarray(grad_y, dtype=np.float64).reshape(-1, grad_y.shape[-1])
    xx = np.asarray(x, dtype=np.float64).reshape(-1, x.shape[-1])
    return g.T @ xx, g.sum(0)


# ------------------------------------------------------------------------------------------
# The decoder layer's Linears under torch.autocast(bfloat16): numpy restatement of the reference's
# op sequence with its rounding points (test infrastructure: the checker of dfine_linear_fwd,
# dfine_gate_fwd, dfine_ffn_out_fwd; pinned to tests/golden/layer.npz by tests/test_oracle_golden.py)
# ------------------------------------------------------------------------------------------
def bf16_round(x) -> np.ndarray:
    """float32 -> nearest bfloat16 (ties to even) -> float32, the cast autocast applies to the operands and
    that F.linear applies to its result."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = r.view(np.float32).copy()
    nan = np.isnan(a)
    out[nan] = a[nan]
    return out.reshape(a.shape)


def linear_bf16(x, w, b, x_add=None, relu: bool = False) -> np.ndarray:
    """F.linear(x [+ x_add], w, b) under autocast(bfloat16) (reference dfine_decoder.py:139-147 with the
    caller's with_pos_embed :245; linear1 + activation :229-230): the add in float32, operands rounded to
    bf16, exact products accumulated (float64 here; fp32 in an unspecified order on the device), bias
    (rounded to bf16) added before the single rounding of the result to bf16."""
    xs = np.asarray(x, np.float32)
    if x_add is not None:
        xs = xs + np.asarray(x_add, np.float32)
    y = bf16_round(xs).astype(np.float64) @ bf16_round(w).astype(np.float64).T + bf16_round(b).astype(np.float64)
    y = bf16_round(y.astype(np.float32))
    return np.maximum(y, 0.0) if relu else y


def layer_norm(x, w, b, eps: float) -> np.ndarray:
    x = np.asarray(x, np.float64)
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return ((x - mu) / np.sqrt(var + eps) * np.asarray(w, np.float64) + np.asarray(b, np.float64)).astype(np.float32)


def gate_fwd(x1, x2, w, b, ln_w, ln_b, eps: float) -> np.ndarray:
    """Gate.forward (reference dfine_decoder.py:265-271) under autocast(bfloat16): gates =
    sigmoid(Linear(cat(x1, x2))) is a bf16 tensor (sigmoid evaluated in float32 on the bf16 Linear output,
    rounded once); gate * x promotes to float32; LayerNorm in float32."""
    C = np.asarray(x1).shape[-1]
    z = linear_bf16(np.concatenate([x1, x2], -1), w, b)
    with np.errstate(over="ignore"):
        gates = bf16_round((1.0 / (1.0 + np.exp(-z.astype(np.float64)))).astype(np.float32))
    mix = (gates[..., :C] * np.asarray(x1, np.float32)) + (gates[..., C:] * np.asarray(x2, np.float32))
    return layer_norm(mix, ln_w, ln_b, eps)


def ffn_tail(hidden, w2, b2, residual, ln_w, ln_b, eps: float) -> np.ndarray:
    """The last lines of TransformerDecoderLayer.forward (reference dfine_decoder.py:251-253) under
    autocast(bfloat16): linear2 (bf16 result) + residual in float32, clamp to +-65504, norm3."""
    t2 = linear_bf16(hidden, w2, b2)
    return layer_norm(np.clip(np.asarray(residual, np.float32) + t2, -65504.0, 65504.0), ln_w, ln_b, eps)


def lqe_fwd(scores, corners, w1, b1, w2, b2, k: int = 4, reg_max: int = 32, emulate_bf16: bool = False) -> np.ndarray:
    """LQE.forward (reference dfine_decoder.py:307-313; MLP :33-46): softmax over the reg_max+1 bins of the four
    edges, the k largest probabilities per edge and their mean, reg_conf = Linear(4(k+1), H) -> ReLU -> Linear(H, 1),
    broadcast add to the scores.  emulate_bf16: the rounding points of autocast(bfloat16) (statistics, parameters
    and each Linear's result rounded to bf16; the sum with the bf16 scores rounded once)."""
    sc = np.asarray(scores, np.float32)
    x = np.asarray(corners, np.float32).reshape(*sc.shape[:-1], 4, reg_max + 1).astype(np.float64)
    e = np.exp(x - x.max(-1, keepdims=True))
    prob = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    top = -np.sort(-prob, axis=-1)[..., :k]
    stat = np.concatenate([top, top.mean(-1, keepdims=True, dtype=np.float32)], -1).reshape(*sc.shape[:-1], 4 * (k + 1))
    r = bf16_round if emulate_bf16 else (lambda a: np.asarray(a, np.float32))
    h = r((r(stat).astype(np.float64) @ r(w1).astype(np.float64).T + r(b1).astype(np.float64)).astype(np.float32))
    h = np.maximum(h, 0.0)
    q = r((h.astype(np.float64) @ r(w2).astype(np.float64).T + r(b2).astype(np.float64)).astype(np.float32))
    return r(sc + q)
