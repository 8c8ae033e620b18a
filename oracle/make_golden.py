"""Generates tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (the reference tree is not shipped to the GPU box):

    python oracle/make_golden.py [--ref /root/reference] [--out tests/golden]

Every fixture holds the seeded inputs AND the outputs the reference produced for them
(torch CPU, float32), so the tests never need the reference or a particular RNG.
Tensors that feed both the fp32 and the bf16 kernels are rounded to bf16-representable
values first, so one fixture serves both element types with identical inputs.

Reference entry points exercised (all imported, nothing restated here):
  src/d_fine/arch/utils.py:191   deformable_attention_core_func_v2
  src/d_fine/arch/dfine_decoder.py:49   MSDeformableAttention (forward :119-178)
  src/d_fine/arch/dfine_decoder.py:416  TransformerDecoder.value_op
  src/d_fine/arch/utils.py:145   weighting_function
  src/d_fine/arch/dfine_decoder.py:274  Integral
  src/d_fine/arch/utils.py:119   distance2bbox
  src/d_fine/arch/dfine_decoder.py:937  DFINETransformer._mask_logits_from_h
  src/d_fine/dfine_criterion.py:273-312  DFINECriterion._focal_loss_mask / _dice_loss
  src/d_fine/arch/dfine_decoder.py:180-256  TransformerDecoderLayer (with_pos_embed, forward_ffn, norm3), :258-271 Gate
  src/d_fine/arch/dfine_decoder.py:298-313  LQE
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch


def bf16r(t: torch.Tensor) -> torch.Tensor:
    """Round to the nearest bf16-representable float32."""
    return t.to(torch.bfloat16).to(torch.float32)


def make_boxes(g: torch.Generator, B: int, Lq: int) -> torch.Tensor:
    """Object-like cxcywh boxes plus ~10 % border / oversize boxes."""
    cxy = torch.rand(B, Lq, 2, generator=g) * 0.9 + 0.05
    wh = torch.exp(torch.rand(B, Lq, 2, generator=g) * (np.log(0.6) - np.log(0.02)) + np.log(0.02))
    big = torch.rand(B, Lq, 1, generator=g) < 0.10
    wh_big = torch.rand(B, Lq, 2, generator=g) * 0.6 + 0.6
    cxy_edge = torch.rand(B, Lq, 2, generator=g) * 1.2 - 0.1
    wh = torch.where(big, wh_big, wh)
    cxy = torch.where(big, cxy_edge, cxy)
    return torch.cat([cxy, wh], -1)


def ref_value_views(TransformerDecoder, memory: torch.Tensor, H: int, shapes):
    stub = types.SimpleNamespace(num_head=H)
    return TransformerDecoder.value_op(stub, memory, None, None, None, shapes)


def case_core(ref, name, seed, B, Lq, H, c, shapes, npts, loc_override=None, attn_override=None):
    """deformable_attention_core_func_v2 forward + all three gradients."""
    g = torch.Generator().manual_seed(seed)
    L = sum(h * w for h, w in shapes)
    P = sum(npts)
    memory = bf16r(torch.randn(B, L, H * c, generator=g)).requires_grad_(True)
    if loc_override is None:
        boxes = make_boxes(g, B, Lq)
        off = torch.randn(B, Lq, H, P, 2, generator=g) * 0.6
        loc = boxes[:, :, None, None, :2] + off * boxes[:, :, None, None, 2:] * 0.5
    else:
        loc = loc_override
    loc = loc.clone().requires_grad_(True)
    if attn_override is None:
        attn = torch.softmax(torch.randn(B, Lq, H, P, generator=g) * 1.5, -1)
    else:
        attn = attn_override
    attn = attn.clone().requires_grad_(True)
    grad_out = bf16r(torch.randn(B, Lq, H * c, generator=g))

    value = ref_value_views(ref.TransformerDecoder, memory, H, shapes)
    out = ref.core(value, shapes, loc, attn, npts)
    assert out.shape == (B, Lq, H * c)
    out.backward(grad_out)
    return name, dict(
        memory_bf16=memory.detach().to(torch.bfloat16).view(torch.int16).numpy(),
        loc=loc.detach().numpy(), attn=attn.detach().numpy(),
        grad_out_bf16=grad_out.to(torch.bfloat16).view(torch.int16).numpy(),
        shapes=np.asarray(shapes, np.int32), npts=np.asarray(npts, np.int32),
        H=np.int32(H), c=np.int32(c),
        out=out.detach().contiguous().numpy(),
        grad_memory=memory.grad.numpy(), grad_loc=loc.grad.numpy(), grad_attn=attn.grad.numpy(),
    )


def case_edge(ref):
    """Engineered positions: pixel centres, half-pixel borders, far outside, NaN / inf."""
    shapes, npts, H, c = [[5, 7], [3, 4]], [4, 4], 2, 16
    vals = []
    for (h, w) in shapes:
        xs = [(k + 0.5) / w for k in range(-1, w + 1)] + [0.0, 1.0, -1.0 / w, 1.0 + 1.0 / w,
                                                          0.5 / w * 0.999999, 1e-8, 1 - 1e-8]
        ys = [(k + 0.5) / h for k in range(-1, h + 1)] + [0.0, 1.0, 0.5, 0.25]
        vals.append((xs, ys))
    pts = []
    for xs, ys in vals:
        lv = [(x, y) for x in xs for y in ys]
        pts.append(lv)
    n = max(len(p) for p in pts)
    Lq = (n + 3) // 4 + 2
    loc = torch.full((1, Lq, H, 8, 2), 0.5)
    for lvl, lv in enumerate(pts):
        for i, (x, y) in enumerate(lv):
            q, p = divmod(i, 4)
            loc[0, q, :, lvl * 4 + p, 0] = x
            loc[0, q, :, lvl * 4 + p, 1] = y
    special = torch.tensor([[float("nan"), 0.5], [0.5, float("inf")], [-float("inf"), 0.5],
                            [1e30, 0.5], [0.5, -1e30], [3.0, 3.0], [-2.0, 0.5], [0.5, 2.5]])
    loc[0, Lq - 2, 0, :, :] = special
    loc[0, Lq - 1, 1, :, :] = special.flip(0)
    g = torch.Generator().manual_seed(77)
    attn = torch.softmax(torch.randn(1, Lq, H, 8, generator=g), -1)
    attn[0, 0, 0] = 0.0
    attn[0, 0, 0, 3] = 1.0  # one-hot attention
    # Only the forward is pinned for the NaN/inf rows: autograd through NaN positions
    # yields NaN gradients in the reference, checked separately by the finite mask.
    return case_core(ref, "core_edge", 5, 1, Lq, H, c, shapes, npts, loc_override=loc,
                     attn_override=attn)


def case_near_centre(ref):
    """Positions within a few ulps of pixel centres at the real 640^2 level sizes (80, 40, 20): there the
    unnormalise ((g+1)*size-1)/2 gives floor() one value when the multiply-subtract is rounded once (what
    the shipped ATen kernels execute: FFMA on CUDA, vfmsub on CPU) and another when it is rounded twice.
    The gradient wrt the location and the touched pixels of grad_memory expose the chosen cell.  A freshly
    initialised decoder samples exactly such positions (offset bias rings on pixel centres)."""
    shapes, npts, H, c = [[80, 80], [40, 40], [20, 20]], [4, 4, 4], 2, 16
    per_level = []
    for (h, w) in shapes:
        xs = []
        for k in range(w):
            base = np.float32((k + 0.5) / w)
            for d in range(-3, 4):
                v = base
                for _ in range(abs(d)):
                    v = np.nextafter(v, np.float32(2.0 if d > 0 else -2.0), dtype=np.float32)
                xs.append(float(v))
        per_level.append(xs)
    n = max(len(x) for x in per_level)
    Lq = (n + 3) // 4
    loc = torch.full((1, Lq, H, 12, 2), 0.5)
    for lvl, xs in enumerate(per_level):
        for i, x in enumerate(xs):
            q, p = divmod(i, 4)
            # head 0: x sweeps the near-centre values, y sits mid-cell; head 1: the transpose
            loc[0, q, 0, lvl * 4 + p, 0] = x
            loc[0, q, 0, lvl * 4 + p, 1] = xs[(i * 7 + 3) % len(xs)] * 0.5 + 0.2
            loc[0, q, 1, lvl * 4 + p, 1] = x
            loc[0, q, 1, lvl * 4 + p, 0] = xs[(i * 5 + 1) % len(xs)] * 0.5 + 0.3
    g = torch.Generator().manual_seed(78)
    attn = torch.softmax(torch.randn(1, Lq, H, 12, generator=g), -1)
    return case_core(ref, "core_near_centre", 6, 1, Lq, H, c, shapes, npts, loc_override=loc,
                     attn_override=attn)


def case_probe(ref):
    """Which pixels does aten::grid_sampler_2d touch?  One sample per (image, head):
    the non-zeros of grad_input are the in-bounds corners, their values the weights."""
    h, w = 6, 9
    g = torch.Generator().manual_seed(9)
    xs = [(k + 0.5) / w for k in range(-2, w + 2)] + [k / w for k in range(-1, w + 2)]
    ys = [(k + 0.5) / h for k in range(-2, h + 2)] + [k / h for k in range(-1, h + 2)]
    eng = torch.tensor([(x, y) for x in xs for y in ys], dtype=torch.float32)
    rnd = torch.rand(400, 2, generator=g) * 1.3 - 0.15
    loc = torch.cat([eng, rnd], 0)
    n = loc.shape[0]
    inp = torch.ones(n, 1, h, w, requires_grad=True)
    grid = (2 * loc - 1).reshape(n, 1, 1, 2)
    out = torch.nn.functional.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros",
                                          align_corners=False)
    out.sum().backward()
    return "aten_corner_probe", dict(loc=loc.numpy(), hw=np.asarray([h, w], np.int32),
                                     grad_input=inp.grad.reshape(n, h * w).numpy(),
                                     out=out.detach().reshape(n).numpy())


def case_module(ref, name, seed, B, Lq, C, H, shapes, npts):
    """MSDeformableAttention.forward with trained-like Linear weights; grads wrt
    query, memory and the four Linear parameters.  Also records the raw Linear
    outputs so that the fused kernel can be checked without the module."""
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    m = ref.MSDeformableAttention(C, H, len(shapes), npts)
    with torch.no_grad():
        m.sampling_offsets.weight.copy_(bf16r(torch.randn(m.sampling_offsets.weight.shape, generator=g) * 0.02))
        m.attention_weights.weight.copy_(bf16r(torch.randn(m.attention_weights.weight.shape, generator=g) * 0.05))
        m.attention_weights.bias.copy_(bf16r(torch.randn(m.attention_weights.bias.shape, generator=g) * 0.1))
        m.sampling_offsets.bias.copy_(bf16r(m.sampling_offsets.bias))
    L = sum(h * w for h, w in shapes)
    P = sum(npts)
    memory = bf16r(torch.randn(B, L, C, generator=g)).requires_grad_(True)
    query = bf16r(torch.randn(B, Lq, C, generator=g)).requires_grad_(True)
    ref_pts = make_boxes(g, B, Lq).unsqueeze(2)  # [B, Lq, 1, 4]
    grad_out = bf16r(torch.randn(B, Lq, C, generator=g))
    value = ref_value_views(ref.TransformerDecoder, memory, H, shapes)
    assert not value[0].is_contiguous()
    out = m(query, ref_pts, value, shapes)
    out.backward(grad_out)
    with torch.no_grad():
        raw_off = m.sampling_offsets(query).reshape(B, Lq, H, P, 2)
        raw_logit = m.attention_weights(query).reshape(B, Lq, H, P)
    return name, dict(
        memory_bf16=memory.detach().to(torch.bfloat16).view(torch.int16).numpy(),
        query=query.detach().numpy(), ref_points=ref_pts.numpy(),
        grad_out_bf16=grad_out.to(torch.bfloat16).view(torch.int16).numpy(),
        shapes=np.asarray(shapes, np.int32), npts=np.asarray(npts, np.int32), H=np.int32(H),
        so_w=m.sampling_offsets.weight.detach().numpy(), so_b=m.sampling_offsets.bias.detach().numpy(),
        aw_w=m.attention_weights.weight.detach().numpy(), aw_b=m.attention_weights.bias.detach().numpy(),
        num_points_scale=m.num_points_scale.numpy(), offset_scale=np.float32(m.offset_scale),
        raw_off=raw_off.numpy(), raw_logit=raw_logit.numpy(),
        out=out.detach().contiguous().numpy(), grad_memory=memory.grad.numpy(),
        grad_query=query.grad.numpy(),
        g_so_w=m.sampling_offsets.weight.grad.numpy(), g_so_b=m.sampling_offsets.bias.grad.numpy(),
        g_aw_w=m.attention_weights.weight.grad.numpy(), g_aw_b=m.attention_weights.bias.grad.numpy(),
        state_dict_keys=np.asarray(sorted(m.state_dict().keys())),
    )


def case_fdr(ref):
    g = torch.Generator().manual_seed(21)
    out = {}
    for tag, (up, rs) in {"m": (0.5, 4.0), "x": (0.5, 8.0), "odd": (-0.37, 5.5)}.items():
        upt, rst = torch.tensor([up]), torch.tensor([rs])
        out[f"project_{tag}"] = ref.weighting_function(32, upt, rst).numpy()
        out[f"project_deploy_{tag}"] = ref.weighting_function(32, upt, rst, deploy=True).numpy()
        out[f"up_rs_{tag}"] = np.asarray([up, rs], np.float32)
    reg_max, N = 32, 96
    corners = (torch.randn(2, N // 2, 4 * (reg_max + 1), generator=g) * 3.0)
    corners[0, 0] = 0.0            # uniform distribution
    corners[0, 1, :33] = 80.0      # overflow-prone logits
    corners[0, 2, 5] = 60.0        # one-hot
    corners = bf16r(corners).requires_grad_(True)
    ref_init = make_boxes(g, 2, N // 2)
    project = ref.weighting_function(reg_max, torch.tensor([0.5]), torch.tensor([4.0]))
    integral = ref.Integral(reg_max)
    dist = integral(corners, project)
    boxes = ref.distance2bbox(ref_init, dist, torch.tensor([4.0]))
    gb = torch.randn(boxes.shape, generator=g)
    gd = torch.randn(dist.shape, generator=g)
    (boxes * gb).sum().backward(retain_graph=True)
    g_from_boxes = corners.grad.clone()
    corners.grad = None
    ((boxes * gb).sum() + (dist * gd).sum()).backward()
    out.update(corners=corners.detach().numpy(), ref_init=ref_init.numpy(),
               project=project.numpy(), reg_scale=np.float32(4.0), dist=dist.detach().numpy(),
               boxes=boxes.detach().numpy(), grad_boxes=gb.numpy(), grad_dist=gd.numpy(),
               grad_corners_from_boxes=g_from_boxes.numpy(),
               grad_corners_from_both=corners.grad.numpy())
    return "fdr", out


def case_mask(ref):
    g = torch.Generator().manual_seed(31)
    B, Q, C, Hm, Wm = 2, 37, 64, 12, 20
    h = bf16r(torch.randn(B, Q, C, generator=g))
    feat = bf16r(torch.randn(B, C, Hm, Wm, generator=g))
    stub = types.SimpleNamespace(mask_head=torch.nn.Identity())
    logits = ref.DFINETransformer._mask_logits_from_h(stub, h, feat)
    return "mask", dict(coef=h.numpy(), proto=feat.numpy(), logits=logits.numpy(),
                        probs=torch.sigmoid(logits).numpy())


def case_mask_bwd(ref):
    """Gradients of the reference's mask contraction (autograd of dfine_decoder.py:940) for a
    bf16-representable upstream gradient: what the two backward tensor-core contractions must return."""
    g = torch.Generator().manual_seed(32)
    B, Q, C, Hm, Wm = 2, 37, 128, 12, 20
    h = bf16r(torch.randn(B, Q, C, generator=g)).requires_grad_(True)
    feat = bf16r(torch.randn(B, C, Hm, Wm, generator=g)).requires_grad_(True)
    go = bf16r(torch.randn(B, Q, Hm, Wm, generator=g))
    stub = types.SimpleNamespace(mask_head=torch.nn.Identity())
    logits = ref.DFINETransformer._mask_logits_from_h(stub, h, feat)
    logits.backward(go)
    return "mask_bwd", dict(coef=h.detach().numpy(), proto=feat.detach().numpy(), grad_out=go.numpy(),
                            logits=logits.detach().numpy(), grad_coef=h.grad.numpy(), grad_proto=feat.grad.numpy())


def case_mask_loss(ref):
    """DFINECriterion._focal_loss_mask / _dice_loss (dfine_criterion.py:273-312) on matched mask rows:
    binary targets with small / large / empty / full foreground (the adaptive alpha, :279-282), one row of
    soft target values, large-magnitude logits; losses and their gradients w.r.t. the logits."""
    g = torch.Generator().manual_seed(41)
    M, Hm, Wm = 7, 12, 20
    pred = (torch.randn(M, Hm, Wm, generator=g) * 3.0)
    pred[5] *= 12.0                                   # saturated sigmoid / large |x|
    tgt = torch.zeros(M, Hm, Wm)
    tgt[0, 2:5, 3:6] = 1                              # small foreground
    tgt[1, :, :15] = 1                                # large foreground
    tgt[2] = 0                                        # empty
    tgt[3] = 1                                        # full
    tgt[4] = (torch.rand(Hm, Wm, generator=g) < 0.5).float()
    tgt[5] = (torch.rand(Hm, Wm, generator=g) < 0.3).float()
    tgt[6] = torch.rand(Hm, Wm, generator=g)          # soft values (float masks)
    out = dict(pred=pred.numpy().copy(), tgt=tgt.numpy().copy())
    for name, fn in (("bce", ref.DFINECriterion._focal_loss_mask), ("dice", ref.DFINECriterion._dice_loss)):
        p = pred.clone().requires_grad_(True)
        loss = fn(p, tgt)
        loss.backward()
        out["loss_" + name] = np.asarray(loss.item(), np.float32)
        out["grad_" + name] = p.grad.numpy().copy()
    return "mask_loss", out


def case_layer(ref):
    """The decoder layer's Linears as the reference runs them under autocast(bfloat16) (the shipped trainer's
    AMP and the inference wrapper's half mode): MSDeformableAttention's two Linears on with_pos_embed(target,
    pos) (dfine_decoder.py:139-147, :245), Gate.forward (:258-271), and the FFN tail of
    TransformerDecoderLayer.forward (:251-253: forward_ffn, residual, clamp, norm3).  torch CPU autocast: the
    same rounding points as on CUDA (Linear operands and outputs in bf16, sigmoid on bf16, LayerNorm on fp32)."""
    g = torch.Generator().manual_seed(51)
    M, C, Fd, H, P = 150, 256, 1024, 8, 12
    layer = ref.TransformerDecoderLayer(d_model=C, n_head=H, dim_feedforward=Fd, n_levels=3, n_points=[3, 6, 3])
    layer.eval()
    with torch.no_grad():
        for p_ in layer.parameters():                       # trained-like values everywhere
            # (weight matrices bf16-representable: autocast rounds them anyway, and the fixture stores 2 bytes each)
            p_.copy_(bf16r(torch.randn(p_.shape, generator=g) * 0.05) if p_.dim() > 1
                     else torch.randn(p_.shape, generator=g) * 0.3)
        for ln in (layer.gateway.norm, layer.norm3):
            ln.weight.add_(1.0)
    target = torch.randn(1, M, C, generator=g) * 1.3
    pos = torch.randn(1, M, C, generator=g)
    x2 = torch.randn(1, M, C, generator=g) * 0.7 + 0.1
    target[0, 0, :4] = torch.tensor([7e4, -7e4, 65504.0, 3.0])     # reaches the clamp of :253
    ca = layer.cross_attn
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        q = layer.with_pos_embed(target, pos)
        raw = torch.cat([ca.sampling_offsets(q), ca.attention_weights(q)], -1)
        gate_out = layer.gateway(target, x2)
        hid = layer.activation(layer.linear1(target))
        t2 = layer.forward_ffn(target)
        ffn_out = layer.norm3((target + t2).clamp(min=-65504, max=65504))
    assert raw.dtype == torch.bfloat16 and hid.dtype == torch.bfloat16
    assert gate_out.dtype == torch.float32 and ffn_out.dtype == torch.float32
    f = lambda t: t.detach().float().numpy().copy()
    h = lambda t: t.detach().to(torch.bfloat16).view(torch.int16).numpy().copy()     # bf16 bit patterns
    return "layer", dict(
        target=f(target[0]), pos=f(pos[0]), x2=f(x2[0]),
        so_w_bf16=h(ca.sampling_offsets.weight), so_b=f(ca.sampling_offsets.bias),
        aw_w_bf16=h(ca.attention_weights.weight), aw_b=f(ca.attention_weights.bias), raw_bf16=h(raw[0]),
        gate_w_bf16=h(layer.gateway.gate.weight), gate_b=f(layer.gateway.gate.bias),
        gate_ln_w=f(layer.gateway.norm.weight), gate_ln_b=f(layer.gateway.norm.bias),
        gate_eps=np.float32(layer.gateway.norm.eps), gate_out=f(gate_out[0]),
        w1_bf16=h(layer.linear1.weight), b1=f(layer.linear1.bias), w2_bf16=h(layer.linear2.weight),
        b2=f(layer.linear2.bias), ln3_w=f(layer.norm3.weight), ln3_b=f(layer.norm3.bias),
        ln3_eps=np.float32(layer.norm3.eps), hidden_bf16=h(hid[0]), ffn_out=f(ffn_out[0]))


def case_lqe(ref):
    """LQE.forward (dfine_decoder.py:307-313) of the reference's LQE(4, 64, 2, 32): float32, and under CPU
    autocast(bfloat16) with bf16 scores / corners (what the decoder hands it under AMP; note that torch's CPU
    autocast keeps softmax / topk / mean in bf16 where CUDA autocast computes the softmax in float32).  Engineered
    rows: a uniform distribution (all ties), saturated logits, a one-hot."""
    g = torch.Generator().manual_seed(61)
    B, L, nc, reg_max = 2, 45, 7, 32
    lqe = ref.LQE(4, 64, 2, reg_max)
    with torch.no_grad():
        for p_ in lqe.parameters():
            p_.copy_(bf16r(torch.randn(p_.shape, generator=g) * 0.3))     # (the output layer is zero-initialised)
    pc = bf16r(torch.randn(B, L, 4 * (reg_max + 1), generator=g) * 3.0)
    pc[0, 0] = 0.0
    pc[0, 1, :33] = 80.0
    pc[0, 2, 5] = 60.0
    sc = bf16r(torch.randn(B, L, nc, generator=g))
    with torch.no_grad():
        out32 = lqe(sc, pc)
        with torch.autocast("cpu", dtype=torch.bfloat16):
            out16 = lqe(sc.bfloat16(), pc.bfloat16())
    assert out32.dtype == torch.float32 and out16.dtype == torch.bfloat16
    f = lambda t: t.detach().float().numpy().copy()
    l1, l2 = lqe.reg_conf.layers
    return "lqe", dict(scores=f(sc), corners=f(pc), w1=f(l1.weight), b1=f(l1.bias), w2=f(l2.weight), b2=f(l2.bias),
                       out_f32=f(out32), out_bf16=f(out16))


def _criterion_class():
    from src.d_fine.dfine_criterion import DFINECriterion  # noqa: E402
    return DFINECriterion


def load_reference(path: str):
    sys.path.insert(0, path)
    from src.d_fine.arch import dfine_decoder as dd  # noqa: E402
    from src.d_fine.arch import utils as au  # noqa: E402
    return types.SimpleNamespace(
        core=au.deformable_attention_core_func_v2, weighting_function=au.weighting_function,
        distance2bbox=au.distance2bbox, MSDeformableAttention=dd.MSDeformableAttention,
        Integral=dd.Integral, TransformerDecoder=dd.TransformerDecoder,
        TransformerDecoderLayer=dd.TransformerDecoderLayer, LQE=dd.LQE,
        DFINETransformer=dd.DFINETransformer, DFINECriterion=_criterion_class())


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    ap.add_argument("--only", default=None, help="comma-separated fixture names to (re)write; default all")
    a = ap.parse_args()
    torch.set_num_threads(1)
    torch.use_deterministic_algorithms(True)
    ref = load_reference(a.ref)
    os.makedirs(a.out, exist_ok=True)
    cases = [
        case_core(ref, "core_m_small", 1, 2, 40, 8, 32, [[12, 16], [6, 8], [3, 4]], [3, 6, 3]),
        case_core(ref, "core_n_small", 2, 1, 30, 8, 16, [[10, 10], [5, 5]], [6, 6]),
        case_core(ref, "core_x444", 3, 1, 20, 8, 32, [[16, 16], [8, 8], [4, 4]], [4, 4, 4]),
        case_edge(ref),
        case_near_centre(ref),
        case_probe(ref),
        case_module(ref, "module_m_small", 11, 2, 33, 256, 8, [[12, 16], [6, 8], [3, 4]], [3, 6, 3]),
        case_module(ref, "module_n_small", 12, 1, 21, 128, 8, [[10, 12], [5, 6]], [6, 6]),
        case_fdr(ref),
        case_mask(ref),
        case_mask_bwd(ref),
        case_mask_loss(ref),
        case_layer(ref),
        case_lqe(ref),
    ]
    for name, arrs in cases:
        if a.only and name not in a.only.split(","):
            continue
        p = os.path.join(a.out, name + ".npz")
        np.savez_compressed(p, **arrs)
        print(f"{name}: {os.path.getsize(p) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
