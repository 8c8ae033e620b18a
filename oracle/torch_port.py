"""Torch restatement of the reference hot path -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference implements this path as a sequence of PyTorch ATen calls (F.grid_sample,
F.softmax, F.linear, torch.einsum); the reference tree itself does not travel to the GPU
box, so this file restates those call sequences with the same ATen operators.  It is what
bench.py times as the CPU baseline (`cpu_baseline`, `--impl reference`; kind "port") and,
on the GPU, as the eager-PyTorch bar that the CUDA kernels are compared with.  It is pinned
to the real reference by tests/test_oracle_golden.py (same golden vectors as the C oracle).

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F


def value_views(memory: torch.Tensor, num_heads: int, spatial_shapes) -> Sequence[torch.Tensor]:
    """TransformerDecoder.value_op with no mask / projection (dfine_decoder.py:416-426)."""
    B, L, C = memory.shape
    v = memory.reshape(B, L, num_heads, C // num_heads).permute(0, 2, 3, 1)
    return v.split([h * w for h, w in spatial_shapes], dim=-1)


def msda_core(value, spatial_shapes, sampling_locations, attention_weights,
              num_points_list: List[int]) -> torch.Tensor:
    """deformable_attention_core_func_v2, method='default' (arch/utils.py:191-264)."""
    B, H, c, _ = value[0].shape
    Lq = sampling_locations.shape[1]
    grids = (2 * sampling_locations - 1).permute(0, 2, 1, 3, 4).flatten(0, 1)
    per_level = grids.split(num_points_list, dim=-2)
    sampled = []
    for lvl, (h, w) in enumerate(spatial_shapes):
        fmap = value[lvl].reshape(B * H, c, h, w)
        sampled.append(F.grid_sample(fmap, per_level[lvl], mode="bilinear",
                                     padding_mode="zeros", align_corners=False))
    aw = attention_weights.permute(0, 2, 1, 3).reshape(B * H, 1, Lq, sum(num_points_list))
    out = (torch.cat(sampled, dim=-1) * aw).sum(-1).reshape(B, H * c, Lq)
    return out.permute(0, 2, 1)


def msda_module(query, reference_points, value, spatial_shapes, so_w, so_b, aw_w, aw_b,
                num_points_scale, num_points_list, num_heads, offset_scale=0.5):
    """MSDeformableAttention.forward, reference_points last-dim 4 (dfine_decoder.py:119-178)."""
    B, Lq = query.shape[:2]
    P = sum(num_points_list)
    so = F.linear(query, so_w, so_b).reshape(B, Lq, num_heads, P, 2)
    aw = F.softmax(F.linear(query, aw_w, aw_b).reshape(B, Lq, num_heads, P), dim=-1)
    nps = num_points_scale.to(dtype=query.dtype).unsqueeze(-1)
    offset = so * nps * reference_points[:, :, None, :, 2:] * offset_scale
    loc = reference_points[:, :, None, :, :2] + offset
    return msda_core(value, spatial_shapes, loc, aw, num_points_list)


def msda_from_raw(raw_off, raw_logit, reference_points, value, spatial_shapes,
                  num_points_scale, num_points_list, offset_scale=0.5):
    """Same as msda_module after the two Linears (dfine_decoder.py:144-176)."""
    aw = F.softmax(raw_logit, dim=-1)
    nps = num_points_scale.to(dtype=torch.float32).unsqueeze(-1)
    offset = raw_off * nps * reference_points[:, :, None, :, 2:] * offset_scale
    loc = reference_points[:, :, None, :, :2] + offset
    return msda_core(value, spatial_shapes, loc, aw, num_points_list)


def weighting_function(reg_max: int, up: torch.Tensor, reg_scale: torch.Tensor) -> torch.Tensor:
    """arch/utils.py:145-188, non-deploy branch."""
    ub1 = abs(up[0]) * abs(reg_scale)
    ub2 = ub1 * 2
    step = (ub1 + 1) ** (2 / (reg_max - 2))
    left = [-(step ** i) + 1 for i in range(reg_max // 2 - 1, 0, -1)]
    right = [step ** i - 1 for i in range(1, reg_max // 2)]
    vals = [-ub2] + left + [torch.zeros_like(up[0][None])] + right + [ub2]
    return torch.cat([v.reshape(1) for v in vals], 0)


def integral(x: torch.Tensor, project: torch.Tensor, reg_max: int = 32) -> torch.Tensor:
    """Integral.forward (dfine_decoder.py:291-295)."""
    shape = x.shape
    p = F.softmax(x.reshape(-1, reg_max + 1), dim=1)
    d = F.linear(p, project.to(p.device)).reshape(-1, 4)
    return d.reshape(list(shape[:-1]) + [-1])


def distance2bbox(points, distance, reg_scale):
    """arch/utils.py:119-142 followed by box_xyxy_to_cxcywh (:70-73)."""
    rs = abs(reg_scale)
    x1 = points[..., 0] - (0.5 * rs + distance[..., 0]) * (points[..., 2] / rs)
    y1 = points[..., 1] - (0.5 * rs + distance[..., 1]) * (points[..., 3] / rs)
    x2 = points[..., 0] + (0.5 * rs + distance[..., 2]) * (points[..., 2] / rs)
    y2 = points[..., 1] + (0.5 * rs + distance[..., 3]) * (points[..., 3] / rs)
    return torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], -1)


def mask_logits(coef, mask_feat):
    """einsum of DFINETransformer._mask_logits_from_h (dfine_decoder.py:940)."""
    return torch.einsum("bqc,bchw->bqhw", coef, mask_feat)


def hot_path_step(mem, queries, refs, lins, spatial_shapes, npts, H, corners, ref_init, up,
                  reg_scale, grad_outs, grad_boxes, train: bool = True):
    """One pass of the decoder hot path for one batch: for every decoder layer the
    MSDeformableAttention forward (+ backward) and the FDR decode (+ backward).
    Used by bench.py for the eager-PyTorch baselines (CPU and GPU)."""
    outs = []
    project = weighting_function(32, up, reg_scale)
    for i in range(len(queries)):
        so_w, so_b, aw_w, aw_b, nps = lins[i]
        m = mem.detach().requires_grad_(train)
        q = queries[i].detach().requires_grad_(train)
        value = value_views(m, H, spatial_shapes)
        out = msda_module(q, refs[i], value, spatial_shapes, so_w, so_b, aw_w, aw_b, nps, npts, H)
        pc = corners[i].detach().requires_grad_(train)
        boxes = distance2bbox(ref_init, integral(pc, project), reg_scale)
        if train:
            torch.autograd.backward([out, boxes], [grad_outs[i], grad_boxes[i]])
            outs.append((out.detach(), boxes.detach(), m.grad, q.grad, pc.grad))
        else:
            outs.append((out, boxes))
    return outs
