"""Restatement of the reference's detection metric (TEST INFRASTRUCTURE -- only tests/ import it).

The north star asks that "detection F1 / mask IoU must be unchanged on a fixed synthetic set".
The reference computes them in src/dl/validator.py, which cannot be imported on the GPU box
(torchmetrics / faster_coco_eval are not installed, /root/reference does not travel), so the
pieces on that path are restated here in numpy/torch:

* postprocess          src/dl/train.py:227-319  (sigmoid -> top-K over Q*C -> label = idx % C,
                        query = idx // C -> confidence threshold)
* box matching / F1    src/dl/validator.py:340-437  (pairs with IoU >= thresh, matched greedily in
                        descending IoU, class-aware TP / FP / FN; unmatched predictions are FPs,
                        unmatched ground truths FNs), F1 from the summed counts (:300-338)
* mask IoU             src/dl/validator.py:269-279  (intersection / union of binary masks)
* mask postprocess     src/dl/train.py:296-316 + src/dl/utils.py:715-786 (kept queries' mask probabilities -> half
                        -> bilinear resize to the input size -> clamp, >= conf_thresh -> zero outside the box)
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch


def box_cxcywh_to_xyxy(b: torch.Tensor) -> torch.Tensor:
    cx, cy, w, h = b.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def box_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """torchvision.ops.box_iou on xyxy boxes."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter)


def postprocess(logits: torch.Tensor, boxes_cxcywh: torch.Tensor, conf_thresh: float = 0.5,
                num_top_queries: int = 300) -> List[Dict[str, torch.Tensor]]:
    """train.py:227-319, focal-loss branch, boxes kept in normalised xyxy."""
    B, Q, C = logits.shape
    scores_all = torch.sigmoid(logits.float())
    flat = scores_all.flatten(1)
    K = min(num_top_queries, flat.shape[1])
    topk_scores, topk_idx = torch.topk(flat, K, dim=-1)
    topk_labels = topk_idx - (topk_idx // C) * C
    topk_q = topk_idx // C
    xyxy = box_cxcywh_to_xyxy(boxes_cxcywh.float())
    out = []
    for b in range(B):
        keep = topk_scores[b] >= conf_thresh
        q = topk_q[b][keep]
        out.append({"labels": topk_labels[b][keep].cpu(), "scores": topk_scores[b][keep].cpu(),
                    "boxes": xyxy[b][q].cpu(), "queries": q.cpu(),
                    "all_scores": topk_scores[b].cpu(), "all_labels": topk_labels[b].cpu(),
                    "all_queries": topk_q[b].cpu()})
    return out


def f1_counts(preds: List[Dict[str, torch.Tensor]], gts: List[Dict[str, torch.Tensor]], iou_thresh: float = 0.5):
    """validator.py:340-437 -> (TPs, FPs, FNs, matched IoUs, matches per image as (pred, gt) pairs)."""
    tps = fps = fns = 0
    ious_tp, matches = [], []
    for pred, gt in zip(preds, gts):
        pb, pl, gb, gl = pred["boxes"], pred["labels"], gt["boxes"], gt["labels"]
        mp, mg, pairs = set(), set(), []
        if len(pb) and len(gb):
            iou = box_iou(pb, gb)
            pi, gi = torch.nonzero(iou >= iou_thresh, as_tuple=True)
            vals = iou[pi, gi]
            order = torch.argsort(-vals)
            for k in order.tolist():
                p_, g_ = int(pi[k]), int(gi[k])
                if p_ in mp or g_ in mg:
                    continue
                mp.add(p_)
                mg.add(g_)
                pairs.append((p_, g_))
                if int(pl[p_]) == int(gl[g_]):
                    tps += 1
                    ious_tp.append(float(vals[k]))
                else:
                    fns += 1
                    fps += 1
        fps += len(pb) - len(mp)
        fns += len(gb) - len(mg)
        matches.append(pairs)
    return tps, fps, fns, ious_tp, matches


def f1_score(tps: int, fps: int, fns: int) -> float:
    precision = tps / (tps + fps) if tps + fps else 0.0
    recall = tps / (tps + fns) if tps + fns else 0.0
    return 2 * precision * recall / (precision + recall) if precision + recall else 0.0


def mask_iou(a: np.ndarray, b: np.ndarray) -> float:
    """validator.py:269-279 on binary masks."""
    a, b = a.astype(bool), b.astype(bool)
    union = np.logical_or(a, b).sum()
    return float(np.logical_and(a, b).sum() / union) if union else 0.0


def postprocess_masks(pred_masks: torch.Tensor, preds: List[Dict[str, torch.Tensor]], size: int,
                      conf_thresh: float = 0.5) -> List[torch.Tensor]:
    """train.py:296-316 for square inputs without letterboxing (orig size = network input size): per image the
    uint8 masks [N, size, size] of the kept predictions (`preds` from `postprocess`: "queries", "boxes" in
    normalised xyxy)."""
    out = []
    for b, p in enumerate(preds):
        q = p["queries"].to(pred_masks.device)
        if q.numel() == 0:
            out.append(torch.zeros((0, size, size), dtype=torch.uint8))
            continue
        mb = pred_masks[b, q].to(torch.float16).unsqueeze(0)
        m = torch.nn.functional.interpolate(mb, size=(size, size), mode="bilinear", align_corners=False)[0]
        m = (m.clamp(0, 1) >= conf_thresh).to(torch.uint8).cpu()
        ys = torch.arange(size)[None, :, None]
        xs = torch.arange(size)[None, None, :]
        x1, y1, x2, y2 = (p["boxes"] * size).T          # utils.py:772-786 cleanup_masks
        inside = (xs >= x1[:, None, None]) & (xs < x2[:, None, None]) & (ys >= y1[:, None, None]) & (ys < y2[:, None, None])
        out.append(m * inside.to(m.dtype))
    return out
