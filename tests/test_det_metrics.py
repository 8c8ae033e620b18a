"""Known-answer tests of the metric restatement (oracle/det_metrics.py): the reference validator's
own cases 1-4 (src/dl/validator.py:706-780: perfect match, partial match IoU 0.75,
misclassification, pure false positive), with boxes for the box metric and masks for mask IoU."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import det_metrics as DM  # noqa: E402


def _sample(boxes, labels):
    return {"boxes": torch.tensor(boxes, dtype=torch.float32).reshape(-1, 4),
            "labels": torch.tensor(labels, dtype=torch.int64)}


def _prf(preds, gts, thr=0.5):
    tps, fps, fns, ious, _ = DM.f1_counts(preds, gts, thr)
    p = tps / (tps + fps) if tps + fps else 0.0
    r = tps / (tps + fns) if tps + fns else 0.0
    return p, r, DM.f1_score(tps, fps, fns), ious


def test_case1_perfect_match():
    p, r, f1, ious = _prf([_sample([[1, 1, 3, 3]], [0])], [_sample([[1, 1, 3, 3]], [0])])
    assert (p, r, f1) == (1.0, 1.0, 1.0) and np.allclose(ious, [1.0])
    m = np.zeros((4, 4), np.uint8)
    m[1:3, 1:3] = 1
    assert DM.mask_iou(m, m) == 1.0


def test_case2_partial_match_iou_075():
    p, r, f1, ious = _prf([_sample([[0, 0, 4, 3]], [0])], [_sample([[0, 0, 4, 4]], [0])])
    assert (p, r) == (1.0, 1.0) and np.allclose(ious, [0.75])
    gt = np.ones((4, 4), np.uint8)
    pr = gt.copy()
    pr[3] = 0
    assert DM.mask_iou(pr, gt) == 0.75


def test_case3_misclassification_counts_fp_and_fn():
    tps, fps, fns, _, matches = DM.f1_counts([_sample([[1, 1, 3, 3]], [1])], [_sample([[1, 1, 3, 3]], [0])])
    assert (tps, fps, fns) == (0, 1, 1) and matches == [[(0, 0)]]
    assert DM.f1_score(tps, fps, fns) == 0.0


def test_case4_pure_false_positive_and_empty_sets():
    assert DM.f1_counts([_sample([[1, 1, 3, 3]], [0])], [_sample([], [])])[:3] == (0, 1, 0)
    assert DM.f1_counts([_sample([], [])], [_sample([[1, 1, 3, 3]], [0])])[:3] == (0, 0, 1)
    assert DM.f1_counts([_sample([], [])], [_sample([], [])])[:3] == (0, 0, 0)
    assert DM.mask_iou(np.zeros((2, 2)), np.zeros((2, 2))) == 0.0


def test_greedy_matching_prefers_the_highest_iou():
    # two predictions overlap one ground truth: the better one matches, the other is a false positive
    preds = [_sample([[0, 0, 4, 4], [0, 0, 4, 3]], [0, 0])]
    tps, fps, fns, ious, matches = DM.f1_counts(preds, [_sample([[0, 0, 4, 4]], [0])])
    assert (tps, fps, fns) == (1, 1, 0) and matches == [[(0, 0)]] and np.allclose(ious, [1.0])


def test_postprocess_topk_over_queries_and_classes():
    logits = torch.full((1, 3, 4), -10.0)
    logits[0, 2, 1] = 4.0      # query 2, class 1
    logits[0, 0, 3] = 0.5      # query 0, class 3
    logits[0, 1, 0] = -0.5     # below the confidence threshold
    boxes = torch.tensor([[[0.5, 0.5, 0.2, 0.2], [0.3, 0.3, 0.1, 0.1], [0.7, 0.7, 0.4, 0.2]]])
    out = DM.postprocess(logits, boxes, conf_thresh=0.5, num_top_queries=300)[0]
    assert out["labels"].tolist() == [1, 3] and out["queries"].tolist() == [2, 0]
    assert torch.allclose(out["boxes"][0], torch.tensor([0.5, 0.6, 0.9, 0.8]))
    assert out["all_scores"].shape == (12,)
