"""Synthetic batches and the restated train / inference step around the UNMODIFIED reference model.

Used by `tests/test_gpu_model.py` and `bench.py`'s `full_model` legs.  The model, criterion,
matcher and optimizer builder are the reference's own (`baseline/_ref/src/d_fine`, imported through
`baseline.ref_install`); only the trainer's inner step is restated here because `src/dl/train.py`
imports hydra / albumentations / torchmetrics, none of which exist in this image:

* `train_step`  -- reference src/dl/train.py:545-557 (autocast forward, criterion outside autocast,
  `sum(loss_dict.values())`, backward) + `optimizer_step` :488-511 (clip 0.1, step, zero_grad).
* `synthetic_batch` -- SURVEY.md section 8(d): images U(0,1) [B,3,S,S]; 10 boxes per image,
  labels randint(0, 80), centres U(0.2, 0.8), sizes U(0.05, 0.25); seg: uint8 rectangle masks.
* `LRS` -- reference config.yaml:109-125.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ref_install

LRS = {"n": (0.0004, 0.0008), "s": (0.00006, 0.00025), "m": (0.00002, 0.00015),
       "l": (0.00000625, 0.000125), "x": (0.0000015, 0.0001)}   # (backbone_lr, base_lr)


def synthetic_batch(batch: int, size: int, device, seed: int = 42, seg: bool = False, n_gt: int = 10,
                    num_classes: int = 80) -> Tuple[torch.Tensor, List[Dict[str, torch.Tensor]]]:
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(batch, 3, size, size, generator=g)
    targets = []
    for _ in range(batch):
        labels = torch.randint(0, num_classes, (n_gt,), generator=g)
        cxcy = torch.rand(n_gt, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(n_gt, 2, generator=g) * 0.2 + 0.05
        t = {"labels": labels, "boxes": torch.cat([cxcy, wh], -1)}
        if seg:
            m = torch.zeros(n_gt, size, size, dtype=torch.uint8)
            x0 = ((cxcy[:, 0] - wh[:, 0] / 2) * size).long().clamp(0, size - 1)
            x1 = ((cxcy[:, 0] + wh[:, 0] / 2) * size).long().clamp(1, size)
            y0 = ((cxcy[:, 1] - wh[:, 1] / 2) * size).long().clamp(0, size - 1)
            y1 = ((cxcy[:, 1] + wh[:, 1] / 2) * size).long().clamp(1, size)
            for i in range(n_gt):
                m[i, int(y0[i]):max(int(y1[i]), int(y0[i]) + 1), int(x0[i]):max(int(x1[i]), int(x0[i]) + 1)] = 1
            t["masks"] = m
        targets.append(t)
    images = images.to(device)
    targets = [{k: v.to(device) for k, v in t.items()} for t in targets]
    return images, targets


def build(model_name: str, device, size: int = 640, seg: bool = False, num_classes: int = 80, seed: int = 0,
          trained: bool = True):
    """(model, criterion) built by the reference's own factories (dfine.py:51-84), seeded; `trained` gives the
    decoder's zero-initialised tensors small values (see trained_like)."""
    dfine = ref_install.import_reference()
    from copy import deepcopy

    torch.manual_seed(seed)
    model = dfine.build_model(model_name, num_classes, seg, str(device), img_size=[size, size])
    # build_loss appends "masks" to the shared config dict on every call (dfine.py:77-78): work on a copy
    saved = deepcopy(dfine.models[model_name])
    try:
        loss_fn = dfine.build_loss(model_name, num_classes, 0.0, seg)
    finally:
        dfine.models[model_name] = saved
    if trained:
        trained_like(model, seed + 1)
    return model, loss_fn.to(device)


@torch.no_grad()
def trained_like(model, seed: int = 1) -> int:
    """A freshly built D-FINE has many zero-initialised tensors on the hot path (sampling_offsets.weight,
    attention_weights.*, the last layers of the bbox / LQE heads, the Gate weight: dfine_decoder.py:100-117,
    :264-266, :302-303, :826-850), so attention is uniform, every FDR distribution is flat and the decoded
    boxes equal the reference points.  Give every all-zero parameter small seeded values (in place, same on
    every arm because it runs before the deepcopy) so that the parity tests exercise softmax, offsets and the
    box refinement with non-trivial numbers.  Returns the number of tensors touched."""
    g = torch.Generator().manual_seed(seed)
    n = 0
    for name, p in model.named_parameters():
        if name.startswith("decoder.") and p.numel() > 1 and float(p.abs().max()) == 0.0:
            std = 0.01 if "sampling_offsets" in name else 0.05
            p.copy_((torch.randn(p.shape, generator=g) * std).to(p.device, p.dtype))
            n += 1
    return n


def build_optimizer(model, model_name: str):
    dfine = ref_install.import_reference()
    backbone_lr, base_lr = LRS[model_name]
    return dfine.build_optimizer(model, lr=base_lr, backbone_lr=backbone_lr, betas=(0.9, 0.999),
                                 weight_decay=0.000125, base_lr=base_lr)


def forward_loss(model, loss_fn, images, targets, amp_dtype: Optional[torch.dtype] = None):
    """Reference src/dl/train.py:545-556 up to the loss (no backward)."""
    dev = images.device.type
    if amp_dtype is not None:
        with torch.autocast(dev, dtype=amp_dtype, cache_enabled=True):
            output = model(images, targets=targets)
        with torch.autocast(dev, enabled=False):
            loss_dict = loss_fn(output, targets)
    else:
        output = model(images, targets=targets)
        loss_dict = loss_fn(output, targets)
    return output, loss_dict, sum(loss_dict.values())


def train_step(model, loss_fn, images, targets, amp_dtype: Optional[torch.dtype] = None, optimizer=None,
               scaler=None, clip_max_norm: float = 0.1):
    """One optimisation step as the reference trainer runs it (train.py:545-557, :488-511)."""
    output, loss_dict, loss = forward_loss(model, loss_fn, images, targets, amp_dtype)
    if scaler is not None:
        scaler.scale(loss).backward()
    else:
        loss.backward()
    if optimizer is not None:
        if scaler is not None:
            if clip_max_norm:
                scaler.unscale_(optimizer)
                torch.nn.utils.clip_grad_norm_(model.parameters(), clip_max_norm)
            scaler.step(optimizer)
            scaler.update()
        else:
            if clip_max_norm:
                torch.nn.utils.clip_grad_norm_(model.parameters(), clip_max_norm)
            optimizer.step()
        optimizer.zero_grad()
    return output, loss_dict, loss


@torch.no_grad()
def infer_step(model, images, amp_dtype: Optional[torch.dtype] = None):
    """Reference src/infer/torch_model.py:303-344 restated: eval forward under no_grad (+ autocast)."""
    if amp_dtype is not None:
        with torch.autocast(images.device.type, dtype=amp_dtype):
            return model(images)
    return model(images)


def flat_outputs(output: dict) -> Dict[str, torch.Tensor]:
    """Every tensor of the decoder's output dict (main, aux, pre, enc_aux, dn heads) under a flat name."""
    flat = {}

    def visit(prefix, o):
        if isinstance(o, torch.Tensor):
            flat[prefix] = o
        elif isinstance(o, dict):
            for k, v in o.items():
                visit(f"{prefix}.{k}" if prefix else k, v)
        elif isinstance(o, (list, tuple)):
            for i, v in enumerate(o):
                visit(f"{prefix}[{i}]", v)

    visit("", output)
    return flat
