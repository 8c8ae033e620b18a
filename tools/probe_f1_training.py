"""Feasibility probe (CPU or GPU): does a briefly trained D-FINE-n-seg on synthetic coloured rectangles reach a
non-trivial F1 / mask IoU?  (for the north star's "F1 / mask IoU unchanged on a fixed synthetic set")"""
import sys, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/d-fine-seg_b200")
from baseline import model_harness as H
from oracle import det_metrics as DM

COLORS = torch.tensor([[1.0, 0.1, 0.1], [0.1, 1.0, 0.1], [0.1, 0.1, 1.0], [1.0, 1.0, 0.1]])


def rect_batch(batch, size, device, seed, n_gt=3, nc=4):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(batch, 3, size, size, generator=g) * 0.15
    targets = []
    for b in range(batch):
        labels = torch.randint(0, nc, (n_gt,), generator=g)
        cxcy = torch.rand(n_gt, 2, generator=g) * 0.6 + 0.2
        wh = torch.rand(n_gt, 2, generator=g) * 0.2 + 0.12
        m = torch.zeros(n_gt, size, size, dtype=torch.uint8)
        for i in range(n_gt):
            x0, x1 = int((cxcy[i, 0] - wh[i, 0] / 2) * size), int((cxcy[i, 0] + wh[i, 0] / 2) * size)
            y0, y1 = int((cxcy[i, 1] - wh[i, 1] / 2) * size), int((cxcy[i, 1] + wh[i, 1] / 2) * size)
            m[i, y0:y1, x0:x1] = 1
            images[b, :, y0:y1, x0:x1] = COLORS[labels[i]].view(3, 1, 1) * (0.8 + 0.2 * torch.rand(1, generator=g))
        targets.append({"labels": labels.to(device), "boxes": torch.cat([cxcy, wh], -1).to(device), "masks": m.to(device)})
    return images.to(device), targets


def evaluate(model, size, dev, seeds, amp=None):
    model.eval()
    tp = fp = fn = 0
    for s in seeds:
        images, targets = rect_batch(4, size, dev, s)
        out = H.infer_step(model, images, amp)
        preds = DM.postprocess(out["pred_logits"], out["pred_boxes"], 0.5)
        gts = [{"boxes": DM.box_cxcywh_to_xyxy(t["boxes"].cpu()), "labels": t["labels"].cpu()} for t in targets]
        a, b, c, _, _ = DM.f1_counts(preds, gts)
        tp += a; fp += b; fn += c
    model.train()
    return DM.f1_score(tp, fp, fn), tp, fp, fn


if __name__ == "__main__":
    dev = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
    size, steps = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 200
    model, loss_fn = H.build("n", dev, size, True, num_classes=4, trained=False)
    model.train(); loss_fn.train()
    opt = H.build_optimizer(model, "n")
    t0 = time.time()
    for it in range(steps):
        images, targets = rect_batch(8, size, dev, 1000 + it)
        _, _, loss = H.train_step(model, loss_fn, images, targets, None, optimizer=opt)
        if it % 100 == 99:
            print(it + 1, "loss %.3f" % float(loss), "f1 tp fp fn", evaluate(model, size, dev, range(5)), "%.0fs" % (time.time() - t0), flush=True)
