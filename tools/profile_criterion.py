"""Finer split of the reference criterion's time (synchronised wall clock)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-fine-seg_b200"))
from loguru import logger
logger.remove()
from baseline import model_harness as H
dev = torch.device("cuda:0")
model, loss_fn = H.build("m", dev, 640, False)
model.train(); loss_fn.train()
images, targets = H.synthetic_batch(32, 640, dev, seed=42)
with torch.autocast("cuda", dtype=torch.bfloat16):
    out = model(images, targets=targets)
acc = {}
def wrap(obj, name, key=None):
    orig = getattr(obj, name)
    key = key or name
    def f(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = orig(*a, **k)
        torch.cuda.synchronize(); acc[key] = acc.get(key, 0.0) + time.perf_counter() - t
        acc[key + "#"] = acc.get(key + "#", 0) + 1
        return r
    setattr(obj, name, f)
wrap(loss_fn.matcher, "forward", "matcher")
for n in ("loss_boxes", "loss_labels_vfl", "loss_local", "_get_go_indices", "get_loss_meta_info"):
    wrap(loss_fn, n)
import scipy.optimize, src.d_fine.matcher as M
orig_lsa = M.linear_sum_assignment
def lsa(c):
    t = time.perf_counter(); r = orig_lsa(c); acc["scipy"] = acc.get("scipy", 0.0) + time.perf_counter() - t; acc["scipy#"] = acc.get("scipy#", 0) + 1; return r
M.linear_sum_assignment = lsa
for it in range(3):
    acc.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.autocast("cuda", enabled=False):
        ld = loss_fn(out, targets)
    torch.cuda.synchronize(); tot = time.perf_counter() - t0
    print("criterion total ms", round(tot * 1e3, 1), {k: (round(v * 1e3, 2) if not k.endswith("#") else v) for k, v in acc.items()}, flush=True)
print(len(ld), "loss terms")
