import sys, copy, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/d-fine-seg_b200")
from baseline import model_harness as H
import dfine_b200
dev = torch.device("cuda:0")
model, loss_fn = H.build("s", dev, 640, False)
H.trained_like(model)
model.train()
a, b, c = copy.deepcopy(model), copy.deepcopy(model), copy.deepcopy(model)
dfine_b200.patch_model(a); dfine_b200.patch_model(b, layer=True); dfine_b200.patch_model(c)
images, targets = H.synthetic_batch(2, 640, dev, seed=5)
res = []
for m in (a, b, c):
    torch.manual_seed(3)
    m.zero_grad(set_to_none=True)
    _, _, loss = H.forward_loss(m, loss_fn, images, targets, torch.bfloat16)
    loss.backward()
    res.append((float(loss), {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}))
print([r[0] for r in res])
def err(x, y): return float((x.double()-y.double()).abs().max()) / max(float(y.double().abs().max()), 1e-30)
rows = sorted(((err(res[1][1][k], v), err(res[2][1][k], v), float(v.abs().max()), k) for k, v in res[0][1].items()), reverse=True)
for r in rows[:12]: print("%.3e noise %.3e max %.3e %s" % r)
