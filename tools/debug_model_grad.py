"""Debug: captured inputs of every cross_attn core call of the reference model -> reference core vs dfine_b200 core."""
import copy, os, sys, functools
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-fine-seg_b200"))
from baseline import model_harness as H, ref_install
import dfine_b200
from dfine_b200 import ops

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "n"
model, loss_fn = H.build(name, dev, 640, False)
model.train(); loss_fn.train()
images, targets = H.synthetic_batch(4, 640, dev, seed=42)
from src.d_fine.arch.utils import deformable_attention_core_func_v2 as ref_core

cap = []
for i, layer in enumerate(model.decoder.decoder.layers):
    m = layer.cross_attn
    def hook(value, shapes, loc, aw, npl, _i=i, _m=m):
        cap.append((_i, [v.detach().clone() for v in value], shapes, loc.detach().clone(), aw.detach().clone(), npl))
        return ref_core(value, shapes, loc, aw, npl, method="default")
    m.ms_deformable_attn_core = hook
torch.manual_seed(1234)
out, ld, loss = H.forward_loss(model, loss_fn, images, targets, None)
loss.backward()
for (i, value, shapes, loc, aw, npl) in cap:
    print("layer", i, "loc range", float(loc.min()), float(loc.max()), "nan", bool(torch.isnan(loc).any()), loc.shape)
    B, Lq = loc.shape[:2]
    spec = ops.level_spec(shapes, npl)
    # rebuild memory [B, L, C] and views like value_op
    Hh, c = value[0].shape[1], value[0].shape[2]
    mem = torch.cat([v.permute(0, 3, 1, 2) for v in value], 1).reshape(B, spec.L, Hh * c).contiguous()
    go = torch.randn(B, Lq, Hh * c, device=dev)
    res = []
    for which in ("ref", "b200"):
        m_ = mem.clone().requires_grad_(True)
        views = m_.reshape(B, spec.L, Hh, c).permute(0, 2, 3, 1).split(spec.sizes, dim=-1)
        l_ = loc.clone().requires_grad_(True); a_ = aw.clone().requires_grad_(True)
        fn = ref_core if which == "ref" else ops.msda_core
        o = fn(views, shapes, l_, a_, npl, method="default")
        o.backward(go)
        res.append((o.detach(), m_.grad, l_.grad, a_.grad))
    for nm, a, b in zip(("out", "g_value", "g_loc", "g_attn"), res[1], res[0]):
        d = (a - b).abs()
        k = int(d.argmax())
        print(f"   {nm}: max|d| {float(d.max()):.3e} scale {float(b.abs().max()):.3e} at {k} got {float(a.reshape(-1)[k]):.6e} want {float(b.reshape(-1)[k]):.6e}")
        if nm == "g_loc":
            idx = torch.unravel_index(torch.tensor(k), d.shape)
            idx = [int(t) for t in idx]
            print("      index", idx, "loc", loc[idx[0], idx[1], idx[2], idx[3]].tolist())
            bad = (d > 1e-4 * b.abs().max()).nonzero()
            print("      #bad", bad.shape[0])
            for r in bad[:10].tolist():
                lx, ly = loc[r[0], r[1], r[2], r[3]].tolist()
                lvl = 0
                pp = r[3]
                for li, n in enumerate(npl):
                    if pp < n: lvl = li; break
                    pp -= n
                hh, ww = shapes[lvl]
                print("       ", r, "loc", lx, ly, "ix", lx * ww - 0.5, "iy", ly * hh - 0.5, "got", float(a[tuple(r)]), "want", float(b[tuple(r)]))
