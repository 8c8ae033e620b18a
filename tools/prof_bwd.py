"""One decoder layer of config 3: fused forward (writes the records) + backward, a few times (ncu target)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-fine-seg_b200"))
from dfine_b200 import ops
dev = torch.device("cuda:0")
B, Lq, H, c = 32, 500, 8, 32
shapes, npts = [[80, 80], [40, 40], [20, 20]], [3, 6, 3]
spec = ops.level_spec(shapes, npts)
P = spec.P
g = torch.Generator(device=dev).manual_seed(7)
mem = torch.randn(B, spec.L, H * c, device=dev, generator=g).to(torch.bfloat16)
ref = torch.cat([torch.rand(B, Lq, 2, device=dev, generator=g) * 0.9 + 0.05,
                 torch.exp(torch.rand(B, Lq, 2, device=dev, generator=g) * 3.4 - 3.9)], -1)
raw = torch.randn(B, Lq, 3 * H * P, device=dev, generator=g).to(torch.bfloat16)
attn_view = raw.reshape(-1)[2 * H * P:]
rs = raw.shape[-1]
nps = torch.tensor([1.0 / n for n in npts for _ in range(n)], device=dev)
rec = ops.new_records(mem, spec, H, Lq)
go = torch.randn(B, Lq, H * c, device=dev, generator=g)
g_raw = torch.empty_like(raw)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(reps):
    ops.msda_forward_raw(mem, spec, H, raw, attn_view, ref, nps, 0.5, True, torch.float32, samp_rs=rs, attn_rs=rs, records=rec)
    ops.msda_backward_raw(mem, spec, H, raw, attn_view, ref, nps, 0.5, True, go, gv_dtype=mem.dtype, samp_rs=rs,
                          attn_rs=rs, grad_raw=g_raw, records=rec)
torch.cuda.synchronize()
print("ok")
