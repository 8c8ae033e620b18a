import sys, torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/d-fine-seg_b200")
from dfine_b200 import ops
dev = torch.device("cuda:0")
for B in (3, 16, 64):
    L, nc = 300, 80
    torch.manual_seed(B + L + nc)
    l1, l2 = nn.Linear(20, 64).to(dev), nn.Linear(64, 1).to(dev)
    with torch.no_grad():
        l2.weight.normal_(0, 0.3); l2.bias.normal_(0, 0.3)
    pc = (torch.randn(B, L, 132, device=dev) * 3.0).bfloat16()
    sc = torch.randn(B, L, nc, device=dev).bfloat16()
    r = lambda t: t.bfloat16().float()
    with torch.no_grad():
        prob = F.softmax(pc.float().reshape(B, L, 4, 33), dim=-1)
        tk, _ = prob.topk(4, dim=-1)
        stat = torch.cat([tk, tk.mean(-1, keepdim=True)], -1).reshape(B, L, 20)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            h_ref = F.relu(l1(stat)); q_ref = l2(h_ref)
        h_em = F.relu(r(r(stat) @ r(l1.weight).t() + r(l1.bias)))
        q_em = r(h_em @ r(l2.weight).t() + r(l2.bias))
        zero = torch.zeros(B, L, nc, device=dev, dtype=torch.bfloat16)
        q_got = ops.lqe_fwd(zero, pc, l1.weight.detach(), l1.bias.detach(), l2.weight.detach(), l2.bias.detach(), emulate_bf16=True)[..., :1].float()
    print(B, "h ref vs em", float((h_ref.float() - h_em).abs().max()), "q ref vs em", float((q_ref.float() - q_em).abs().max()),
          "q got vs em", float((q_got - q_em).abs().max()), "q got vs ref", float((q_got - q_ref.float()).abs().max()),
          "frac q got!=ref", float((q_got != q_ref.float()).float().mean()))
