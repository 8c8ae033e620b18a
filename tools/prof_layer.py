"""One launch of each fused decoder-layer kernel at the config-3 shape (for ncu)."""
import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/d-fine-seg_b200")
from dfine_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(11)
M, C, Fd, N = 16000, 256, 1024, 288
x = torch.randn(M, C, device=dev, generator=g); pos = torch.randn(M, C, device=dev, generator=g)
x2 = torch.randn(M, C, device=dev, generator=g)
wb = (torch.randn(N, C, device=dev, generator=g) * 0.05).bfloat16(); bb = torch.randn(N, device=dev, generator=g).bfloat16()
wg = (torch.randn(2 * C, 2 * C, device=dev, generator=g) * 0.05).bfloat16(); bg = torch.randn(2 * C, device=dev, generator=g).bfloat16()
w1 = (torch.randn(Fd, C, device=dev, generator=g) * 0.05).bfloat16(); b1 = torch.randn(Fd, device=dev, generator=g).bfloat16()
w2 = (torch.randn(C, Fd, device=dev, generator=g) * 0.05).bfloat16(); b2 = torch.randn(C, device=dev, generator=g).bfloat16()
lnw, lnb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ops.linear_fwd(x, wb, bb, x_add=pos, save_input=True)
    ops.gate_fwd(x, x2, wg, bg, lnw, lnb, 1e-5)
    h = ops.linear_fwd(x, w1, b1, relu=True)
    ops.ffn_out_fwd(h, w2, b2, x, lnw, lnb, 1e-5)
    ops.ffn_fwd(x, w1, b1, w2, b2, lnw, lnb, 1e-5)
torch.cuda.synchronize()
print("ok")
