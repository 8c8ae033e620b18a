import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/d-fine-seg_b200")
import dfine_b200
dev = torch.device("cuda:0")
N = 16000
g = torch.Generator(device=dev).manual_seed(0)
corners = (torch.randn(N, 132, device=dev, generator=g) * 2).bfloat16().requires_grad_(True)
ref = torch.rand(N, 4, device=dev, generator=g) * 0.5 + 0.1
up, rs = torch.tensor([0.5], device=dev), torch.tensor([4.0], device=dev)
project = dfine_b200.fdr_project(up, rs, 32)
gb = torch.randn(N, 4, device=dev, generator=g)
for _ in range(3):
    corners.grad = None
    boxes = dfine_b200.fdr_decode(corners, ref, project, rs, 32)
    boxes.backward(gb)
torch.cuda.synchronize(); print("ok")
