"""Cycle counters of the fused FFN kernel's MMA thread and one epilogue thread (build with -DDFINE_FFN_PROF):
    DFINE_NVCC_EXTRA=-DDFINE_FFN_PROF python -c "import sys; sys.path.insert(0,'d-fine-seg_b200'); from dfine_b200 import build; build.build(force=True)"
"""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/d-fine-seg_b200")
from dfine_b200 import ops, _lib
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
M, C, F = 16000, 256, 1024
x = torch.randn(M, C, device=dev, generator=g)
w1 = (torch.randn(F, C, device=dev, generator=g) * 0.05).bfloat16(); b1 = torch.randn(F, device=dev, generator=g).bfloat16()
w2 = (torch.randn(C, F, device=dev, generator=g) * 0.05).bfloat16(); b2 = torch.randn(C, device=dev, generator=g).bfloat16()
lnw, lnb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for _ in range(3):
    ops.ffn_fwd(x, w1, b1, w2, b2, lnw, lnb, 1e-5)
torch.cuda.synchronize()
buf = np.zeros((256, 8), np.int64)
lib = _lib.lib()
lib.dfine_debug_ffn_prof.argtypes = [ctypes.c_void_p]
assert lib.dfine_debug_ffn_prof(buf.ctypes.data) == 0
p = buf[:125].astype(np.float64) / 1.965e3   # us at 1965 MHz
names = ["mma total", "mma wait weights", "mma wait d1_empty", "mma wait h_full", "epi total", "epi wait d1_full", "epi wait h_empty"]
for i, n in enumerate(names):
    print(f"{n:22s} mean {p[:, i].mean():7.2f} us  min {p[:, i].min():7.2f}  max {p[:, i].max():7.2f}")
