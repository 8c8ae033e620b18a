"""Probe: cuBLAS variants for the Linear weight gradient dW[288,256] = g^T[288,M] x[M,256] (M = 16000, bf16 in,
fp32 out) -- which operand order / split the library picks fastest.  Device-timed, 50 back-to-back launches."""
import torch

def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3

def main():
    torch.manual_seed(0)
    M, N, K = 16000, 288, 256
    g = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    x = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    big = torch.empty(64 << 20, device="cuda", dtype=torch.float32)   # L2 flush between variants
    ref = (g.float().t() @ x.float())
    out = {}
    def chk(name, fn, post=lambda r: r):
        big.zero_()
        us = timed(fn)
        r = post(fn()).float()
        err = float((r - ref).abs().max() / ref.abs().max())
        out[name] = (round(us, 1), err)
        print(f"{name:40s} {us:7.1f} us  rel err {err:.2e}", flush=True)
    chk("g.t() @ x -> f32 (current)", lambda: torch.mm(g.t(), x, out_dtype=torch.float32))
    chk("(x.t() @ g).t() -> f32", lambda: torch.mm(x.t(), g, out_dtype=torch.float32), lambda r: r.t())
    chk("g.t() @ x -> bf16", lambda: torch.mm(g.t(), x))
    for S in (4, 8, 16, 32):
        gs, xs = g.view(S, M // S, N), x.view(S, M // S, K)
        chk(f"bmm split {S} + sum", lambda: torch.bmm(gs.transpose(1, 2), xs, out_dtype=torch.float32).sum(0))
        chk(f"bmm split {S} only", lambda: torch.bmm(gs.transpose(1, 2), xs, out_dtype=torch.float32), lambda r: r.sum(0))
    # ones column appended to x: dW and db from one GEMM
    xp = torch.cat([x, torch.ones(M, 8, device="cuda", dtype=torch.bfloat16)], 1)
    chk("g.t() @ [x|1] (N=264) -> f32", lambda: torch.mm(g.t(), xp, out_dtype=torch.float32), lambda r: r[:, :K])

if __name__ == "__main__":
    main()
