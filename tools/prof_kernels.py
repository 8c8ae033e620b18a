"""Runs the hot-path kernels at a BASELINE.json shape a few times (for ncu / timing).

    python tools/prof_kernels.py [--config m|x|s] [--iters 3] [--dtype bf16|f32] [--mask]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "d-fine-seg_b200"))
import torch  # noqa: E402

from dfine_b200 import ops  # noqa: E402

CFG = {
    "m": dict(B=32, Lq=500, shapes=[[80, 80], [40, 40], [20, 20]], npts=[3, 6, 3]),
    "s": dict(B=64, Lq=300, shapes=[[80, 80], [40, 40], [20, 20]], npts=[3, 6, 3]),
    "x": dict(B=8, Lq=500, shapes=[[128, 128], [64, 64], [32, 32]], npts=[4, 4, 4]),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="m")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--mask", action="store_true")
    a = ap.parse_args()
    cfg = CFG[a.config]
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B, Lq, H, c = cfg["B"], cfg["Lq"], 8, 32
    spec = ops.level_spec(cfg["shapes"], cfg["npts"])
    vdt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    mem = torch.randn(B, spec.L, H * c, device=dev).to(vdt)
    cxy = torch.rand(B, Lq, 2, device=dev) * 0.9 + 0.05
    wh = torch.exp(torch.rand(B, Lq, 2, device=dev) * 3.4 - 3.9)
    ref = torch.cat([cxy, wh], -1)
    ang = torch.arange(H, device=dev) * (2 * torch.pi / H)
    dirs = torch.stack([ang.cos(), ang.sin()], -1)
    dirs = dirs / dirs.abs().max(-1, keepdim=True).values
    rank = torch.cat([torch.arange(1, n + 1) for n in cfg["npts"]]).to(dev).float()
    bias = dirs[:, None, :] * rank[None, :, None]
    raw_off = (bias[None, None] + torch.randn(B, Lq, H, spec.P, 2, device=dev) * 0.3).to(vdt)
    raw_log = torch.randn(B, Lq, H, spec.P, device=dev).to(vdt)
    nps = torch.tensor([1.0 / n for n in cfg["npts"] for _ in range(n)], device=dev)
    go = torch.randn(B, Lq, H * c, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf, tb = [], []
    # the layout bench.py / the patched module use: one concatenated Linear output, geometry
    # records written by the forward, gradients in the inputs' dtypes
    raw = torch.cat([raw_off.reshape(B, Lq, -1), raw_log.reshape(B, Lq, -1)], -1).contiguous()
    attn_view = raw.reshape(-1)[2 * H * spec.P:]
    rs = raw.shape[-1]
    g_raw = torch.empty_like(raw)
    xq = torch.randn(B * Lq, H * c, device=dev).to(torch.bfloat16)   # the Linear's input (bf16 queries)
    for _ in range(a.iters):
        rec = ops.new_records(mem, spec, H, Lq)
        ev[0].record()
        ops.msda_forward_raw(mem, spec, H, raw, attn_view, ref, nps, 0.5, True, torch.float32,
                             samp_rs=rs, attn_rs=rs, records=rec)
        ev[1].record()
        ops.msda_backward_raw(mem, spec, H, raw, attn_view, ref, nps, 0.5, True, go, gv_dtype=mem.dtype,
                              samp_rs=rs, attn_rs=rs, grad_raw=g_raw, records=rec)
        ev[2].record()
        if vdt == torch.bfloat16:   # weight + bias gradient of the concatenated Linear over the kernel's own output
            ops.linear_wgrad(g_raw.reshape(B * Lq, rs), xq)
        torch.cuda.synchronize()
        tf.append(ev[0].elapsed_time(ev[1]))
        tb.append(ev[1].elapsed_time(ev[2]))
    print(f"config {a.config} {a.dtype}: fwd {min(tf)*1e3:.1f} us, bwd(+memset) {min(tb)*1e3:.1f} us")
    if a.mask:
        coef = torch.randn(16, 500, 256, device=dev).to(torch.bfloat16)
        proto = torch.randn(16, 256, 160 * 160, device=dev).to(torch.bfloat16)
        for odt in (torch.bfloat16, torch.float32):
            ts = []
            for _ in range(a.iters):
                ev[0].record()
                ops.mask_gemm_raw(coef, proto, odt, False)
                ev[1].record()
                torch.cuda.synchronize()
                ts.append(ev[0].elapsed_time(ev[1]))
            ref_t = []
            for _ in range(a.iters):
                ev[0].record()
                torch.bmm(coef, proto) if odt == torch.bfloat16 else torch.bmm(coef.float(), proto.float())
                ev[1].record()
                torch.cuda.synchronize()
                ref_t.append(ev[0].elapsed_time(ev[1]))
            flops = 2 * 16 * 500 * 256 * 25600
            print(f"mask gemm out={odt}: ours {min(ts)*1e3:.1f} us ({flops/min(ts)/1e9:.1f} TFLOP/s), "
                  f"torch.bmm {min(ref_t)*1e3:.1f} us")


if __name__ == "__main__":
    main()
