"""Per-CTA phase timeline of the grad_value kernel (needs a build with DFINE_NVCC_EXTRA=-DDFINE_BV_PROF).

    DFINE_NVCC_EXTRA=-DDFINE_BV_PROF python d-fine-seg_b200/dfine_b200/build.py --force
    python tools/bv_phases.py            # on the GPU box
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "d-fine-seg_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dfine_b200 import _lib, ops  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B, Lq, H, c = 32, 500, 8, 32
    shapes, npts = [[80, 80], [40, 40], [20, 20]], [3, 6, 3]
    spec = ops.level_spec(shapes, npts)
    mem = torch.randn(B, spec.L, H * c, device=dev).to(torch.bfloat16)
    cxy = torch.rand(B, Lq, 2, device=dev) * 0.9 + 0.05
    wh = torch.exp(torch.rand(B, Lq, 2, device=dev) * 3.4 - 3.9)
    ref = torch.cat([cxy, wh], -1)
    raw = torch.randn(B, Lq, 3 * H * spec.P, device=dev).to(torch.bfloat16)
    attn_view = raw.reshape(-1)[2 * H * spec.P:]
    rs = raw.shape[-1]
    nps = torch.tensor([1.0 / n for n in npts for _ in range(n)], device=dev)
    go = torch.randn(B, Lq, H * c, device=dev)
    g_raw = torch.empty_like(raw)
    for _ in range(3):
        rec = ops.new_records(mem, spec, H, Lq)
        ops.msda_forward_raw(mem, spec, H, raw, attn_view, ref, nps, 0.5, True, torch.float32,
                             samp_rs=rs, attn_rs=rs, records=rec)
        ops.msda_backward_raw(mem, spec, H, raw, attn_view, ref, nps, 0.5, True, go, gv_dtype=mem.dtype,
                              samp_rs=rs, attn_rs=rs, grad_raw=g_raw, records=rec)
    torch.cuda.synchronize()
    lib = _lib.lib()
    n = 8192
    buf = np.zeros((n, 10), dtype=np.uint64)
    fn = lib.dfine_debug_bv_prof
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rc = fn(buf.ctypes.data, n)
    assert rc == 0, rc
    # CTAs of the LAST launch: stamps within 1 ms of the newest one
    newest = buf[:, 3].max()
    buf = buf[(buf[:, 3] > 0) & (newest - buf[:, 3] < 1_000_000)]
    n = len(buf)
    t = buf[:, :4].astype(np.int64)
    t0 = t[:, 0].min()
    print(f"kernel span {(t[:, 3].max() - t0) / 1e3:.1f} us, {n} CTAs")
    d = np.diff(t, axis=1) / 1e3
    t6 = buf[:, 6:9].astype(np.int64)
    for chunk in sorted(set(buf[:, 5].tolist())):
        m = buf[:, 5] == chunk
        print(f"   chunk {chunk}: start->cleared {(t6[m, 0] - t[m, 0]).mean() / 1e3:.2f}, ->bulk issued {(t6[m, 1] - t[m, 0]).mean() / 1e3:.2f}, "
              f"->first record done {(t6[m, 2] - t[m, 0]).mean() / 1e3:.2f}, ->lists built {(t[m, 1] - t[m, 0]).mean() / 1e3:.2f}; "
              f"walk: touched pixels {(buf[m, 9].astype(np.int64) - t[m, 2]).mean() / 1e3:.2f}, zero fill {(t[m, 3] - buf[m, 9].astype(np.int64)).mean() / 1e3:.2f}")
        print(f"chunk {chunk}: n={m.sum()} build {d[m, 0].mean():.2f} sort {d[m, 1].mean():.2f} walk {d[m, 2].mean():.2f} "
              f"total {(t[m, 3] - t[m, 0]).mean() / 1e3:.2f} us (max {(t[m, 3] - t[m, 0]).max() / 1e3:.2f})")
    # per-SM timeline: gaps between consecutive CTAs on the same SM
    gaps, busy = [], []
    for sm in sorted(set(buf[:, 4].tolist())):
        m = buf[:, 4] == sm
        tt = t[m]
        o = np.argsort(tt[:, 0])
        tt = tt[o]
        gaps += ((tt[1:, 0] - tt[:-1, 3]) / 1e3).tolist()
        busy.append(((tt[:, 3] - tt[:, 0]).sum() / 1e3, len(tt), (tt[-1, 3] - t0) / 1e3, (tt[0, 0] - t0) / 1e3))
    gaps = np.asarray(gaps)
    print(f"gap between CTAs on one SM: mean {gaps.mean():.2f} us, p50 {np.median(gaps):.2f}, max {gaps.max():.2f}")
    b = np.asarray(busy)
    print(f"per SM: CTAs {b[:, 1].min():.0f}..{b[:, 1].max():.0f}, busy {b[:, 0].mean():.1f} us (min {b[:, 0].min():.1f}, max {b[:, 0].max():.1f}), "
          f"first start {b[:, 3].mean():.2f} us, last end mean {b[:, 2].mean():.1f} (min {b[:, 2].min():.1f} max {b[:, 2].max():.1f})")


if __name__ == "__main__":
    main()
