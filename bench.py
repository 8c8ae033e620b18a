#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native D-FINE-seg decoder hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun)
    python bench.py --impl reference ...                      (CPU arm: the reference's torch path)

A "step" is one pass of the hot path over one batch of synthetic input at BASELINE.json's
config 3 (D-FINE-m, 640x640, batch 32 per GPU, bf16 autocast, Lq = 300 + 200 denoising
queries): for each of the 4 decoder layers MSDeformableAttention forward (2 Linears + fused
sampling kernel) and the FDR box decode, then the backward of all of it (gradients to
`memory`, the queries, the Linear parameters and pred_corners).  `value` is whole-job
images/s with inputs resident in HBM; `e2e` is the same step with every input copied from
pinned host memory and the decoded boxes read back inside the timed region.

The reference arm and `cpu_baseline` time oracle/torch_port.py -- the reference's own
PyTorch (ATen grid_sample) call sequence -- on the host cores; they are the only places this
file touches oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "d-fine-seg_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "D-FINE-m 640x640 train imgs/s through the decoder hot path (MSDeformAttn fwd+bwd + FDR)"
UNIT = "imgs/s"

WORKLOADS = {
    # BASELINE.json configs[2]: D-FINE-m detection training, 640x640, batch 32 / GPU
    "dfine_m_train_640_b32": dict(B=32, Lq=500, C=256, H=8, shapes=[[80, 80], [40, 40], [20, 20]],
                                  npts=[3, 6, 3], layers=4, reg_max=32, up=0.5, reg_scale=4.0),
}
DEFAULT_WORKLOAD = "dfine_m_train_640_b32"


# ------------------------------------------------------------------------------------------
# synthetic inputs: object-like boxes + 10 % border / oversize boxes,
# "trained-like" Linear weights on top of the reference's directional bias init
# ------------------------------------------------------------------------------------------
def make_inputs(wl: dict, B: int, seed: int, device: str):
    g = torch.Generator().manual_seed(seed)
    Lq, C, H, layers = wl["Lq"], wl["C"], wl["H"], wl["layers"]
    L = sum(h * w for h, w in wl["shapes"])
    P = sum(wl["npts"])

    def boxes():
        cxy = torch.rand(B, Lq, 2, generator=g) * 0.9 + 0.05
        wh = torch.exp(torch.rand(B, Lq, 2, generator=g) * 3.4 - 3.9)  # logU(0.02, 0.6)
        big = torch.rand(B, Lq, 1, generator=g) < 0.10
        wh = torch.where(big, torch.rand(B, Lq, 2, generator=g) * 0.6 + 0.6, wh)
        cxy = torch.where(big, torch.rand(B, Lq, 2, generator=g) * 1.2 - 0.1, cxy)
        return torch.cat([cxy, wh], -1)

    inp = dict(
        memory=torch.randn(B, L, C, generator=g),
        queries=[torch.randn(B, Lq, C, generator=g) for _ in range(layers)],
        refs=[boxes().unsqueeze(2) for _ in range(layers)],
        corners=[torch.randn(B, Lq, 4 * (wl["reg_max"] + 1), generator=g) * 2 for _ in range(layers)],
        ref_init=boxes(),
        grad_outs=[torch.randn(B, Lq, C, generator=g) for _ in range(layers)],
        grad_boxes=[torch.randn(B, Lq, 4, generator=g) for _ in range(layers)],
    )
    lin = []
    for _ in range(layers):
        lin.append(dict(so_w=torch.randn(H * P * 2, C, generator=g) * 0.02,
                        aw_w=torch.randn(H * P, C, generator=g) * 0.02,
                        aw_b=torch.randn(H * P, generator=g) * 0.1))
    inp["lin"] = lin
    return inp


def algorithmic_bytes(wl: dict, B: int, e_g: int = 2):
    """SURVEY.md section 8(d), per decoder layer: compulsory traffic, every tensor touched
    once.  bf16 value (e_v = 2), fp32 out / grad_out (e_o = 4); e_g is the element size of the
    grad_value the kernel actually writes: 2 here (bf16 written directly under AMP -- the
    SURVEY formula assumed an fp32 buffer, e_g = 4, which would flatter the result)."""
    L = sum(h * w for h, w in wl["shapes"])
    P = sum(wl["npts"])
    blc, samples, bqc = B * L * wl["C"], B * wl["Lq"] * wl["H"] * P, B * wl["Lq"] * wl["C"]
    e_v, e_o = 2, 4
    fwd = blc * e_v + samples * 12 + bqc * e_o
    bwd = bqc * e_o + blc * e_v + samples * 12 + blc * e_g + samples * 12
    return fwd, bwd


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class HotPath:
    """The patched decoder hot path over one batch, driven through the package's public API."""

    def __init__(self, wl: dict, inp: dict, device: torch.device, world: int = 1,
                 sync_mode: str = "eager"):
        import dfine_b200
        self.api = dfine_b200
        self.wl, self.dev = wl, device
        # data parallel (world > 1): the gradients of the path's own parameters (the two Linears
        # of every layer) are averaged over the ranks each step -- what DDP does for them in the
        # reference (src/dl/train.py:161-166).  "overlap" (measured SLOWER than "eager": 1.041 vs 1.005 ms at
        # N = 2 -- four latency-bound collectives whose kernels take SMs from the backward cost more than
        # the one exposed all-reduce they replace; kept selectable): one bucket per decoder layer, its
        # NCCL all-reduce forked onto a side stream the moment the layer's weight-gradient kernel is
        # done and captured INSIDE the step's CUDA graph, so it runs under the backward kernels of
        # the remaining layers (DDP's bucketed overlap; only the first layer's exchange is exposed).
        # "eager": one all-reduce of everything on the compute stream after every graph replay (fully
        # exposed); "off": no exchange.  Graphs holding collectives are destroyed before the
        # process group (HotPath.close) -- leaving them alive hung the ranks at teardown.
        # "nvls" (default): NO collective launch -- dfine_multicast_add adds every rank's gradient blocks into
        # all replicas of a symmetric buffer through its NVLS multicast mapping (multimem.red: the sum over
        # the ranks is formed inside the NVSwitch); two tiny device-side barriers per step
        # (grad_sync.NvlsGradExchange).
        self.world, self.sync_mode = world, (sync_mode if world > 1 else "off")
        self.bucket = None
        self.lsync = None
        self.nvls = None
        self.nvls_error = None
        self.mods = []
        for lw in inp["lin"]:
            m = dfine_b200.MSDeformableAttention(wl["C"], wl["H"], len(wl["shapes"]), wl["npts"])
            with torch.no_grad():
                m.sampling_offsets.weight.copy_(lw["so_w"])
                m.attention_weights.weight.copy_(lw["aw_w"])
                m.attention_weights.bias.copy_(lw["aw_b"])
            self.mods.append(m.to(device))
        self.up = torch.tensor([wl["up"]], device=device)
        self.reg_scale = torch.tensor([wl["reg_scale"]], device=device)
        self.host = None
        self.d = None
        self.g_grads = None
        if self.sync_mode != "off":
            from dfine_b200 import grad_sync
            self.bucket = grad_sync.GradBucket(grad_sync.path_parameters(self.mods))
            if self.sync_mode == "overlap":
                self.lsync = grad_sync.LayerwiseGradSync(self.mods)
            if self.sync_mode == "nvls":
                P = sum(wl["npts"])
                n_out, K = 3 * wl["H"] * P, wl["C"]
                try:
                    self.nvls = grad_sync.NvlsGradExchange([n_out * K + n_out] * wl["layers"], device)
                except Exception as exc:  # noqa: BLE001  (no multicast on this box: NCCL after the replay)
                    self.nvls_error = f"{type(exc).__name__}: {exc}"[:200]
                    self.sync_mode = "eager"

    # every input of a step lives in ONE slab (256-byte aligned slots): the device copy is what the CUDA
    # graph reads, the pinned host copy is what an end-to-end step sends -- one cudaMemcpyAsync of the
    # whole slab instead of one small copy per tensor (22 at config 3)
    _DTYPES = dict(memory=torch.bfloat16, corners=torch.bfloat16)

    def _layout(self, inp: dict):
        slots, off = [], 0
        for k in ("memory", "queries", "refs", "corners", "ref_init", "grad_outs", "grad_boxes"):
            v = inp[k]
            dt = self._DTYPES.get(k, torch.float32)
            for i, t in enumerate(v if isinstance(v, list) else [v]):
                nbytes = t.numel() * torch.empty((), dtype=dt).element_size()
                slots.append((k, i if isinstance(v, list) else None, dt, tuple(t.shape), off, nbytes))
                off += (nbytes + 255) & ~255
        return slots, off

    def _views(self, slab: torch.Tensor, slots):
        d = {}
        for k, i, dt, shape, off, nbytes in slots:
            t = slab[off:off + nbytes].view(dt).view(shape)
            if i is None:
                d[k] = t
            else:
                d.setdefault(k, []).append(t)
        return d

    def load_device(self, inp: dict):
        slots, total = self._layout(inp)
        self.d_slab = torch.empty(total, dtype=torch.uint8, device=self.dev)
        self.d = self._views(self.d_slab, slots)
        for k, i, dt, shape, off, nbytes in slots:
            src = inp[k] if i is None else inp[k][i]
            (self.d[k] if i is None else self.d[k][i]).copy_(src.to(dt))

    def pin_host(self, inp: dict):
        slots, total = self._layout(inp)
        self.h_slab = torch.empty(total, dtype=torch.uint8).pin_memory()
        self.host = self._views(self.h_slab, slots)
        for k, i, dt, shape, off, nbytes in slots:
            src = inp[k] if i is None else inp[k][i]
            (self.host[k] if i is None else self.host[k][i]).copy_(src.to(dt))
        self.h2d_bytes = total
        self.h2d_copies = 1
        L = len(inp["queries"])
        self.boxes_host = torch.empty((L, *inp["grad_boxes"][0].shape), dtype=torch.float32).pin_memory()
        self.d2h_bytes = self.boxes_host.numel() * 4

    def step(self, d: dict):
        wl = self.wl
        mem = d["memory"].detach().requires_grad_(True)
        queries = [q.detach().requires_grad_(True) for q in d["queries"]]
        corners = [c.detach().requires_grad_(True) for c in d["corners"]]
        for m in self.mods:
            m.zero_grad(set_to_none=True)
        if self.nvls is not None:
            self.nvls.begin_step()    # zero the local replica + barrier A (hidden under the forward)
        outs, boxes = [], []
        with torch.autocast("cuda", dtype=torch.bfloat16):
            project = self.api.fdr_project(self.up, self.reg_scale, wl["reg_max"])
            # TransformerDecoder.value_op: per-level strided views of `memory` (zero copy)
            B, L, C = mem.shape
            value = mem.reshape(B, L, wl["H"], C // wl["H"]).permute(0, 2, 3, 1).split(
                [h * w for h, w in wl["shapes"]], dim=-1)
            for i, m in enumerate(self.mods):
                outs.append(m(queries[i], d["refs"][i], value, wl["shapes"]))
                boxes.append(self.api.fdr_decode(corners[i], d["ref_init"], project, self.reg_scale,
                                                 wl["reg_max"]))
        if self.nvls is not None:
            with self.nvls:           # dW / db of every layer land in the exchange's local blocks
                torch.autograd.backward(outs + boxes, d["grad_outs"] + d["grad_boxes"])
            self.nvls.end_step()      # multimem.red into all replicas + barrier B: the rank average
        else:
            torch.autograd.backward(outs + boxes, d["grad_outs"] + d["grad_boxes"])
        if self.lsync is not None:
            self.lsync.finish()       # the compute stream joins the per-layer exchanges
        return boxes, mem.grad

    def check_exchange(self):
        """N > 1 self-check of the in-kernel exchange: the gradients an "nvls" step leaves in param.grad
        against the same step's local gradients averaged by an NCCL all-reduce.  Max error relative to the
        largest gradient element."""
        import torch.distributed as dist
        if self.nvls is None:
            return None
        params = self.bucket.params
        self.step(self.d)
        got = [p.grad.detach().clone() for p in params]
        ex, self.nvls = self.nvls, None
        try:
            self.step(self.d)
        finally:
            self.nvls = ex
        err = 0.0
        for g, p in zip(got, params):
            want = p.grad.detach().clone()
            dist.all_reduce(want, op=dist.ReduceOp.AVG)
            err = max(err, float((g - want).abs().max() / want.abs().max().clamp_min(1e-30)))
        return err

    def step_eager(self, d: dict):
        """step() plus the gradient all-reduce a graph replay is followed by (N > 1)."""
        r = self.step(d)
        if self.sync_mode == "eager":
            self.bucket.reduce()
        return r

    def capture(self, warmup: int = 3):
        """Capture one step (forward + backward of all layers) over the static device
        buffers self.d into a CUDA graph: the ~120 launches of a step replay without any
        Python / launch latency between them."""
        from dfine_b200 import ops
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.step(self.d)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        n0 = ops.LAUNCHES["count"]
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            boxes, mem_grad = self.step(self.d)
            self.g_boxes = torch.stack(boxes)
            self.g_mem_grad = mem_grad
        if self.bucket is not None:   # the graph's own (static) parameter-gradient tensors
            self.g_grads = [p.grad for p in self.bucket.params]
        self.launches_per_step = ops.LAUNCHES["count"] - n0
        return self.graph

    def replay(self):
        self.graph.replay()
        if self.sync_mode == "eager":
            self.bucket.reduce(self.g_grads)

    def close(self):
        """Drop the CUDA graphs (they hold captured collectives in "overlap" mode) and drain the device:
        call before torch.distributed.destroy_process_group()."""
        for st in getattr(self, "sets", []):
            st.pop("graph", None)
        self.sets = []
        self.graph = None
        if self.lsync is not None:
            self.lsync.remove()
        torch.cuda.synchronize(self.dev)

    def setup_pipeline(self, inp: dict):
        """Second set of static buffers + graph, a copy stream and events: the H2D copies of
        step i+1 overlap the compute of step i (copies and kernels on separate streams)."""
        first = dict(d=self.d, slab=self.d_slab, graph=self.graph, boxes=self.g_boxes, grads=self.g_grads)
        self.load_device(inp)             # fresh static buffers -> self.d
        self.capture(warmup=3)            # second graph over them
        second = dict(d=self.d, slab=self.d_slab, graph=self.graph, boxes=self.g_boxes, grads=self.g_grads)
        self.d, self.d_slab, self.graph, self.g_boxes, self.g_grads = (first["d"], first["slab"], first["graph"],
                                                                        first["boxes"], first["grads"])
        self.sets = [first, second]
        self.copy_stream = torch.cuda.Stream(self.dev)
        for st in self.sets:
            st["copied"] = torch.cuda.Event()
            st["done"] = torch.cuda.Event()
            st["done"].record(torch.cuda.current_stream(self.dev))
            st["host_boxes"] = torch.empty_like(self.boxes_host).pin_memory()
        self.pipe_i = 0

    def step_e2e_pipelined(self):
        """One step of the double-buffered pipeline: enqueue H2D of this step's inputs on the
        copy stream (after the previous user of the buffer set finished), then replay the
        graph and read the boxes back on the compute stream.  Host waits only for the step
        issued two calls ago, so copies of step i+1 overlap kernels of step i."""
        st = self.sets[self.pipe_i & 1]
        self.pipe_i += 1
        comp = torch.cuda.current_stream(self.dev)
        st["done"].synchronize()          # result of the step that last used this set is on the host
        with torch.cuda.stream(self.copy_stream):
            st["slab"].copy_(self.h_slab, non_blocking=True)      # ONE copy: every input of the step
            st["copied"].record(self.copy_stream)
        comp.wait_event(st["copied"])
        st["graph"].replay()
        if self.sync_mode == "eager":
            self.bucket.reduce(st["grads"])
        st["host_boxes"].copy_(st["boxes"], non_blocking=True)
        st["done"].record(comp)
        return st["host_boxes"]

    def drain_pipeline(self):
        for st in self.sets:
            st["done"].synchronize()

    def step_e2e(self):
        """Host buffers in, host result out: H2D of every input of the step from pinned
        memory into the graph's static buffers, graph replay, D2H of the decoded boxes."""
        self.d_slab.copy_(self.h_slab, non_blocking=True)
        self.replay()
        self.boxes_host.copy_(self.g_boxes, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return self.boxes_host


def kernel_leg(wl: dict, hp: "HotPath", device, reps: int = 20):
    """Per-kernel durations behind the roofline figures: every C-ABI kernel of one decoder layer
    launched `reps` times back to back on the current stream (the device never idles between
    launches, so the CUDA-event bracket measures kernel time, not host launch gaps), on the
    step's own layer-0 tensors.  Between timed kernels of the same kind nothing is flushed: the
    working set of one launch (memory 137.6 MB + gradient 137.6 MB + ...) is larger than L2."""
    from dfine_b200 import ops
    import torch.nn.functional as F
    d, m = hp.d, hp.mods[0]
    H = wl["H"]
    spec = ops.level_spec(wl["shapes"], wl["npts"])
    mem = d["memory"]
    B, Lq = d["queries"][0].shape[:2]
    with torch.no_grad():
        w = torch.cat([m.sampling_offsets.weight, m.attention_weights.weight], 0).to(torch.bfloat16)
        bias = torch.cat([m.sampling_offsets.bias, m.attention_weights.bias], 0).to(torch.bfloat16)
        raw = F.linear(d["queries"][0].to(torch.bfloat16), w, bias).contiguous()
    attn_view = raw.reshape(-1)[2 * H * spec.P:]
    rs = raw.shape[-1]
    ref = d["refs"][0].reshape(B, Lq, 4).float().contiguous()
    nps = m.num_points_scale.float().contiguous()
    go = d["grad_outs"][0].contiguous()
    g_raw = torch.empty_like(raw)
    rec = ops.new_records(mem, spec, H, Lq)
    run = torch.zeros((B, spec.L, mem.shape[-1]), dtype=mem.dtype, device=device)
    corners, ref_init, gb = d["corners"][0], d["ref_init"], d["grad_boxes"][0]
    project = hp.api.fdr_project(hp.up, hp.reg_scale, wl["reg_max"])

    def fwd():
        ops.msda_forward_raw(mem, spec, H, raw, attn_view, ref, nps, 0.5, True, torch.float32,
                             samp_rs=rs, attn_rs=rs, records=rec)

    def bwd(acc):
        ops.msda_backward_raw(mem, spec, H, raw, attn_view, ref, nps, 0.5, True, go, gv_dtype=mem.dtype,
                              samp_rs=rs, attn_rs=rs, grad_raw=g_raw, records=rec, accumulate_into=acc)

    def fdr():
        hp.api.fdr_decode(corners, ref_init, project, hp.reg_scale, wl["reg_max"])

    legs = {"msda_fwd": fwd, "msda_bwd": lambda: bwd(None), "msda_bwd_accumulate": lambda: bwd(run),
            "fdr_fwd": fdr}
    out = {}
    for name, fn in legs.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize(device)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize(device)
        out[name] = s.elapsed_time(e) / reps
    return out


def secondary_kernel_legs(device, peak: float, reps: int = 10):
    """Kernel-level numbers at the other BASELINE.json configs (they are parity-test cases, not
    the bench line): config 2 (D-FINE-s inference, B=64, Lq=300: forward only), config 5
    (1024x1024, B=8, Lq=500, 3 levels x 4 points: forward + backward) and the config-4 mask
    assembly (16 x 500 x 256 x 160^2, bf16 out).  Same back-to-back timing as kernel_leg."""
    import dfine_b200
    from dfine_b200 import ops
    H, c = 8, 32
    out = {}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(device)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize(device)
        return s.elapsed_time(e) / reps

    cfgs = {
        "config2_dfine_s_infer_b64": dict(B=64, Lq=300, shapes=[[80, 80], [40, 40], [20, 20]], npts=[3, 6, 3], bwd=False),
        "config5_dfine_x_1024_b8": dict(B=8, Lq=500, shapes=[[128, 128], [64, 64], [32, 32]], npts=[4, 4, 4], bwd=True),
    }
    g = torch.Generator(device=device).manual_seed(7)
    for name, cf in cfgs.items():
        B, Lq = cf["B"], cf["Lq"]
        spec = ops.level_spec(cf["shapes"], cf["npts"])
        P = spec.P
        mem = torch.randn(B, spec.L, H * c, device=device, generator=g).to(torch.bfloat16)
        ref = torch.cat([torch.rand(B, Lq, 2, device=device, generator=g) * 0.9 + 0.05,
                         torch.exp(torch.rand(B, Lq, 2, device=device, generator=g) * 3.4 - 3.9)], -1)
        raw = torch.randn(B, Lq, 3 * H * P, device=device, generator=g).to(torch.bfloat16)
        attn_view = raw.reshape(-1)[2 * H * P:]
        rs = raw.shape[-1]
        nps = torch.tensor([1.0 / n for n in cf["npts"] for _ in range(n)], device=device)
        rec = ops.new_records(mem, spec, H, Lq) if cf["bwd"] else None
        blc, smp, bqc = mem.numel(), B * Lq * H * P, B * Lq * H * c
        fwd_bytes = blc * 2 + smp * 12 + bqc * 4
        res = {}
        ms = timed(lambda: ops.msda_forward_raw(mem, spec, H, raw, attn_view, ref, nps, 0.5, True, torch.float32,
                                                samp_rs=rs, attn_rs=rs, records=rec))
        res["msda_fwd"] = {"ms": ms, "algorithmic_bytes": fwd_bytes, "frac": fwd_bytes / (ms / 1e3) / 1e9 / peak,
                           "note": "SURVEY's formula charges the whole `memory` tensor once; the gather touches a subset "
                                   "of it (ncu at config 3: 112.8 MB read of 156 MB algorithmic), so this fraction is an "
                                   "upper-bound reading of the kernel's DRAM efficiency, most of all at large B x L"}
        if cf["bwd"]:
            go = torch.randn(B, Lq, H * c, device=device, generator=g)
            g_raw = torch.empty_like(raw)
            bwd_bytes = bqc * 4 + blc * 2 + smp * 12 + blc * 2 + smp * 12
            ms = timed(lambda: ops.msda_backward_raw(mem, spec, H, raw, attn_view, ref, nps, 0.5, True, go,
                                                     gv_dtype=mem.dtype, samp_rs=rs, attn_rs=rs, grad_raw=g_raw,
                                                     records=rec))
            res["msda_bwd"] = {"ms": ms, "algorithmic_bytes": bwd_bytes, "frac": bwd_bytes / (ms / 1e3) / 1e9 / peak}
        out[name] = res
        del mem, raw
    Bm, M, K, N = 16, 500, 256, 160 * 160
    coef = torch.randn(Bm, M, K, device=device, generator=g).to(torch.bfloat16)
    proto = torch.randn(Bm, K, N, device=device, generator=g).to(torch.bfloat16)
    ms = timed(lambda: ops.mask_gemm_raw(coef, proto, torch.bfloat16, False))
    mbytes = Bm * (M * K * 2 + K * N * 2 + M * N * 2)
    flops = 2.0 * Bm * M * K * N
    m4 = {"mask_gemm_fwd": {"ms": ms, "algorithmic_bytes": mbytes, "frac": mbytes / (ms / 1e3) / 1e9 / peak,
                            "tflops": flops / (ms / 1e3) / 1e12}}
    # the library bar for the same contraction, same run (cuBLAS batched GEMM through torch.bmm)
    ms_lib = timed(lambda: torch.bmm(coef, proto))
    m4["mask_gemm_fwd"]["cublas_bmm_ms"] = ms_lib
    m4["mask_gemm_fwd"]["vs_cublas"] = ms_lib / ms
    # backward: grad_proto = coef^T x go (reads go, writes [B, K, N]); grad_coef = go x proto^T (reads go + proto)
    go = torch.randn(Bm, M, N, device=device, generator=g).to(torch.bfloat16)
    for key, wc, wp, nbytes in (("mask_gemm_bwd_proto", False, True, Bm * (M * K * 2 + M * N * 2 + K * N * 2)),
                                ("mask_gemm_bwd_coef", True, False, Bm * (M * N * 2 + K * N * 2 + M * K * 4))):
        ms = timed(lambda: ops.mask_gemm_bwd_raw(coef, proto, go, want_coef=wc, want_proto=wp))
        ms_lib = timed((lambda: torch.bmm(coef.transpose(1, 2), go)) if wp else
                       (lambda: torch.bmm(go, proto.transpose(1, 2))))
        m4[key] = {"ms": ms, "algorithmic_bytes": nbytes, "frac": nbytes / (ms / 1e3) / 1e9 / peak,
                   "tflops": flops / (ms / 1e3) / 1e12, "cublas_bmm_ms": ms_lib, "vs_cublas": ms_lib / ms}
    # fused focal + dice over the matched rows: 16 images x 100 denoising rows of 160 x 160 (bf16 logits)
    Mr = 1600
    xl = (torch.randn(Mr, N, device=device, generator=g) * 3).to(torch.bfloat16)
    tl = (torch.rand(Mr, N, device=device, generator=g) < 0.2).float()
    ms = timed(lambda: ops._MaskLossFn.apply(xl, tl))
    nb = Mr * N * (2 + 4)
    m4["mask_loss_fwd"] = {"ms": ms, "rows": Mr, "algorithmic_bytes": nb, "frac": nb / (ms / 1e3) / 1e9 / peak}
    st = ops._MaskLossFn.apply(xl, tl)
    gst = torch.randn(Mr, 4, device=device, generator=g)
    gl = torch.empty_like(xl)

    def bwd():
        rc = dfine_b200._lib.lib().dfine_mask_loss_bwd(xl.data_ptr(), 1, N, tl.data_ptr(), Mr, N, st.data_ptr(),
                                                       gst.data_ptr(), gl.data_ptr(), 1,
                                                       torch.cuda.current_stream(device).cuda_stream)
        assert rc == 0
    ms = timed(bwd)
    nb = Mr * N * (2 + 4 + 2)
    m4["mask_loss_bwd"] = {"ms": ms, "rows": Mr, "algorithmic_bytes": nb, "frac": nb / (ms / 1e3) / 1e9 / peak}
    out["config4_mask_assembly_bf16"] = m4
    out["config3_decoder_layer_kernels"] = decoder_layer_legs(device, peak, graph_timed(device))
    return out


def graph_timed(device, reps: int = 20):
    """Timer for kernels shorter than their host launch path (10-30 us): `reps` calls are captured into ONE
    CUDA graph and the replay is bracketed by two events -- device time per call, no host gaps (for the
    library op sequences it is compared with as well)."""
    def timed(fn):
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize(device)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        g.replay()
        e.record()
        torch.cuda.synchronize(device)
        del g
        return s.elapsed_time(e) / reps
    return timed


def decoder_layer_legs(device, peak: float, timed):
    """The fused Linear kernels of the decoder layer at the config-3 shape
    (M = B*Lq = 16000 rows, C = 256, FFN 1024), each beside the reference's own op sequence under bf16
    autocast on the same GPU (library bar: elementwise kernels + cuBLAS).  Both rooflines are reported: the
    algorithmic HBM bytes (every operand and result touched once) and the tensor-core rate."""
    import torch.nn.functional as F
    from dfine_b200 import ops
    g = torch.Generator(device=device).manual_seed(11)
    M, C, Fd, N = 16000, 256, 1024, 288
    x = torch.randn(M, C, device=device, generator=g)
    pos = torch.randn(M, C, device=device, generator=g)
    x2 = torch.randn(M, C, device=device, generator=g)
    w = (torch.randn(N, C, device=device, generator=g) * 0.05)
    b = torch.randn(N, device=device, generator=g)
    wg = torch.randn(2 * C, 2 * C, device=device, generator=g) * 0.05
    bg = torch.randn(2 * C, device=device, generator=g)
    w1 = torch.randn(Fd, C, device=device, generator=g) * 0.05
    b1 = torch.randn(Fd, device=device, generator=g)
    w2 = torch.randn(C, Fd, device=device, generator=g) * 0.05
    b2 = torch.randn(C, device=device, generator=g)
    lnw, lnb = torch.ones(C, device=device), torch.zeros(C, device=device)
    wb, bb, wgb, bgb = w.bfloat16(), b.bfloat16(), wg.bfloat16(), bg.bfloat16()
    w1b, b1b, w2b, b2b = w1.bfloat16(), b1.bfloat16(), w2.bfloat16(), b2.bfloat16()
    h = ops.linear_fwd(x, w1b, b1b, relu=True)
    res = {}

    def entry(ms, nbytes, flops, ms_lib, what):
        return {"ms": ms, "algorithmic_bytes": nbytes, "frac": nbytes / (ms / 1e3) / 1e9 / peak,
                "tflops": flops / (ms / 1e3) / 1e12, "reference_ops_ms": ms_lib, "vs_reference_ops": ms_lib / ms,
                "reference_ops": what}

    def ac(fn):
        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                return fn()
        return run

    res["msda_linear_fwd"] = entry(
        timed(lambda: ops.linear_fwd(x, wb, bb, x_add=pos)), M * C * 8 + N * C * 2 + M * N * 2, 2.0 * M * N * C,
        timed(ac(lambda: F.linear(x + pos, w, b))), "add + 3 casts + cuBLAS GEMM (autocast F.linear(x + pos))")
    res["msda_linear_fwd_train"] = entry(
        timed(lambda: ops.linear_fwd(x, wb, bb, x_add=pos, save_input=True)),
        M * C * 8 + N * C * 2 + M * N * 2 + M * C * 2, 2.0 * M * N * C,
        timed(ac(lambda: F.linear(x + pos, w, b))), "same (autograd keeps the bf16 cast of the input)")
    res["gate_fwd"] = entry(
        timed(lambda: ops.gate_fwd(x, x2, wgb, bgb, lnw, lnb, 1e-5)), M * C * 12 + 4 * C * C * 2, 2.0 * M * 4 * C * C,
        timed(ac(lambda: F.layer_norm((lambda gt: gt[..., :C] * x + gt[..., C:] * x2)(
            torch.sigmoid(F.linear(torch.cat([x, x2], -1), wg, bg))), (C,), lnw, lnb, 1e-5))),
        "cat + casts + cuBLAS GEMM + sigmoid + 2 mul + add + LayerNorm (Gate.forward)")
    res["ffn_linear1_relu"] = entry(
        timed(lambda: ops.linear_fwd(x, w1b, b1b, relu=True)), M * C * 4 + Fd * C * 2 + M * Fd * 2, 2.0 * M * Fd * C,
        timed(ac(lambda: F.relu(F.linear(x, w1, b1)))), "casts + cuBLAS GEMM + ReLU")
    res["ffn_out_fwd"] = entry(
        timed(lambda: ops.ffn_out_fwd(h, w2b, b2b, x, lnw, lnb, 1e-5)), M * Fd * 2 + C * Fd * 2 + M * C * 8,
        2.0 * M * Fd * C,
        timed(ac(lambda: F.layer_norm((x + F.linear(h, w2, b2)).clamp(min=-65504, max=65504), (C,), lnw, lnb, 1e-5))),
        "cast + cuBLAS GEMM + add + clamp + LayerNorm")
    res["ffn_fwd_one_kernel"] = entry(
        timed(lambda: ops.ffn_fwd(x, w1b, b1b, w2b, b2b, lnw, lnb, 1e-5)), M * C * 8 + 2 * Fd * C * 2, 4.0 * M * Fd * C,
        res["ffn_linear1_relu"]["reference_ops_ms"] + res["ffn_out_fwd"]["reference_ops_ms"],
        "casts + 2 cuBLAS GEMMs + ReLU + add + clamp + LayerNorm (forward_ffn + norm3)")
    res["ffn_fwd_one_kernel"]["two_kernel_route_ms"] = res["ffn_linear1_relu"]["ms"] + res["ffn_out_fwd"]["ms"]
    tail = res["gate_fwd"]["ms"] + res["ffn_fwd_one_kernel"]["ms"]
    tail_ref = res["gate_fwd"]["reference_ops_ms"] + res["ffn_linear1_relu"]["reference_ops_ms"] + \
        res["ffn_out_fwd"]["reference_ops_ms"]
    res["layer_tail_total"] = {"ms": tail, "reference_ops_ms": tail_ref, "vs_reference_ops": tail_ref / tail,
                               "launches": 2}
    return res


def inference_leg(device, steps: int = 30):
    """BASELINE config 2 through the same public API, forward only: D-FINE-s decoder hot path at
    batch 64, 300 queries, 3 decoder layers (module forward + FDR decode per layer) under
    no_grad + bf16 autocast, replayed from a CUDA graph.  Informational (not the bench line)."""
    import dfine_b200
    B, Lq, C, H, layers = 64, 300, 256, 8, 3
    shapes, npts = [[80, 80], [40, 40], [20, 20]], [3, 6, 3]
    L = sum(h * w for h, w in shapes)
    g = torch.Generator(device=device).manual_seed(11)
    mods = []
    for _ in range(layers):
        m = dfine_b200.MSDeformableAttention(C, H, len(shapes), npts).to(device)
        with torch.no_grad():
            m.sampling_offsets.weight.normal_(0, 0.02, generator=g)
            m.attention_weights.weight.normal_(0, 0.02, generator=g)
        mods.append(m.eval())
    mem = torch.randn(B, L, C, device=device, generator=g).to(torch.bfloat16)
    queries = [torch.randn(B, Lq, C, device=device, generator=g) for _ in range(layers)]
    ref = torch.cat([torch.rand(B, Lq, 2, device=device, generator=g) * 0.9 + 0.05,
                     torch.exp(torch.rand(B, Lq, 2, device=device, generator=g) * 3.4 - 3.9)], -1)
    corners = [(torch.randn(B, Lq, 132, device=device, generator=g) * 2).to(torch.bfloat16) for _ in range(layers)]
    up, rs = torch.tensor([0.5], device=device), torch.tensor([4.0], device=device)

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            project = dfine_b200.fdr_project(up, rs, 32)
            value = mem.reshape(B, L, H, C // H).permute(0, 2, 3, 1).split([h * w for h, w in shapes], dim=-1)
            outs = []
            for i, m in enumerate(mods):
                y = m(queries[i], ref.unsqueeze(2), value, shapes)
                outs.append(dfine_b200.fdr_decode(corners[i], ref, project, rs, 32))
            return y, torch.stack(outs)

    side = torch.cuda.Stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream(device).wait_stream(side)
    torch.cuda.synchronize(device)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep = step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(device)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        graph.replay()
    e.record()
    torch.cuda.synchronize(device)
    ms = s.elapsed_time(e) / steps
    del keep
    return {"workload": "dfine_s_infer_640_b64 (decoder hot path, forward only, 3 layers, CUDA graph)",
            "ms_per_step": ms, "imgs_per_s": B / (ms / 1e3)}


# ------------------------------------------------------------------------------------------
# full model: the UNMODIFIED reference model (baseline/_ref) with and without the patched hot path
# ------------------------------------------------------------------------------------------
def full_model_leg(device, rank: int, world: int, dist_on: bool, steps: int, warmup: int):
    """BASELINE configs 3 and 2 through the real model on this GPU: `build_model` of the reference
    (baseline/_ref/src/d_fine/dfine.py:51-73), the reference criterion + Hungarian matcher, AdamW from the
    reference's `build_optimizer`, the train step restated from src/dl/train.py:545-557 / :488-511 (bf16
    autocast, clip 0.1) -- once unpatched (the reference's own eager CUDA path: the bar), once with
    `dfine_b200.patch_model`.  Every step copies its images from pinned host memory (end to end: only
    images and targets cross PCIe).  N > 1: both arms are wrapped in DistributedDataParallel exactly as
    src/dl/train.py:161-166 does, so the curve includes the real gradient exchange (78 MB for D-FINE-m)."""
    import copy

    import dfine_b200
    from baseline import model_harness as MH
    from baseline import ref_install
    from dfine_b200 import ops

    if not ref_install.installed():
        return {"unavailable": "baseline/_ref (reference model package) is not installed: run __graft_entry__.build()"}
    try:
        from loguru import logger
        logger.remove()
    except Exception:  # noqa: BLE001
        pass
    out = {}

    def timed(fn, n, w):
        return time_steps(fn, n, w, device, dist_on) / n

    # ---- config 3: D-FINE-m training, 640x640, batch 32 per GPU ----
    B = 32
    model, loss_fn = MH.build("m", device, 640, False, seed=0)
    model.train(), loss_fn.train()
    patched = copy.deepcopy(model)
    counts = dfine_b200.patch_model(patched, layer=True)
    # the patched arm also routes the criterion's matching stage through the device (dfine_lsap):
    # same assignments, same loss terms, no host round trip of the cost matrices
    patched_loss = copy.deepcopy(loss_fn)
    counts.update(dfine_b200.patch_criterion(patched_loss))
    arms = {"reference": (model, loss_fn), "patched": (patched, patched_loss)}
    images, targets = MH.synthetic_batch(B, 640, "cpu", seed=rank_seed(rank))
    host_images = images.pin_memory()
    host_targets = [{k: v.pin_memory() for k, v in t.items()} for t in targets]
    h2d = host_images.numel() * 4 + sum(v.numel() * v.element_size() for t in host_targets for v in t.values())
    res = {"workload": "dfine_m_train_640_b32 (full model: HGNetv2-B2 + HybridEncoder + DFINETransformer, "
                       "criterion + Hungarian matcher, AdamW, clip 0.1, bf16 autocast)",
           "images_per_gpu": B, "patched_modules": counts, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
    for name, (m, crit) in arms.items():
        opt = MH.build_optimizer(m, "m")
        run = m
        if dist_on:
            from torch.nn.parallel import DistributedDataParallel as DDP
            run = DDP(m, device_ids=[device.index], output_device=device.index, find_unused_parameters=False)
        loss_host = torch.zeros(1).pin_memory()

        def step(run=run, opt=opt, crit=crit):
            img = host_images.to(device, non_blocking=True)
            tg = [{k: v.to(device, non_blocking=True) for k, v in t.items()} for t in host_targets]
            _, _, loss = MH.train_step(run, crit, img, tg, torch.bfloat16, optimizer=opt)
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)

        n0 = ops.LAUNCHES["count"]
        ms = timed(step, steps, warmup)
        res[name] = {"ms_per_step": ms, "imgs_per_s": B * world / (ms / 1e3),
                     "dfine_b200_launches_per_step": (ops.LAUNCHES["count"] - n0) // (steps + warmup),
                     "loss_after": float(loss_host)}
        if name == "patched":
            ops.enable_kernel_timers(True)
            step()
            torch.cuda.synchronize(device)
            tm = ops.kernel_timers()
            hot = {k: sum(s.elapsed_time(e) for s, e in v) for k, v in tm.items()}
            ops.enable_kernel_timers(False)
            res[name]["hot_path_kernel_ms_per_step"] = sum(hot.values())
            res[name]["hot_path_kernels_ms"] = hot
        del opt, run
    res["speedup"] = res["patched"]["imgs_per_s"] / res["reference"]["imgs_per_s"]
    out["train_config3"] = res
    del model, patched, arms, patched_loss
    torch.cuda.empty_cache()

    # ---- config 4: D-FINE-m instance segmentation training, 640x640, batch 16 per GPU: reference vs
    #      patch_model(mask="matched") + patch_criterion (matched-rows-only mask assembly, fused focal + dice) ----
    try:
        out["train_config4_seg"] = _seg_train_leg(device, rank, world, dist_on, max(3, steps // 2), warmup, timed)
    except Exception as exc:  # noqa: BLE001
        out["train_config4_seg"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    torch.cuda.empty_cache()

    # ---- config 2: D-FINE-s inference, 640x640, batch 64, bf16 autocast (rank 0 only does not matter:
    #      inference needs no collective, every rank runs its own batch) ----
    B = 64
    model, _ = MH.build("s", device, 640, False, seed=0)
    model.eval()
    patched = copy.deepcopy(model)
    dfine_b200.patch_model(patched, layer=True)
    images, _ = MH.synthetic_batch(B, 640, "cpu", seed=rank_seed(rank))
    host_images = images.pin_memory()
    res = {"workload": "dfine_s_infer_640_b64 (full model, eval, no_grad, bf16 autocast)", "images_per_gpu": B,
           "h2d_bytes_per_step": host_images.numel() * 4}
    for name, m in {"reference": model, "patched": patched}.items():
        box_host = torch.empty(B, 300, 4).pin_memory()
        logit_host = torch.empty(B, 300, 80).pin_memory()

        def step(m=m):
            o = MH.infer_step(m, host_images.to(device, non_blocking=True), torch.bfloat16)
            box_host.copy_(o["pred_boxes"], non_blocking=True)
            logit_host.copy_(o["pred_logits"], non_blocking=True)

        ms = timed(step, steps, warmup)
        res[name] = {"ms_per_step": ms, "imgs_per_s": B * world / (ms / 1e3)}
    res["d2h_bytes_per_step"] = (box_host.numel() + logit_host.numel()) * 4
    res["speedup"] = res["patched"]["imgs_per_s"] / res["reference"]["imgs_per_s"]
    out["infer_config2"] = res
    # the patched model replayed from a CUDA graph (dfine_b200.GraphedInference: same kernels, one launch)
    try:
        graphed = dfine_b200.GraphedInference(patched, host_images.to(device), amp_dtype=torch.bfloat16)

        def gstep():
            o = graphed(host_images.to(device, non_blocking=True))
            box_host.copy_(o["pred_boxes"], non_blocking=True)
            logit_host.copy_(o["pred_logits"], non_blocking=True)

        ms = timed(gstep, steps, warmup)
        res["patched_graphed"] = {"ms_per_step": ms, "imgs_per_s": B * world / (ms / 1e3)}
        res["speedup_graphed"] = res["patched_graphed"]["imgs_per_s"] / res["reference"]["imgs_per_s"]
        del graphed
    except Exception as exc:  # noqa: BLE001
        res["patched_graphed"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    del model, patched
    torch.cuda.empty_cache()

    # ---- config 1's model and shape on this GPU: D-FINE-n, batch 1 (latency; the reference's own config-1 run is
    #      the CPU one, see cpu_baseline / BASELINE.md) ----
    try:
        out["infer_config1_shape_gpu"] = _latency_leg(device, rank, world, steps, warmup, timed)
    except Exception as exc:  # noqa: BLE001
        out["infer_config1_shape_gpu"] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    return out


def _latency_leg(device, rank, world, steps, warmup, timed):
    import copy

    import dfine_b200
    from baseline import model_harness as MH
    model, _ = MH.build("n", device, 640, False, seed=0)
    model.eval()
    patched = copy.deepcopy(model)
    dfine_b200.patch_model(patched, layer=True)
    images, _ = MH.synthetic_batch(1, 640, "cpu", seed=rank_seed(rank))
    host_images = images.pin_memory()
    res = {"workload": "dfine_n_infer_640_b1 (full model, eval, no_grad, bf16 autocast; latency)", "images_per_gpu": 1,
           "h2d_bytes_per_step": host_images.numel() * 4}
    box_host = torch.empty(1, 300, 4).pin_memory()
    graphed = dfine_b200.GraphedInference(patched, host_images.to(device), amp_dtype=torch.bfloat16)
    for name, fn in {"reference": lambda x: MH.infer_step(model, x, torch.bfloat16),
                     "patched": lambda x: MH.infer_step(patched, x, torch.bfloat16),
                     "patched_graphed": graphed}.items():
        def step(fn=fn):
            o = fn(host_images.to(device, non_blocking=True))
            box_host.copy_(o["pred_boxes"], non_blocking=True)
            torch.cuda.current_stream(device).synchronize()     # latency: the result is on the host

        ms = timed(step, max(steps, 20), warmup)
        res[name] = {"ms_per_image": ms, "imgs_per_s": world / (ms / 1e3)}
    res["speedup"] = res["reference"]["ms_per_image"] / res["patched"]["ms_per_image"]
    res["speedup_graphed"] = res["reference"]["ms_per_image"] / res["patched_graphed"]["ms_per_image"]
    return res


def _seg_train_leg(device, rank, world, dist_on, steps, warmup, timed):
    import copy

    import dfine_b200
    from baseline import model_harness as MH
    from dfine_b200 import ops
    B = 16
    model, loss_fn = MH.build("m", device, 640, True, seed=0)
    model.train(), loss_fn.train()
    patched, ploss = copy.deepcopy(model), copy.deepcopy(loss_fn)
    counts = dfine_b200.patch_model(patched, mask="matched", layer=True)
    counts.update(dfine_b200.patch_criterion(ploss))
    images, targets = MH.synthetic_batch(B, 640, "cpu", seed=rank_seed(rank), seg=True)
    host_images = images.pin_memory()
    host_targets = [{k: v.pin_memory() for k, v in t.items()} for t in targets]
    h2d = host_images.numel() * 4 + sum(v.numel() * v.element_size() for t in host_targets for v in t.values())
    res = {"workload": "dfine_m_seg_train_640_b16 (full model + MaskPixelDecoder + mask head, criterion with mask "
                       "losses, AdamW, clip 0.1, bf16 autocast)", "images_per_gpu": B, "patched_modules": counts,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}
    for name, (m, crit) in {"reference": (model, loss_fn), "patched": (patched, ploss)}.items():
        opt = MH.build_optimizer(m, "m")
        run = m
        if dist_on:
            from torch.nn.parallel import DistributedDataParallel as DDP
            run = DDP(m, device_ids=[device.index], output_device=device.index, find_unused_parameters=False)
        loss_host = torch.zeros(1).pin_memory()

        def step(run=run, opt=opt, crit=crit):
            img = host_images.to(device, non_blocking=True)
            tg = [{k: v.to(device, non_blocking=True) for k, v in t.items()} for t in host_targets]
            _, _, loss = MH.train_step(run, crit, img, tg, torch.bfloat16, optimizer=opt)
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)

        torch.cuda.reset_peak_memory_stats(device)
        n0 = ops.LAUNCHES["count"]
        ms = timed(step, steps, warmup)
        res[name] = {"ms_per_step": ms, "imgs_per_s": B * world / (ms / 1e3),
                     "dfine_b200_launches_per_step": (ops.LAUNCHES["count"] - n0) // (steps + warmup),
                     "loss_after": float(loss_host),
                     "peak_memory_gb": torch.cuda.max_memory_allocated(device) / 2 ** 30}
        del opt, run
    res["speedup"] = res["patched"]["imgs_per_s"] / res["reference"]["imgs_per_s"]
    return res


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(gpu_index), "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def load_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (or None)."""
    for name in ("r5_traffic.json", "r3_traffic.json", "r2_traffic.json"):   # newest capture first
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f).get(kernel)
        except (OSError, ValueError):
            continue
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def rank_seed(rank: int) -> int:
    """Per-rank synthetic batch (the reference seeds 42 + rank, src/dl/train.py:131)."""
    return 42 + rank


def max_over_ranks(ms: float, device, dist_on: bool) -> float:
    """Step time of the job = the slowest rank's device time."""
    if not dist_on:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def job_throughput(images_per_gpu: int, world: int, steps: int, ms: float) -> float:
    """Whole-job images/s: every rank processes its own shard of the batch (weak scaling)."""
    return images_per_gpu * world * steps / (ms / 1e3)


def time_steps(fn, steps: int, warmup: int, device, dist_on: bool):
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(device)
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(device)
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    torch.cuda.synchronize(device)
    if dist_on:
        dist.barrier()
    return max_over_ranks(s.elapsed_time(e), device, dist_on)


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's torch path (oracle/torch_port.py), bounded sample
# ------------------------------------------------------------------------------------------
def _reference_cpu_step(wl: dict, inp: dict):
    """One hot-path step through the UNMODIFIED reference (baseline/_ref: the reference's own
    MSDeformableAttention, TransformerDecoder.value_op, weighting_function, Integral, distance2bbox --
    src/d_fine/arch/dfine_decoder.py, utils.py) on the host CPU, float32: per decoder layer the module forward +
    the FDR decode, then the backward of all of it -- the same work the CUDA arm's step does.  Returns a
    callable, or None when the reference package was not installed by build()."""
    import types
    try:
        from baseline import ref_install
        if not ref_install.installed():
            return None
        ref_install.import_reference()
        from src.d_fine.arch import dfine_decoder as dd
        from src.d_fine.arch import utils as au
    except Exception:  # noqa: BLE001
        return None
    H = wl["H"]
    mods = []
    for lw in inp["lin"]:
        m = dd.MSDeformableAttention(wl["C"], H, len(wl["shapes"]), list(wl["npts"]))
        with torch.no_grad():
            m.sampling_offsets.weight.copy_(lw["so_w"])
            m.attention_weights.weight.copy_(lw["aw_w"])
            m.attention_weights.bias.copy_(lw["aw_b"])
        mods.append(m)
    integral = dd.Integral(wl["reg_max"])
    up, rs = torch.tensor([wl["up"]]), torch.tensor([wl["reg_scale"]])
    stub = types.SimpleNamespace(num_head=H)

    def step():
        for m in mods:
            m.zero_grad(set_to_none=True)
        mem = inp["memory"].detach().requires_grad_(True)
        project = au.weighting_function(wl["reg_max"], up, rs)
        value = dd.TransformerDecoder.value_op(stub, mem, None, None, None, wl["shapes"])
        outs, gos = [], []
        for i, m in enumerate(mods):
            q = inp["queries"][i].detach().requires_grad_(True)
            pc = inp["corners"][i].detach().requires_grad_(True)
            outs.append(m(q, inp["refs"][i], value, wl["shapes"]))
            outs.append(au.distance2bbox(inp["ref_init"], integral(pc, project), rs))
            gos += [inp["grad_outs"][i], inp["grad_boxes"][i]]
        torch.autograd.backward(outs, gos)
    return step


def cpu_reference_run(wl: dict, sample_b: int, steps: int, warmup: int, seed: int = 42):
    """Times the path's CPU implementation on the host cores: the unmodified reference when baseline/_ref is
    present (kind "reference"), else its restatement oracle/torch_port.py (kind "port").  Returns
    (imgs/s, ms per step, kind)."""
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1)
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    inp = make_inputs(wl, sample_b, seed, "cpu")
    kind = "reference"
    step = _reference_cpu_step(wl, inp)
    if step is None:
        kind = "port"
        from oracle import torch_port as TP  # the only use of oracle/ in this file
        H, P = wl["H"], sum(wl["npts"])
        nps = torch.tensor([1.0 / n for n in wl["npts"] for _ in range(n)])
        ang = torch.arange(H, dtype=torch.float32) * (2.0 * torch.pi / H)
        dirs = torch.stack([ang.cos(), ang.sin()], -1)
        dirs = dirs / dirs.abs().max(-1, keepdim=True).values
        rank = torch.cat([torch.arange(1, n + 1) for n in wl["npts"]]).float()
        so_b = (dirs[:, None, :] * rank[None, :, None]).reshape(-1)
        up, rs = torch.tensor([wl["up"]]), torch.tensor([wl["reg_scale"]])
        lins = [(*(t.detach().requires_grad_(True) for t in (lw["so_w"], so_b, lw["aw_w"], lw["aw_b"])), nps)
                for lw in inp["lin"]]

        def step():
            for t in (x for lin in lins for x in lin[:4]):
                t.grad = None
            TP.hot_path_step(inp["memory"], inp["queries"], inp["refs"], lins, wl["shapes"], wl["npts"], H,
                             inp["corners"], inp["ref_init"], up, rs, inp["grad_outs"], inp["grad_boxes"],
                             train=True)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps * 1e3, kind


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_b = args.cpu_sample or wl["B"]
    v, ms, kind = cpu_reference_run(wl, sample_b, args.steps, args.warmup)
    cores = torch.get_num_threads()
    what = ("the unmodified reference modules (baseline/_ref/src/d_fine/arch)" if kind == "reference"
            else "reference call sequence restated in oracle/torch_port.py")
    sample = (f"{sample_b} of {wl['B']} images per step, {args.steps} steps, fp32, torch "
              f"{torch.__version__} CPU, {what}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "images_per_gpu": wl["B"], "queries": wl["Lq"],
                   "levels": wl["shapes"], "points": wl["npts"], "decoder_layers": wl["layers"],
                   "sample_images_per_step": sample_b, "device": "cpu", "value_dtype": "f32"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="images per CPU-baseline step (0 = the workload's full batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager-PyTorch GPU bar")
    ap.add_argument("--no-full-model", action="store_true",
                    help="skip the full-model legs (reference model patched / unpatched on this GPU)")
    ap.add_argument("--full-model-steps", type=int, default=8)
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the kernel-level legs at the other BASELINE configs")
    ap.add_argument("--grad-sync", default="nvls", choices=["nvls", "overlap", "eager", "off"],
                    help="N > 1: NCCL all-reduce of the path's Linear gradients after every step / no exchange")
    ap.add_argument("--launch-list", action="store_true",
                    help="profiling aid: run W+K eager steps of the B200 path and exit (for an ncu launch list)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import dfine_b200
    from dfine_b200 import ops
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist_on = world > 1
    if dist_on:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    dfine_b200._lib.lib()  # fail loudly if the extension is missing

    inp = make_inputs(wl, wl["B"], rank_seed(rank), "cpu")
    hp = HotPath(wl, inp, device, world=world, sync_mode=args.grad_sync)
    hp.load_device(inp)
    hp.pin_host(inp)

    if args.launch_list:
        for _ in range(args.warmup + args.steps):
            hp.step(hp.d)
        torch.cuda.synchronize(device)
        print(json.dumps({"launch_list": True, "steps": args.warmup + args.steps}))
        return

    exchange_err = hp.check_exchange() if dist_on else None
    sampler = ClockSampler(local) if rank == 0 else None
    # ---- headline: device-resident step, replayed from a CUDA graph ----
    ops.enable_kernel_timers(False)
    hp.capture(warmup=max(3, args.warmup))
    ms = time_steps(hp.replay, args.steps, args.warmup, device, dist_on)
    launches = hp.launches_per_step * args.steps
    # ---- end to end: host buffers, H2D + D2H inside the timed region, same graph ----
    ms_e2e_serial = time_steps(hp.step_e2e, args.steps, args.warmup, device, dist_on)
    # same, double-buffered: H2D of step i+1 on a copy stream while step i computes
    hp.setup_pipeline(inp)
    ms_e2e = time_steps(hp.step_e2e_pipelined, args.steps, args.warmup, device, dist_on)
    hp.drain_pipeline()
    # the box's host -> device ceiling with all ranks copying at once: the same slab, copies back to back
    ms_h2d = time_steps(lambda: hp.d_slab.copy_(hp.h_slab, non_blocking=True), args.steps, 2, device, dist_on)
    h2d_ceiling = hp.h2d_bytes * args.steps / (ms_h2d / 1e3) / 1e9
    # ---- eager pass (no graph) with CUDA-event brackets around every C-ABI launch: the
    #      per-kernel durations behind the roofline figures ----
    for _ in range(2):
        hp.step_eager(hp.d)
    ops.enable_kernel_timers(True)
    ms_eager = time_steps(lambda: hp.step_eager(hp.d), args.steps, 0, device, dist_on)
    timers = ops.kernel_timers()
    kernel_ms = {k: sum(s.elapsed_time(e) for s, e in v) / len(v) for k, v in timers.items()}
    kernel_calls = {k: len(v) // args.steps for k, v in timers.items()}
    ops.enable_kernel_timers(False)
    # ---- kernel leg: each kernel of a layer launched back to back (device-bound timing) ----
    kms = kernel_leg(wl, hp, device)
    full_model = None
    if not args.no_full_model:
        try:
            full_model = full_model_leg(device, rank, world, dist_on, args.full_model_steps, 3)
        except Exception as exc:  # noqa: BLE001  (must not take the bench line down)
            import traceback
            full_model = {"error": f"{type(exc).__name__}: {exc}"[:300], "where": traceback.format_exc()[-400:]}
            if dist_on:   # keep the ranks in step for the teardown
                torch.cuda.synchronize(device)
    clocks = sampler.stop() if sampler else None
    other = None
    if rank == 0 and world == 1 and not args.no_secondary:
        try:
            other = secondary_kernel_legs(device, load_peaks()[0])
            other["config2_inference_step"] = inference_leg(device)
        except Exception as exc:  # noqa: BLE001  (informational legs must not take the bench line down)
            other = {"error": str(exc)[:200]}

    value = job_throughput(wl["B"], world, args.steps, ms)
    e2e = job_throughput(wl["B"], world, args.steps, ms_e2e)

    if rank != 0:
        if dist_on:
            hp.close()
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    fwd_b, bwd_b = algorithmic_bytes(wl, wl["B"])
    dom = "msda_bwd" if kms["msda_bwd"] >= kms["msda_fwd"] else "msda_fwd"
    dom_bytes = bwd_b if dom == "msda_bwd" else fwd_b
    achieved = dom_bytes / (kms[dom] / 1e3) / 1e9
    fb_ms = kms["msda_fwd"] + kms["msda_bwd"]
    fwd4, bwd4 = algorithmic_bytes(wl, wl["B"], e_g=4)
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": load_traffic(dom), "peak_source": peak_src,
        "traffic_source": "profiles/r5_traffic.json (ncu --set full, dram read + write per launch)",
        "algorithmic_bytes_per_launch": dom_bytes, "avg_launch_ms": kms[dom],
        "msda_fwd_bwd": {"algorithmic_bytes": fwd_b + bwd_b, "ms": fb_ms,
                         "achieved": (fwd_b + bwd_b) / (fb_ms / 1e3) / 1e9,
                         "frac": (fwd_b + bwd_b) / (fb_ms / 1e3) / 1e9 / peak,
                         "frac_with_survey_bytes_e_g4": (fwd4 + bwd4) / (fb_ms / 1e3) / 1e9 / peak,
                         "note": "per decoder layer: dfine_msda_fwd + dfine_msda_bwd (dots kernel + "
                                 "atomic-free grad_value kernel writing every row of a bf16 gradient: "
                                 "the first layer of a backward pass; the other layers run in accumulate "
                                 "mode, msda_bwd_accumulate)"},
        "kernel_ms": kms,
        "eager_kernel_ms": kernel_ms, "kernel_calls_per_step": kernel_calls,
        "measured_in": "kernel leg after the timed region: 20 back-to-back launches of each C-ABI "
                       "kernel between two CUDA events on the launching stream (kernel_ms); "
                       "eager_kernel_ms brackets every launch of an eager step and includes host gaps; "
                       "the timed region itself replays a CUDA graph",
    }

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "eager_ms_per_step": ms_eager / args.steps,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": args.workload, "images_per_gpu": wl["B"], "queries": wl["Lq"],
                   "levels": wl["shapes"], "points": wl["npts"], "decoder_layers": wl["layers"],
                   "value_dtype": "bf16", "accumulate": "f32", "cuda_graph": True,
                   "memory_grad": "one buffer shared by the 4 layers (hub node), layers 2-4 accumulate in the kernel",
                   "l2_policy": "inputs_larger_than_l2 (per layer: memory 137.6 MB + grad 137.6 MB + queries, "
                                "grads, records; 4 layers per step)",
                   "parallelism": f"dp{world} (batch-sharded; the kernels need no collective)",
                   "collective": (None if hp.bucket is None else
                                  f"NCCL all-reduce (avg) of the {wl['layers']} layers' Linear gradients, "
                                  f"{hp.bucket.nbytes} B per step, " +
                                  ("one bucket per layer, each forked onto a side stream when the layer's "
                                   "weight-gradient kernel is done and captured inside the step's CUDA graph "
                                   "(overlaps the backward of the remaining layers)" if hp.sync_mode == "overlap"
                                   else "one bucket, enqueued on the compute stream after every graph replay")
                                  if hp.nvls is None else
                                  f"none launched: dfine_multicast_add adds every layer's dW / db, scaled by "
                                  f"1/{world}, into all ranks' replicas of a symmetric buffer with multimem.red "
                                  f"(NVLS: summed in the switch); {hp.nvls.nbytes} B per step, two device-side "
                                  f"barriers per step, all inside the CUDA graph"),
                   "collective_check_max_rel_err_vs_nccl": exchange_err,
                   "collective_fallback": hp.nvls_error},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": hp.h2d_bytes, "d2h_bytes_per_step": hp.d2h_bytes,
                "h2d_copies_per_step": hp.h2d_copies,
                "h2d_gbs_per_gpu": hp.h2d_bytes * args.steps / (ms_e2e / 1e3) / 1e9,
                "h2d_ceiling_gbs_per_gpu": h2d_ceiling,
                "h2d_ceiling_note": f"the same pinned slab copied back to back on all {world} ranks at once "
                                    "(max over ranks): what the host side of this box delivers per GPU",
                "pipeline": "double-buffered: H2D of step i+1 on a copy stream overlaps the graph "
                            "replay of step i; every step still copies all its inputs from pinned host "
                            "memory and reads its boxes back",
                "serial_value": job_throughput(wl["B"], world, args.steps, ms_e2e_serial),
                "serial_ms_per_step": ms_e2e_serial / args.steps},
        "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
    }
    if full_model is not None:   # the real model, patched vs unpatched, on this GPU
        out["full_model"] = full_model
    if other is not None:   # kernel-level legs at configs 2 / 4 / 5 and the config-2 inference step
        out["other_configs_kernel_level"] = other

    if world == 1 and not args.no_eager:
        # informational: the reference's eager PyTorch path on the same GPU (fp32 restatement
        # under bf16 autocast), i.e. the bar a user of the reference sees today
        try:
            out["gpu_eager_reference"] = eager_gpu_bar(wl, inp, device, max(3, args.steps // 4))
        except Exception as exc:  # noqa: BLE001
            out["gpu_eager_reference"] = {"error": str(exc)[:200]}
    if world == 1 and not args.no_cpu_baseline:
        nb = args.cpu_sample or wl["B"]
        v, cms, kind = cpu_reference_run(wl, nb, 5, 1)
        cores = torch.get_num_threads()
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                               "ms_per_step": cms,
                               "sample": f"{nb} of {wl['B']} images per step, 5 timed steps after 1 warm-up, fp32, "
                                         + ("the unmodified reference modules (baseline/_ref) on the host CPU"
                                            if kind == "reference" else "oracle/torch_port.py on the host CPU")}
    print(json.dumps(out))
    if dist_on:
        hp.close()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def eager_gpu_bar(wl, inp, device, steps):
    from oracle import torch_port as TP  # baseline leg
    H = wl["H"]
    nps = torch.tensor([1.0 / n for n in wl["npts"] for _ in range(n)], device=device)
    import dfine_b200
    ref_mod = dfine_b200.MSDeformableAttention(wl["C"], H, len(wl["shapes"]), wl["npts"])
    so_b = ref_mod.sampling_offsets.bias.detach().to(device)
    up = torch.tensor([wl["up"]], device=device)
    rs = torch.tensor([wl["reg_scale"]], device=device)
    mem = inp["memory"].to(device, torch.bfloat16)
    d = dict(q=[q.to(device) for q in inp["queries"]], r=[r.to(device) for r in inp["refs"]],
             c=[c.to(device, torch.bfloat16) for c in inp["corners"]], ri=inp["ref_init"].to(device),
             go=[g.to(device) for g in inp["grad_outs"]], gb=[g.to(device) for g in inp["grad_boxes"]])
    lins = [(*(t.to(device).requires_grad_(True) for t in (lw["so_w"], so_b, lw["aw_w"], lw["aw_b"])), nps)
            for lw in inp["lin"]]

    def step():
        for t in (x for lin in lins for x in lin[:4]):
            t.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            TP.hot_path_step(mem, d["q"], d["r"], lins, wl["shapes"], wl["npts"], H, d["c"], d["ri"],
                             up, rs, d["go"], d["gb"], train=True)

    ms = time_steps(step, steps, 3, device, False)
    return {"value": wl["B"] * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps,
            "what": "reference call sequence (F.grid_sample path) eager on the same GPU, bf16 autocast"}


if __name__ == "__main__":
    main()
