"""GPU parity tests (run with `-m gpu` on a B200): the CUDA kernels, called through the
C-ABI of libdfine_b200.so, against (a) the golden vectors produced by the unmodified
reference, (b) the CPU oracle on seeded inputs, and (c) size-independent properties at the
BASELINE.json sizes.  Tolerances are the ones north_star states: corner indices bit-exact,
fp32 1e-5 relative, bf16 1e-2 relative (relative to the tensor's max magnitude).
"""
import os

import numpy as np
import pytest
import torch

from util import (BF16_RTOL, FP32_RTOL, assert_close, assert_close_elementwise, bf16_bits_to_f32, elementwise_err,
                  golden, rel_err)

pytestmark = pytest.mark.gpu

CORE_CASES = ["core_m_small", "core_n_small", "core_x444", "core_edge", "core_near_centre"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    return torch.device("cuda:0")


def _t(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t.to(dtype) if dtype is not None else t


def _core_case(name, dev, vdtype):
    import dfine_b200.ops as ops
    g = golden(name)
    H, c = int(g["H"]), int(g["c"])
    mem = _t(bf16_bits_to_f32(g["memory_bf16"]), dev, vdtype)
    go = _t(bf16_bits_to_f32(g["grad_out_bf16"]), dev)
    spec = ops.level_spec(g["shapes"].tolist(), g["npts"].tolist())
    return g, ops, spec, H, mem, _t(g["loc"], dev), _t(g["attn"], dev), go


@pytest.mark.parametrize("vdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", CORE_CASES)
def test_core_forward_golden_and_indices(name, vdtype, dev):
    from oracle import cpu_oracle as O
    g, ops, spec, H, mem, loc, attn, _ = _core_case(name, dev, vdtype)
    out, idx = ops.msda_forward_raw(mem, spec, H, loc, attn, None, None, 0.5, False,
                                    torch.float32, want_idx=True)
    # inputs are bf16-representable, accumulation is fp32 in both modes: same tolerance
    assert_close(out.cpu().numpy(), g["out"], FP32_RTOL, "out")
    B, L, C = mem.shape
    _, o_idx, _ = O.msda_fwd(mem.float().cpu().numpy().reshape(B, L, H, C // H), g["shapes"],
                             g["npts"], g["loc"], g["attn"], want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), o_idx), "corner indices / level offsets not bit-exact"


@pytest.mark.parametrize("atomic", [False, True], ids=["gather", "atomic"])
@pytest.mark.parametrize("vdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", CORE_CASES)
def test_core_backward_golden(name, vdtype, atomic, dev):
    """Both grad_value strategies: the atomic-free pixel-CSR gather (default) and the fp32
    vector-reduction fallback."""
    g, ops, spec, H, mem, loc, attn, go = _core_case(name, dev, vdtype)
    gm, gl, ga = ops.msda_backward_raw(mem, spec, H, loc, attn, None, None, 0.5, False, go,
                                       force_atomic=atomic)
    assert_close(gm.cpu().numpy(), g["grad_memory"], FP32_RTOL, "grad_value")
    assert_close(gl.cpu().numpy(), g["grad_loc"], FP32_RTOL, "grad_loc")
    assert_close(ga.cpu().numpy(), g["grad_attn"], FP32_RTOL, "grad_attn")
    # bf16 grad_out (pure-bf16 models)
    gm2, gl2, ga2 = ops.msda_backward_raw(mem, spec, H, loc, attn, None, None, 0.5, False,
                                          go.to(torch.bfloat16), force_atomic=atomic)
    assert_close(gm2.cpu().numpy(), g["grad_memory"], FP32_RTOL, "grad_value (bf16 go)")
    assert_close(ga2.cpu().numpy(), g["grad_attn"], FP32_RTOL, "grad_attn (bf16 go)")
    # bf16 grad_value written directly (AMP): one rounding of the fp32 sum
    gm3, _, _ = ops.msda_backward_raw(mem, spec, H, loc, attn, None, None, 0.5, False, go,
                                      gv_dtype=torch.bfloat16, force_atomic=atomic)
    assert gm3.dtype == torch.bfloat16
    assert_close(gm3.float().cpu().numpy(), g["grad_memory"], BF16_RTOL, "grad_value (bf16 out)")
    want = torch.from_numpy(g["grad_memory"]).to(torch.bfloat16).float().numpy()
    fin = np.isfinite(want)
    assert np.abs(gm3.float().cpu().numpy()[fin] - want[fin]).max() <= 2 ** -7 * np.abs(want[fin]).max()


@pytest.mark.parametrize("atomic", [False, True], ids=["gather", "atomic"])
@pytest.mark.parametrize("gdtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", CORE_CASES)
def test_core_backward_accumulate(name, gdtype, atomic, dev):
    """DFINE_MSDA_GRAD_VALUE_ACCUMULATE: grad_value is added to the caller's running gradient
    (what autograd's accumulation over the decoder layers does, dfine_decoder.py:470-515);
    rows no sample touches keep their bits."""
    if atomic and gdtype == torch.bfloat16:
        pytest.skip("the vector-reduction fallback accumulates in float32 only")
    g, ops, spec, H, mem, loc, attn, go = _core_case(name, dev, torch.float32)
    want = g["grad_memory"]
    fin = np.isfinite(want)
    rng = np.random.default_rng(7)
    base = rng.standard_normal(want.shape).astype(np.float32) * max(float(np.abs(want[fin]).max()), 1e-3)
    run = torch.from_numpy(base).to(dev).to(gdtype).contiguous()
    base_q = run.float().cpu().numpy()
    gm, _, _ = ops.msda_backward_raw(mem, spec, H, loc, attn, None, None, 0.5, False, go,
                                     force_atomic=atomic, accumulate_into=run)
    assert gm.data_ptr() == run.data_ptr()
    got = gm.float().cpu().numpy()
    tol = FP32_RTOL if gdtype == torch.float32 else BF16_RTOL
    assert_close(got, base_q + want, tol, "accumulated grad_value")
    untouched = fin & (want == 0)
    assert np.array_equal(got[untouched], base_q[untouched]), "untouched rows must keep their bits"


def test_core_autograd_through_value_views(dev):
    """The drop-in core called exactly like the reference does: value = tuple of strided
    views from value_op; gradients must reach `memory` through the zero-copy route."""
    import dfine_b200
    from oracle import torch_port as TP
    g = golden("core_m_small")
    H = int(g["H"])
    mem = _t(bf16_bits_to_f32(g["memory_bf16"]), dev).requires_grad_(True)
    loc = _t(g["loc"], dev).requires_grad_(True)
    attn = _t(g["attn"], dev).requires_grad_(True)
    shapes, npts = g["shapes"].tolist(), g["npts"].tolist()
    value = TP.value_views(mem, H, shapes)
    assert not value[0].is_contiguous()
    out = dfine_b200.msda_core(value, shapes, loc, attn, npts)
    assert out.shape == g["out"].shape and out.dtype == torch.float32
    out.backward(_t(bf16_bits_to_f32(g["grad_out_bf16"]), dev))
    assert_close(out.detach().cpu().numpy(), g["out"], FP32_RTOL, "out")
    assert_close(mem.grad.cpu().numpy(), g["grad_memory"], FP32_RTOL, "grad_memory")
    assert_close(loc.grad.cpu().numpy(), g["grad_loc"], FP32_RTOL, "grad_loc")
    assert_close(attn.grad.cpu().numpy(), g["grad_attn"], FP32_RTOL, "grad_attn")
    # packed (non zero-copy) route: values cloned level by level
    mem2 = mem.detach().clone().requires_grad_(True)
    value2 = tuple(v.clone() for v in TP.value_views(mem2, H, shapes))
    out2 = dfine_b200.msda_core(value2, shapes, loc.detach(), attn.detach(), npts)
    out2.backward(_t(bf16_bits_to_f32(g["grad_out_bf16"]), dev))
    assert_close(mem2.grad.cpu().numpy(), g["grad_memory"], FP32_RTOL, "grad_memory (packed)")


@pytest.mark.parametrize("name", ["module_m_small", "module_n_small"])
def test_module_mirror_matches_reference(name, dev):
    """dfine_b200.MSDeformableAttention loaded from the reference module's state dict."""
    import dfine_b200
    from oracle import torch_port as TP
    g = golden(name)
    H = int(g["H"])
    shapes, npts = g["shapes"].tolist(), g["npts"].tolist()
    mem = _t(bf16_bits_to_f32(g["memory_bf16"]), dev).requires_grad_(True)
    C = mem.shape[-1]
    m = dfine_b200.MSDeformableAttention(C, H, len(shapes), npts).to(dev)
    assert sorted(m.state_dict().keys()) == [str(k) for k in g["state_dict_keys"]]
    m.load_state_dict({"sampling_offsets.weight": _t(g["so_w"], dev),
                       "sampling_offsets.bias": _t(g["so_b"], dev),
                       "attention_weights.weight": _t(g["aw_w"], dev),
                       "attention_weights.bias": _t(g["aw_b"], dev),
                       "num_points_scale": _t(g["num_points_scale"], dev)})
    q = _t(g["query"], dev).requires_grad_(True)
    ref = _t(g["ref_points"], dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    out = m(q, ref, TP.value_views(mem, H, shapes), shapes)
    out.backward(_t(bf16_bits_to_f32(g["grad_out_bf16"]), dev))
    assert_close(out.detach().cpu().numpy(), g["out"], FP32_RTOL, "out")
    assert_close(mem.grad.cpu().numpy(), g["grad_memory"], FP32_RTOL, "grad_memory")
    # gradients that went through cuBLAS GEMMs: summation order differs, 5e-5
    assert_close(q.grad.cpu().numpy(), g["grad_query"], 5e-5, "grad_query")
    assert_close(m.sampling_offsets.weight.grad.cpu().numpy(), g["g_so_w"], 5e-5, "g_so_w")
    assert_close(m.sampling_offsets.bias.grad.cpu().numpy(), g["g_so_b"], 5e-5, "g_so_b")
    assert_close(m.attention_weights.weight.grad.cpu().numpy(), g["g_aw_w"], 5e-5, "g_aw_w")
    assert_close(m.attention_weights.bias.grad.cpu().numpy(), g["g_aw_b"], 5e-5, "g_aw_b")
    with pytest.raises(ValueError):
        m(q, ref[..., :3], TP.value_views(mem, H, shapes), shapes)


@pytest.mark.parametrize("name", ["module_m_small", "module_n_small"])
def test_fused_kernel_from_raw_outputs(name, dev):
    """Fused-input kernels fed with the reference's own raw Linear outputs (fp32), and with
    bf16-rounded raw outputs + bf16 value (the AMP layout)."""
    import dfine_b200.ops as ops
    from oracle import cpu_oracle as O
    g = golden(name)
    H = int(g["H"])
    shapes, npts = g["shapes"].tolist(), g["npts"].tolist()
    spec = ops.level_spec(shapes, npts)
    mem32 = bf16_bits_to_f32(g["memory_bf16"])
    B, L, C = mem32.shape
    go = bf16_bits_to_f32(g["grad_out_bf16"])
    ref = _t(g["ref_points"].reshape(B, -1, 4), dev)
    nps = _t(g["num_points_scale"], dev)
    osc = float(g["offset_scale"])
    # fp32
    out = ops.msda_forward_raw(_t(mem32, dev), spec, H, _t(g["raw_off"], dev), _t(g["raw_logit"], dev),
                               ref, nps, osc, True, torch.float32)
    assert_close(out.cpu().numpy(), g["out"], FP32_RTOL, "fused out")
    gm, gs, ga = ops.msda_backward_raw(_t(mem32, dev), spec, H, _t(g["raw_off"], dev),
                                       _t(g["raw_logit"], dev), ref, nps, osc, True, _t(go, dev))
    o_gv, o_goff, o_glog = O.msda_fused_bwd(mem32.reshape(B, L, H, C // H), shapes, npts, g["raw_off"],
                                            g["raw_logit"], g["ref_points"], g["num_points_scale"], go, osc)
    assert_close(gm.cpu().numpy(), g["grad_memory"], FP32_RTOL, "fused grad_memory")
    assert_close(gs.cpu().numpy(), o_goff, FP32_RTOL, "fused grad raw offsets")
    assert_close(ga.cpu().numpy(), o_glog, FP32_RTOL, "fused grad raw logits")
    # AMP layout: bf16 value, bf16 raw Linear outputs; the oracle gets the same rounded inputs
    off_bf = _t(g["raw_off"], dev, torch.bfloat16)
    log_bf = _t(g["raw_logit"], dev, torch.bfloat16)
    out_bf, idx = ops.msda_forward_raw(_t(mem32, dev, torch.bfloat16), spec, H, off_bf, log_bf, ref,
                                       nps, osc, True, torch.float32, want_idx=True)
    off_r, log_r = off_bf.float().cpu().numpy(), log_bf.float().cpu().numpy()
    loc_r = O.msda_locations(off_r, g["ref_points"], g["num_points_scale"], osc)
    o_out, o_idx, _ = O.msda_fwd(mem32.reshape(B, L, H, C // H), shapes, npts, loc_r,
                                 O.softmax(log_r), want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), o_idx), "AMP-mode corner indices not bit-exact"
    assert_close(out_bf.cpu().numpy(), o_out, FP32_RTOL, "AMP out vs oracle on rounded inputs")
    assert_close(out_bf.cpu().numpy(), g["out"], BF16_RTOL, "AMP out vs fp32 reference")


def test_aten_cuda_probe_indices(dev):
    """Touched-pixel sets of aten::grid_sampler_2d ON THIS GPU vs the kernel's corner dump."""
    import dfine_b200.ops as ops
    g = golden("aten_corner_probe")
    h, w = [int(v) for v in g["hw"]]
    loc = _t(g["loc"], dev)
    n = loc.shape[0]
    inp = torch.ones(n, 1, h, w, device=dev, requires_grad=True)
    out = torch.nn.functional.grid_sample(inp, (2 * loc - 1).reshape(n, 1, 1, 2), mode="bilinear",
                                          padding_mode="zeros", align_corners=False)
    out.sum().backward()
    aten = inp.grad.reshape(n, h * w)
    spec = ops.level_spec([[h, w]], [1])
    value = torch.ones(1, h * w, 1 * 16, device=dev)  # H=1, c=16
    attn = torch.ones(1, n, 1, 1, device=dev)
    o, idx = ops.msda_forward_raw(value, spec, 1, loc.reshape(1, n, 1, 1, 2).contiguous(), attn,
                                  None, None, 0.5, False, torch.float32, want_idx=True)
    idx = idx.reshape(n, 4).cpu().numpy()
    touched = np.zeros((n, h * w), bool)
    for j in range(4):
        ok = idx[:, j] >= 0
        touched[np.nonzero(ok)[0], idx[ok, j]] = True
    aten_nz = aten.cpu().numpy() != 0
    # every pixel ATen touched with non-zero weight must be one of our corners; our extra
    # corners may only be zero-weight ones (invisible to the probe)
    assert not (aten_nz & ~touched).any(), "ATen touched a pixel outside our corner set"
    assert np.array_equal(aten_nz, g["grad_input"] != 0), "ATen CUDA and ATen CPU disagree"
    assert_close(o[0, :, 0].cpu().numpy(), g["out"], FP32_RTOL, "probe out")


def test_aten_cuda_near_centre_cells(dev):
    """Within a few ulps of a pixel centre the cell floor() picks depends on how the unnormalise
    ((g+1)*size-1)/2 is rounded.  ATen's CUDA kernel fuses the multiply-subtract (FFMA); the kernel must
    pick the SAME cell as aten::grid_sampler_2d running on this GPU, at the real level sizes."""
    import dfine_b200.ops as ops
    for S in (80, 40, 20, 128, 64, 32):
        xs = []
        for k in range(S):
            base = np.float32((k + 0.5) / S)
            for d in range(-4, 5):
                v = base
                for _ in range(abs(d)):
                    v = np.nextafter(v, np.float32(2.0 if d > 0 else -2.0), dtype=np.float32)
                xs.append(v)
        lx = torch.tensor(np.asarray(xs, np.float32), device=dev)
        n = lx.numel()
        loc = torch.stack([lx, torch.full_like(lx, 0.5)], -1)
        # ATen on this GPU: d out / d ix over v[x] = (x+1)^2 identifies the cell
        img = ((torch.arange(S, device=dev, dtype=torch.float32) + 1) ** 2).reshape(1, 1, 1, S)
        grid = (2 * loc - 1).reshape(1, 1, n, 2).clone().requires_grad_(True)
        torch.nn.functional.grid_sample(img, grid, mode="bilinear", padding_mode="zeros",
                                        align_corners=False).sum().backward()
        gix = (grid.grad.reshape(n, 2)[:, 0] / (S / 2)).cpu().numpy()
        x0_aten = np.where(gix < 0, S - 1, np.round((gix - 3) / 2)).astype(np.int64)
        spec = ops.level_spec([[1, S]], [1])
        value = torch.ones(1, S, 16, device=dev)
        _, idx = ops.msda_forward_raw(value, spec, 1, loc.reshape(1, n, 1, 1, 2).contiguous(),
                                      torch.ones(1, n, 1, 1, device=dev), None, None, 0.5, False,
                                      torch.float32, want_idx=True)
        idx = idx.reshape(n, 4).cpu().numpy()
        x0 = np.where(idx[:, 0] >= 0, idx[:, 0], idx[:, 1] - 1)
        assert np.array_equal(x0, x0_aten), (S, np.nonzero(x0 != x0_aten)[0][:10])


def test_fdr_golden(dev):
    import dfine_b200
    g = golden("fdr")
    for tag in ("m", "x", "odd"):
        up, rs = [float(v) for v in g[f"up_rs_{tag}"]]
        proj = dfine_b200.fdr_project(torch.tensor([up], device=dev), torch.tensor([rs], device=dev), 32)
        assert_close(proj.cpu().numpy(), g[f"project_{tag}"], FP32_RTOL, f"project {tag}")
    corners = _t(g["corners"], dev).requires_grad_(True)
    ref = _t(g["ref_init"], dev)
    project = _t(g["project"], dev)
    rs = torch.tensor([float(g["reg_scale"])], device=dev)
    dist = dfine_b200.fdr_integral(corners, project, 32)
    assert_close(dist.detach().cpu().numpy(), g["dist"], FP32_RTOL, "integral")
    boxes, dist2 = dfine_b200.fdr_decode(corners, ref, project, rs, 32, return_dist=True)
    assert_close(boxes.detach().cpu().numpy(), g["boxes"], FP32_RTOL, "boxes")
    (boxes * _t(g["grad_boxes"], dev)).sum().backward()
    assert_close(corners.grad.cpu().numpy(), g["grad_corners_from_boxes"], FP32_RTOL, "gc boxes")
    corners.grad = None
    boxes, dist2 = dfine_b200.fdr_decode(corners, ref, project, rs, 32, return_dist=True)
    ((boxes * _t(g["grad_boxes"], dev)).sum() + (dist2 * _t(g["grad_dist"], dev)).sum()).backward()
    assert_close(corners.grad.cpu().numpy(), g["grad_corners_from_both"], FP32_RTOL, "gc both")
    # bf16 logits (AMP): same inputs are bf16-representable
    b16 = dfine_b200.fdr_decode(corners.detach().to(torch.bfloat16), ref, project, rs, 32)
    assert_close(b16.cpu().numpy(), g["boxes"], FP32_RTOL, "boxes from bf16 logits")


def test_msda_vs_oracle_full_resolution(dev):
    """640x640 level shapes (80/40/20), B=2, Lq=300, seeded inputs, against the CPU oracle."""
    import dfine_b200.ops as ops
    from oracle import cpu_oracle as O
    rng = np.random.default_rng(3)
    B, Lq, H, c = 2, 300, 8, 32
    shapes, npts = [[80, 80], [40, 40], [20, 20]], [3, 6, 3]
    spec = ops.level_spec(shapes, npts)
    mem = torch.from_numpy(rng.standard_normal((B, spec.L, H * c), dtype=np.float32)).to(torch.bfloat16)
    raw_off = torch.from_numpy(rng.standard_normal((B, Lq, H, 12, 2), dtype=np.float32) * 2).to(torch.bfloat16)
    raw_log = torch.from_numpy(rng.standard_normal((B, Lq, H, 12), dtype=np.float32)).to(torch.bfloat16)
    ref = np.concatenate([rng.uniform(-0.05, 1.05, (B, Lq, 2)), rng.uniform(0.02, 0.9, (B, Lq, 2))], -1).astype(np.float32)
    nps = np.asarray([1 / 3] * 3 + [1 / 6] * 6 + [1 / 3] * 3, np.float32)
    go = rng.standard_normal((B, Lq, H * c), dtype=np.float32)
    args = (spec, H, raw_off.to(dev), raw_log.to(dev), _t(ref, dev), _t(nps, dev), 0.5, True)
    out, idx = ops.msda_forward_raw(mem.to(dev), *args, torch.float32, want_idx=True)
    gm, gs, ga = ops.msda_backward_raw(mem.to(dev), *args, _t(go, dev))
    m32 = mem.float().numpy().reshape(B, spec.L, H, c)
    loc = O.msda_locations(raw_off.float().numpy(), ref, nps, 0.5)
    attn = O.softmax(raw_log.float().numpy())
    o_out, o_idx, _ = O.msda_fwd(m32, shapes, npts, loc, attn, want_idx=True)
    o_gv, o_gs, o_ga = O.msda_fused_bwd(m32, shapes, npts, raw_off.float().numpy(), raw_log.float().numpy(),
                                        ref, nps, go, 0.5)
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    assert (o_idx < 0).mean() > 0.01, "test should exercise out-of-bounds corners"
    assert_close(out.cpu().numpy(), o_out, FP32_RTOL, "out")
    assert_close(gm.cpu().numpy().reshape(o_gv.shape), o_gv, FP32_RTOL, "grad_value")
    assert_close(gs.cpu().numpy(), o_gs, FP32_RTOL, "grad_offsets")
    assert_close(ga.cpu().numpy(), o_ga, FP32_RTOL, "grad_logits")


@pytest.mark.parametrize("cfg", [
    dict(name="config3 m train", B=32, Lq=500, shapes=[[80, 80], [40, 40], [20, 20]], npts=[3, 6, 3]),
    dict(name="config5 x 1024", B=8, Lq=500, shapes=[[128, 128], [64, 64], [32, 32]], npts=[4, 4, 4]),
    dict(name="config2 s infer", B=64, Lq=300, shapes=[[80, 80], [40, 40], [20, 20]], npts=[3, 6, 3]),
    dict(name="config1 n infer", B=1, Lq=300, shapes=[[40, 40], [20, 20]], npts=[6, 6], c=16),
    dict(name="config4 m seg", B=16, Lq=500, shapes=[[80, 80], [40, 40], [20, 20]], npts=[3, 6, 3]),
], ids=lambda c: c["name"].replace(" ", "_"))
def test_full_size_properties(cfg, dev):
    """BASELINE.json sizes, bf16 value: (1) against eager PyTorch (F.grid_sample restatement)
    on the same GPU, (2) the adjoint identity <out(V), G> == <V, grad_value(G)> which holds
    for any size because the op is linear in V, (3) attention weights sum to one => sampling
    a constant map returns the constant wherever all corners are in bounds."""
    import dfine_b200.ops as ops
    from oracle import torch_port as TP
    torch.manual_seed(0)
    B, Lq, H, c = cfg["B"], cfg["Lq"], 8, cfg.get("c", 32)
    shapes, npts = cfg["shapes"], cfg["npts"]
    spec = ops.level_spec(shapes, npts)
    P = spec.P
    mem = torch.randn(B, spec.L, H * c, device=dev).to(torch.bfloat16)
    ref = torch.cat([torch.rand(B, Lq, 2, device=dev) * 0.9 + 0.05,
                     torch.rand(B, Lq, 2, device=dev) * 0.4 + 0.02], -1)
    raw_off = (torch.randn(B, Lq, H, P, 2, device=dev)).to(torch.bfloat16)
    raw_log = torch.randn(B, Lq, H, P, device=dev).to(torch.bfloat16)
    nps = torch.tensor([1.0 / n for n in npts for _ in range(n)], device=dev)
    G = torch.randn(B, Lq, H * c, device=dev)
    out = ops.msda_forward_raw(mem, spec, H, raw_off, raw_log, ref, nps, 0.5, True, torch.float32)
    gm, gs, ga = ops.msda_backward_raw(mem, spec, H, raw_off, raw_log, ref, nps, 0.5, True, G)
    # (1) eager reference on the GPU (fp32 math on the same bf16-rounded inputs)
    m32 = mem.float().requires_grad_(True)
    ro, rl = raw_off.float().requires_grad_(True), raw_log.float().requires_grad_(True)
    want = TP.msda_from_raw(ro, rl, ref.unsqueeze(2), TP.value_views(m32, H, shapes), shapes, nps, npts)
    want.backward(G)
    assert rel_err(out.cpu().numpy(), want.detach().cpu().numpy()) <= FP32_RTOL
    assert rel_err(gm.cpu().numpy(), m32.grad.cpu().numpy()) <= FP32_RTOL
    assert rel_err(gs.cpu().numpy(), ro.grad.cpu().numpy()) <= 2e-5
    assert rel_err(ga.cpu().numpy(), rl.grad.cpu().numpy()) <= 2e-5
    # (2) adjoint identity in float64
    lhs = (out.double() * G.double()).sum().item()
    rhs = (mem.double() * gm.double()).sum().item()
    scale = (out.double() * G.double()).abs().sum().item()  # fp32 rounding scales with sum |terms|
    assert abs(lhs - rhs) <= 1e-6 * scale, (lhs, rhs, scale)
    # (3) constant map
    ones = torch.ones_like(mem)
    o1, idx = ops.msda_forward_raw(ones, spec, H, raw_off, raw_log, ref, nps, 0.5, True,
                                   torch.float32, want_idx=True)
    inb = (idx >= 0).all(-1).all(-1)  # [B, Lq, H]: every corner of every point in bounds
    assert inb.float().mean() > 0.3
    got = o1.reshape(B, Lq, H, c)[inb]
    assert (got - 1.0).abs().max().item() <= 1e-5


def test_mask_gemm(dev):
    """tcgen05 GEMM vs golden (reference einsum on bf16-representable inputs) and vs
    torch.bmm at the config-4 shape; bf16 products are exact in fp32, so only the
    accumulation order differs."""
    import dfine_b200
    import dfine_b200.ops as ops
    g = golden("mask")
    coef, proto = _t(g["coef"], dev), _t(g["proto"], dev)
    out = dfine_b200.mask_logits(coef, proto, out_dtype=torch.float32)
    assert_close(out.cpu().numpy(), g["logits"], 2e-5, "mask logits (fp32 out)")
    out_b = dfine_b200.mask_logits(coef, proto)
    assert out_b.dtype == torch.bfloat16
    assert_close(out_b.float().cpu().numpy(), g["logits"], BF16_RTOL, "mask logits (bf16 out)")
    prob = dfine_b200.mask_logits(coef, proto, apply_sigmoid=True, out_dtype=torch.float32)
    assert_close(prob.cpu().numpy(), g["probs"], 2e-5, "mask probs")
    torch.manual_seed(1)
    for (B, M, K, N) in [(2, 500, 256, 160 * 160), (1, 300, 256, 160 * 160), (3, 200, 128, 1000), (1, 77, 64, 264)]:
        a = torch.randn(B, M, K, device=dev).to(torch.bfloat16)
        b = torch.randn(B, K, N, device=dev).to(torch.bfloat16)
        got = ops.mask_gemm_raw(a, b, torch.float32, False)
        want = torch.bmm(a.float(), b.float())
        assert rel_err(got.cpu().numpy(), want.cpu().numpy()) <= 2e-5, (B, M, K, N)
        gotb = ops.mask_gemm_raw(a, b, torch.bfloat16, False)
        assert rel_err(gotb.float().cpu().numpy(), want.cpu().numpy()) <= BF16_RTOL, (B, M, K, N)


def test_mask_gemm_backward(dev):
    """dfine_mask_gemm_bwd (both contractions on tcgen05) against the reference's gradients (golden,
    autograd of the einsum at dfine_decoder.py:940), against the oracle and against fp32 bmm at the
    config-4 shapes (Q = 300 regular / 200 denoising / 500 bench queries; ragged M and N edges)."""
    import dfine_b200
    import dfine_b200.ops as ops
    from oracle import cpu_oracle as O
    g = golden("mask_bwd")
    coef = _t(g["coef"], dev).requires_grad_(True)
    proto = _t(g["proto"], dev).requires_grad_(True)
    out = dfine_b200.mask_logits(coef, proto, out_dtype=torch.float32)
    assert_close(out.detach().cpu().numpy(), g["logits"], 2e-5, "mask logits (K = 128)")
    out.backward(_t(g["grad_out"], dev))
    # through autograd the gradients pass through bf16 (mask_logits casts its operands to bf16, as autocast
    # does for the reference's einsum): bf16 tolerance
    assert coef.grad.dtype == torch.float32 and proto.grad.dtype == torch.float32
    assert_close(coef.grad.cpu().numpy(), g["grad_coef"], BF16_RTOL, "grad_coef (golden, autograd)")
    assert_close(proto.grad.cpu().numpy(), g["grad_proto"], BF16_RTOL, "grad_proto (golden, autograd)")
    # the kernels themselves: inputs and grad_out are bf16-representable, so products are exact and the
    # fp32 results differ from the reference's only by the accumulation order
    cb, pb, gb = (_t(g[k], dev).to(torch.bfloat16) for k in ("coef", "proto", "grad_out"))
    gc, gp = ops.mask_gemm_bwd_raw(cb, pb.flatten(2), gb.flatten(2), proto_dtype=torch.float32)
    assert_close(gc.cpu().numpy(), g["grad_coef"], 2e-5, "grad_coef (golden)")
    assert_close(gp.cpu().numpy().reshape(g["grad_proto"].shape), g["grad_proto"], 2e-5, "grad_proto (golden)")
    torch.manual_seed(2)
    for (B, M, K, N) in [(2, 500, 256, 160 * 160), (2, 300, 256, 160 * 160), (3, 200, 128, 1000), (1, 77, 128, 264),
                         (16, 300, 256, 25600)]:
        a = torch.randn(B, M, K, device=dev).to(torch.bfloat16)
        b = torch.randn(B, K, N, device=dev).to(torch.bfloat16)
        go = torch.randn(B, M, N, device=dev).to(torch.bfloat16)
        # matched-rows-only upstream gradient (what the criterion sends: dfine_criterion.py:336)
        if M == 300:
            keep = torch.zeros(B, M, 1, device=dev, dtype=torch.bfloat16)
            keep[:, ::29] = 1
            go = go * keep
        gc, gp = ops.mask_gemm_bwd_raw(a, b, go, proto_dtype=torch.float32)
        # float64 reference: the kernels' only error is the fp32 accumulation order of a 25600- (grad_coef) or
        # M-term (grad_proto) sum; elementwise bound |a-b| <= 2e-5 (|b| + RMS)
        want_c = torch.bmm(go.double(), b.double().transpose(1, 2))
        want_p = torch.bmm(a.double().transpose(1, 2), go.double())
        if M == 300:   # rows without an upstream gradient are exactly zero; the others are compared on their own scale
            assert not gc[:, 1::29].any() and not gc[:, 2::29].any()
            gc_c, want_cc = gc[:, ::29], want_c[:, ::29]
        else:
            gc_c, want_cc = gc, want_c
        assert elementwise_err(gc_c.cpu().numpy(), want_cc.cpu().numpy(), 2e-5) <= 1.0, ("grad_coef", B, M, K, N)
        assert elementwise_err(gp.cpu().numpy(), want_p.cpu().numpy(), 2e-5) <= 1.0, ("grad_proto", B, M, K, N)
        assert rel_err(gc.cpu().numpy(), want_c.cpu().numpy()) <= 2e-5, ("grad_coef", B, M, K, N)
        assert rel_err(gp.cpu().numpy(), want_p.cpu().numpy()) <= 2e-5, ("grad_proto", B, M, K, N)
        if B * M * N <= 4_000_000:
            oc, op = O.mask_gemm_bwd(a.float().cpu().numpy(), b.float().cpu().numpy(), go.float().cpu().numpy())
            assert rel_err(gc.cpu().numpy(), oc) <= 2e-5 and rel_err(gp.cpu().numpy(), op) <= 2e-5
        _, gpb = ops.mask_gemm_bwd_raw(a, b, go, want_coef=False)
        assert gpb.dtype == torch.bfloat16
        assert rel_err(gpb.float().cpu().numpy(), want_p.cpu().numpy()) <= BF16_RTOL
        gc2, none = ops.mask_gemm_bwd_raw(a, b, go, want_proto=False)
        assert none is None and torch.equal(gc2.isfinite(), gc.isfinite())
        assert rel_err(gc2.cpu().numpy(), want_c.cpu().numpy()) <= 2e-5
    # shapes the backward kernels do not take fail loudly (no silent library fallback)
    a = torch.randn(1, 8, 64, device=dev).to(torch.bfloat16)
    with pytest.raises(dfine_b200.DfineB200Error):
        ops.mask_gemm_bwd_raw(a, torch.randn(1, 64, 16, device=dev).to(torch.bfloat16),
                              torch.randn(1, 8, 16, device=dev).to(torch.bfloat16))


def test_packed_linear_matches_two_linears(dev):
    """dfine_pack_linear + one GEMM against the reference's two nn.Linear calls (fp32 and
    bf16 autocast): outputs and all five gradients."""
    import dfine_b200.ops as ops
    torch.manual_seed(9)
    torch.backends.cuda.matmul.allow_tf32 = False
    B, Lq, C, n0, n1 = 3, 70, 256, 192, 96
    x0 = torch.randn(B, Lq, C, device=dev)
    ws = [torch.randn(n0, C, device=dev) * 0.05, torch.randn(n0, device=dev) * 0.1,
          torch.randn(n1, C, device=dev) * 0.05, torch.randn(n1, device=dev) * 0.1]
    g = torch.randn(B, Lq, n0 + n1, device=dev)
    for amp in (False, True):
        def run(ours):
            x = x0.clone().requires_grad_(True)
            p = [w.clone().requires_grad_(True) for w in ws]
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                if ours:
                    y = ops.packed_linear(x, *p)
                else:
                    y = torch.cat([torch.nn.functional.linear(x, p[0], p[1]),
                                   torch.nn.functional.linear(x, p[2], p[3])], -1)
            y.float().backward(g)
            return [y.detach().float(), x.grad] + [t.grad for t in p]
        got, want = run(True), run(False)
        tol = 2e-2 if amp else 2e-5
        for a, b, n in zip(got, want, ["y", "gx", "gw0", "gb0", "gw1", "gb1"]):
            assert a.shape == b.shape
            assert rel_err(a.cpu().numpy(), b.float().cpu().numpy()) <= tol, (n, amp)


def test_colsum(dev):
    """dfine_colsum (bias gradient of the concatenated Linear) against a float64 sum."""
    import dfine_b200.ops as ops
    torch.manual_seed(2)
    for (M, N, dt) in [(16000, 288, torch.bfloat16), (16000, 288, torch.float32), (77, 96, torch.bfloat16),
                       (1, 2, torch.float32), (4099, 1024, torch.bfloat16)]:
        x = torch.randn(M, N, device=dev).to(dt)
        want = x.double().sum(0)
        got = ops.colsum(x)
        assert got.dtype == torch.float32 and got.shape == (N,)
        scale = x.double().abs().sum(0).max().item()
        assert (got.double() - want).abs().max().item() <= 1e-6 * scale, (M, N, dt)
    # a strided view (row stride > width)
    big = torch.randn(500, 300, device=dev)
    assert torch.allclose(ops.colsum(big[:, :288]), big[:, :288].sum(0), rtol=1e-5, atol=1e-4)
    with pytest.raises(ValueError):
        ops.colsum(torch.randn(8, 7, device=dev))


def test_linear_wgrad(dev):
    """dfine_linear_wgrad (tcgen05: dW = gy^T x and db = column sums of gy in one launch) against a
    float64 product of the same bf16 inputs; shapes with a ragged row tile / row block / column box,
    strided operands, and the packed-Linear backward that uses it against the cuBLAS route."""
    import dfine_b200.ops as ops
    torch.manual_seed(3)
    for (M, N, K) in [(16000, 288, 256), (4800, 288, 128), (77, 96, 64), (1, 8, 8), (8009, 136, 200), (64, 128, 256)]:
        gy = torch.randn(M, N, device=dev).to(torch.bfloat16)
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        gw, gb = ops.linear_wgrad(gy, x)
        assert gw.shape == (N, K) and gb.shape == (N,) and gw.dtype == gb.dtype == torch.float32
        want_w = gy.double().t() @ x.double()
        want_b = gy.double().sum(0)
        # fp32 accumulation of exact bf16 products: error ~ sqrt(M) * 2^-24 of the magnitude
        assert rel_err(gw.double().cpu().numpy(), want_w.cpu().numpy()) <= 1e-5, (M, N, K)
        assert rel_err(gb.double().cpu().numpy(), want_b.cpu().numpy()) <= 1e-5, (M, N, K)
    # row-strided views of wider buffers
    big_g = torch.randn(500, 320, device=dev).to(torch.bfloat16)
    big_x = torch.randn(500, 264, device=dev).to(torch.bfloat16)
    gw, gb = ops.linear_wgrad(big_g[:, :288], big_x[:, :256])
    assert rel_err(gw.double().cpu().numpy(), (big_g[:, :288].double().t() @ big_x[:, :256].double()).cpu().numpy()) <= 1e-5
    assert rel_err(gb.double().cpu().numpy(), big_g[:, :288].double().sum(0).cpu().numpy()) <= 1e-5
    with pytest.raises(ValueError):
        ops.linear_wgrad(torch.zeros(8, 7, device=dev, dtype=torch.bfloat16), torch.zeros(8, 8, device=dev, dtype=torch.bfloat16))
    with pytest.raises(ValueError):   # K > 256 is left to the library GEMM
        ops.linear_wgrad(torch.zeros(8, 8, device=dev, dtype=torch.bfloat16), torch.zeros(8, 512, device=dev, dtype=torch.bfloat16))
    # the packed Linear's backward through it == through cuBLAS + dfine_colsum
    q = torch.randn(4, 50, 256, device=dev)
    lin = [torch.nn.Linear(256, 192).to(dev), torch.nn.Linear(256, 96).to(dev)]
    go = torch.randn(4, 50, 288, device=dev)
    grads = {}
    for mode in ("1", "0"):
        os.environ["DFINE_LINEAR_WGRAD"] = mode
        try:
            for l in lin:
                l.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = ops.packed_linear(q, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias)
            y.backward(go.to(y.dtype))
            grads[mode] = [p.grad.clone() for l in lin for p in (l.weight, l.bias)]
        finally:
            os.environ.pop("DFINE_LINEAR_WGRAD", None)
    for a, b in zip(grads["1"], grads["0"]):
        assert a.shape == b.shape and rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-5


def test_error_behaviour(dev):
    import dfine_b200
    import dfine_b200.ops as ops
    spec = ops.level_spec([[4, 4]], [2])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dfine_b200.msda_core(torch.zeros(1, 16, 2, 16), [[4, 4]], torch.zeros(1, 1, 2, 2, 2),
                             torch.zeros(1, 1, 2, 2), [2])
    with pytest.raises(dfine_b200.DfineB200Error, match="head_dim"):
        ops.msda_forward_raw(torch.zeros(1, 16, 2 * 24, device=dev), spec, 2,
                             torch.zeros(1, 1, 2, 2, 2, device=dev), torch.zeros(1, 1, 2, 2, device=dev),
                             None, None, 0.5, False, torch.float32)
    with pytest.raises(NotImplementedError):
        dfine_b200.msda_core(torch.zeros(1, 16, 2, 16, device=dev), [[4, 4]],
                             torch.zeros(1, 1, 2, 2, 2, device=dev), torch.zeros(1, 1, 2, 2, device=dev),
                             [2], method="discrete")


def test_packed_raw_and_fused_linear(dev):
    """One concatenated Linear output [B, Lq, 3HP] through row strides, bf16 gradients written
    by the kernel, cuBLAS-only Linear backward: against the separate-tensor path."""
    import dfine_b200.ops as ops
    g = golden("module_m_small")
    H = int(g["H"])
    shapes, npts = g["shapes"].tolist(), g["npts"].tolist()
    spec = ops.level_spec(shapes, npts)
    mem32 = bf16_bits_to_f32(g["memory_bf16"])
    B, L, C = mem32.shape
    Lq = g["raw_off"].shape[1]
    go = _t(bf16_bits_to_f32(g["grad_out_bf16"]), dev)
    ref = _t(g["ref_points"].reshape(B, -1, 4), dev)
    nps = _t(g["num_points_scale"], dev)
    for dt, tol in ((torch.float32, FP32_RTOL), (torch.bfloat16, BF16_RTOL)):
        raw = torch.cat([_t(g["raw_off"], dev).reshape(B, Lq, -1), _t(g["raw_logit"], dev).reshape(B, Lq, -1)], -1)
        raw = raw.to(dt).requires_grad_(True)
        mem = _t(mem32, dev, dt).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt == torch.bfloat16):
            out = ops.msda_fused_packed(mem.reshape(B, L, H, C // H), shapes, raw, ref, nps, npts)
        out.backward(go)
        # reference: the two-tensor path on the same (rounded) inputs
        off = raw.detach()[..., :2 * H * spec.P].reshape(B, Lq, H, spec.P, 2).contiguous()
        logit = raw.detach()[..., 2 * H * spec.P:].reshape(B, Lq, H, spec.P).contiguous()
        want = ops.msda_forward_raw(mem.detach(), spec, H, off, logit, ref, nps, 0.5, True, torch.float32)
        gm, gs, ga = ops.msda_backward_raw(mem.detach(), spec, H, off, logit, ref, nps, 0.5, True, go)
        assert torch.equal(out.detach(), want)
        assert_close(mem.grad.float().cpu().numpy(), gm.cpu().numpy(), tol, "packed grad_memory")
        assert_close(raw.grad[..., :2 * H * spec.P].float().cpu().numpy(),
                     gs.reshape(B, Lq, -1).cpu().numpy(), tol, "packed grad offsets")
        assert_close(raw.grad[..., 2 * H * spec.P:].float().cpu().numpy(),
                     ga.reshape(B, Lq, -1).cpu().numpy(), tol, "packed grad logits")
    # fused Linear (cuBLAS) against F.linear autograd
    torch.manual_seed(0)
    x = torch.randn(3, 50, 64, device=dev, requires_grad=True)
    w = torch.randn(36, 64, device=dev, requires_grad=True)
    b = torch.randn(36, device=dev, requires_grad=True)
    gy = torch.randn(3, 50, 36, device=dev)
    ops.fused_linear(x, w, b).backward(gy)
    got = [t.grad.clone() for t in (x, w, b)]
    for t in (x, w, b):
        t.grad = None
    torch.nn.functional.linear(x, w, b).backward(gy)
    for a_, b_ in zip(got, (x.grad, w.grad, b.grad)):
        assert rel_err(a_.cpu().numpy(), b_.cpu().numpy()) <= 5e-5


ZOO = [
    # name, B, Lq, H, c, shapes, npts, value dtype
    ("n_model_c16_two_levels", 2, 300, 8, 16, [[40, 40], [20, 20]], [6, 6], torch.bfloat16),
    ("n_model_c16_fp32", 1, 77, 8, 16, [[40, 40], [20, 20]], [6, 6], torch.float32),
    ("rectangular_multiscale_aug", 2, 123, 8, 32, [[72, 88], [36, 44], [18, 22]], [3, 6, 3], torch.bfloat16),
    ("four_levels_16_points", 1, 50, 8, 32, [[32, 32], [16, 16], [8, 8], [4, 4]], [4, 4, 4, 4], torch.float32),
    ("many_points_single_item_per_warp", 1, 33, 4, 32, [[16, 16], [8, 8]], [10, 10], torch.float32),
    ("odd_heads_runtime_division", 1, 41, 6, 32, [[12, 12], [6, 6]], [4, 4], torch.bfloat16),
    ("wide_heads_c64", 1, 29, 4, 64, [[20, 20], [10, 10], [5, 5]], [3, 6, 3], torch.bfloat16),
    ("wide_heads_c64_fp32", 1, 29, 4, 64, [[20, 20], [10, 10]], [4, 4], torch.float32),
    ("one_level_one_point", 3, 7, 2, 16, [[9, 5]], [1], torch.float32),
    ("x_model_1024", 1, 500, 8, 32, [[128, 128], [64, 64], [32, 32]], [4, 4, 4], torch.bfloat16),
]


@pytest.mark.parametrize("case", ZOO, ids=[z[0] for z in ZOO])
def test_shape_zoo_against_oracle(case, dev):
    """Every template instantiation / runtime path (head widths 16/32/64, fp32 and bf16 value,
    compile-time and runtime P, one or two items per warp, non-power-of-two head counts,
    rectangular and 4-level pyramids, chunked level 0) against the CPU oracle: forward,
    corner indices, and all three gradients in plain and fused input mode."""
    import dfine_b200.ops as ops
    from oracle import cpu_oracle as O
    name, B, Lq, H, c, shapes, npts, vdt = case
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    spec = ops.level_spec(shapes, npts)
    P = spec.P
    mem = torch.from_numpy(rng.standard_normal((B, spec.L, H * c), dtype=np.float32)).to(torch.bfloat16)
    raw_off = torch.from_numpy(rng.standard_normal((B, Lq, H, P, 2), dtype=np.float32) * 1.5).to(torch.bfloat16).float()
    raw_log = torch.from_numpy(rng.standard_normal((B, Lq, H, P), dtype=np.float32)).to(torch.bfloat16).float()
    ref = np.concatenate([rng.uniform(-0.1, 1.1, (B, Lq, 2)), rng.uniform(0.02, 1.0, (B, Lq, 2))], -1).astype(np.float32)
    nps = np.asarray([1.0 / n for n in npts for _ in range(n)], np.float32)
    go = rng.standard_normal((B, Lq, H * c), dtype=np.float32)
    m32 = mem.float().numpy().reshape(B, spec.L, H, c)
    memd = mem.to(dev).to(vdt)
    # fused mode
    args = (spec, H, raw_off.to(dev), raw_log.to(dev), _t(ref, dev), _t(nps, dev), 0.5, True)
    out, idx = ops.msda_forward_raw(memd, *args, torch.float32, want_idx=True)
    loc = O.msda_locations(raw_off.numpy(), ref, nps, 0.5)
    attn = O.softmax(raw_log.numpy())
    o_out, o_idx, _ = O.msda_fwd(m32, shapes, npts, loc, attn, want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), o_idx), "corner indices"
    assert_close(out.cpu().numpy(), o_out, FP32_RTOL, "fused out")
    o_gv, o_gs, o_ga = O.msda_fused_bwd(m32, shapes, npts, raw_off.numpy(), raw_log.numpy(), ref, nps, go, 0.5)
    if P <= 16:
        for atomic in (False, True):
            gm, gs, ga = ops.msda_backward_raw(memd, *args, _t(go, dev), force_atomic=atomic)
            assert_close(gm.cpu().numpy().reshape(o_gv.shape), o_gv, FP32_RTOL, f"grad_value atomic={atomic}")
            assert_close(gs.cpu().numpy(), o_gs, FP32_RTOL, "grad_offsets")
            assert_close(ga.cpu().numpy(), o_ga, FP32_RTOL, "grad_logits")
        # records written by the forward, consumed by the backward
        rec = ops.new_records(memd, spec, H, Lq)
        out2 = ops.msda_forward_raw(memd, *args, torch.float32, records=rec)
        assert torch.equal(out2, out)
        gm, gs, ga = ops.msda_backward_raw(memd, *args, _t(go, dev), records=rec)
        assert_close(gm.cpu().numpy().reshape(o_gv.shape), o_gv, FP32_RTOL, "grad_value (records)")
        assert_close(gs.cpu().numpy(), o_gs, FP32_RTOL, "grad_offsets (records)")
        assert_close(ga.cpu().numpy(), o_ga, FP32_RTOL, "grad_logits (records)")
    # plain mode
    locd, attnd = _t(loc, dev), _t(attn, dev)
    outp = ops.msda_forward_raw(memd, spec, H, locd, attnd, None, None, 0.5, False, torch.float32)
    assert_close(outp.cpu().numpy(), o_out, FP32_RTOL, "plain out")
    if P <= 16:
        p_gv, p_gl, p_ga = O.msda_bwd(m32, shapes, npts, loc, attn, go)
        gm, gl, ga = ops.msda_backward_raw(memd, spec, H, locd, attnd, None, None, 0.5, False, _t(go, dev))
        assert_close(gm.cpu().numpy().reshape(p_gv.shape), p_gv, FP32_RTOL, "plain grad_value")
        assert_close(gl.cpu().numpy(), p_gl, FP32_RTOL, "plain grad_loc")
        assert_close(ga.cpu().numpy(), p_ga, FP32_RTOL, "plain grad_attn")


def test_decoder_loop_integration(dev):
    """A decoder-shaped graph on the GPU: `memory` -> value_op views (once) -> 3 cross-attention
    layers (patched mirror modules, fused path) chained through their queries, FDR decode per
    layer, one backward.  Compared with the same graph built from the reference call
    sequence (oracle/torch_port.py, eager PyTorch fp32) with identical parameters: outputs,
    boxes and the gradients w.r.t. memory (accumulated over the layers through the zero-copy
    route), the first query and every Linear parameter."""
    import dfine_b200
    from oracle import torch_port as TP
    torch.manual_seed(5)
    torch.backends.cuda.matmul.allow_tf32 = False
    B, Lq, C, H = 2, 60, 256, 8
    shapes, npts, n_layers = [[20, 20], [10, 10], [5, 5]], [3, 6, 3], 3
    L = sum(h * w for h, w in shapes)
    mods = []
    for _ in range(n_layers):
        m = dfine_b200.MSDeformableAttention(C, H, len(shapes), npts).to(dev)
        with torch.no_grad():
            m.sampling_offsets.weight.normal_(0, 0.02)
            m.attention_weights.weight.normal_(0, 0.05)
            m.attention_weights.bias.normal_(0, 0.1)
        mods.append(m)
    memory0 = torch.randn(B, L, C, device=dev)
    query0 = torch.randn(B, Lq, C, device=dev)
    ref = torch.cat([torch.rand(B, Lq, 2, device=dev), torch.rand(B, Lq, 2, device=dev) * 0.5 + 0.05], -1)
    corners = torch.randn(n_layers, B, Lq, 132, device=dev)
    up, rs = torch.tensor([0.5], device=dev), torch.tensor([4.0], device=dev)
    g_out = torch.randn(B, Lq, C, device=dev)
    g_box = torch.randn(n_layers, B, Lq, 4, device=dev)

    def run(ours: bool):
        mem = memory0.clone().requires_grad_(True)
        q = query0.clone().requires_grad_(True)
        pc = corners.clone().requires_grad_(True)
        for m in mods:
            m.zero_grad(set_to_none=True)
        value = TP.value_views(mem, H, shapes)   # the reference's value_op layout, built once
        project = dfine_b200.fdr_project(up, rs) if ours else TP.weighting_function(32, up, rs)
        x, boxes = q, []
        for i, m in enumerate(mods):
            if ours:
                y = m(x, ref.unsqueeze(2), value, shapes)
                boxes.append(dfine_b200.fdr_decode(pc[i], ref, project, rs))
            else:
                y = TP.msda_module(x, ref.unsqueeze(2), value, shapes, m.sampling_offsets.weight,
                                   m.sampling_offsets.bias, m.attention_weights.weight,
                                   m.attention_weights.bias, m.num_points_scale, npts, H)
                boxes.append(TP.distance2bbox(ref, TP.integral(pc[i], project), rs))
            x = x + torch.tanh(y)                # the next layer's query depends on this layer
        boxes = torch.stack(boxes)
        torch.autograd.backward([x, boxes], [g_out, g_box])
        grads = [mem.grad, q.grad, pc.grad] + [p.grad for m in mods for p in m.parameters()]
        return [x.detach(), boxes.detach()] + [g.clone() for g in grads]

    got, want = run(True), run(False)
    names = ["out", "boxes", "grad_memory", "grad_query", "grad_corners"] + \
            [f"layer{i}.{n}" for i in range(n_layers) for n, _ in mods[0].named_parameters()]
    for n, a, b in zip(names, got, want):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4, n
    assert rel_err(got[0].cpu().numpy(), want[0].cpu().numpy()) <= 2e-5
    assert rel_err(got[2].cpu().numpy(), want[2].cpu().numpy()) <= 2e-5


@pytest.mark.parametrize("overlap", [False, True], ids=["one_stream", "two_streams"])
@pytest.mark.parametrize("amp", [False, True], ids=["fp32", "amp_bf16"])
def test_shared_memory_gradient_hub(amp, overlap, dev):
    """All layers sample one `memory`; their grad_value kernels accumulate into ONE buffer that
    a hub node hands to autograd once (ops.share_memory_grad).  Must equal autograd's own
    per-layer accumulation -- also when a layer is left out of the loss and for a second
    backward over a retained graph."""
    import dfine_b200
    from dfine_b200 import ops
    from oracle import torch_port as TP
    torch.manual_seed(11)
    B, Lq, C, H = 2, 50, 256, 8
    shapes, npts, n_layers = [[16, 12], [8, 6], [4, 3]], [3, 6, 3], 4
    L = sum(h * w for h, w in shapes)
    mods = []
    for _ in range(n_layers):
        m = dfine_b200.MSDeformableAttention(C, H, len(shapes), npts).to(dev)
        with torch.no_grad():
            m.sampling_offsets.weight.normal_(0, 0.02)
            m.attention_weights.weight.normal_(0, 0.05)
        mods.append(m)
    enc = torch.randn(B, L, C, device=dev)
    qs = [torch.randn(B, Lq, C, device=dev) for _ in range(n_layers)]
    ref = torch.cat([torch.rand(B, Lq, 2, device=dev), torch.rand(B, Lq, 2, device=dev) * 0.5 + 0.05], -1)
    gos = [torch.randn(B, Lq, C, device=dev) for _ in range(n_layers)]

    def run(share, used, twice=False):
        old = ops.share_memory_grad(share)
        old_ov = ops.overlap_grad_value(overlap and share)
        try:
            leaf = enc.clone().requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                mem = (leaf * 1.5).to(torch.bfloat16) if amp else leaf * 1.5  # non-leaf, like the encoder output
                value = TP.value_views(mem, H, shapes)
                outs = [m(q, ref.unsqueeze(2), value, shapes) for m, q in zip(mods, qs)]
            sel = [o for o, u in zip(outs, used) if u]
            gsel = [g for g, u in zip(gos, used) if u]
            torch.autograd.backward(sel, gsel, retain_graph=twice)
            if twice:
                torch.autograd.backward(sel, gsel)
            return leaf.grad.clone()
        finally:
            ops.share_memory_grad(old)
            ops.overlap_grad_value(old_ov)

    tol = BF16_RTOL if amp else FP32_RTOL
    for used in ([True] * 4, [True, False, True, True], [False, False, True, False]):
        want = run(False, used)
        got = run(True, used)
        assert_close(got.cpu().numpy(), want.cpu().numpy(), tol, f"grad through hub, layers {used}")
    want2 = run(False, [True] * 4, twice=True)
    got2 = run(True, [True] * 4, twice=True)
    assert_close(got2.cpu().numpy(), want2.cpu().numpy(), tol, "second backward over a retained graph")
    assert len(ops._HUBS) <= ops._MAX_HUBS


def test_hub_ignores_a_pass_that_never_reached_it(dev):
    """torch.autograd.grad(..., inputs=[query], retain_graph=True) runs the layers' backward kernels but not the
    hub node: the sum it leaves in the shared buffer must not leak into the next backward() over the retained
    graph (the first layer of a new pass -- a new autograd graph task -- starts a fresh buffer)."""
    import dfine_b200
    from dfine_b200 import ops
    from oracle import torch_port as TP
    torch.manual_seed(12)
    B, Lq, C, H = 2, 40, 256, 8
    shapes, npts = [[16, 12], [8, 6], [4, 3]], [3, 6, 3]
    L = sum(h * w for h, w in shapes)
    mods = []
    for _ in range(3):
        m = dfine_b200.MSDeformableAttention(C, H, len(shapes), npts).to(dev)
        with torch.no_grad():
            m.sampling_offsets.weight.normal_(0, 0.02)
            m.attention_weights.weight.normal_(0, 0.05)
        mods.append(m)
    enc = torch.randn(B, L, C, device=dev)
    qs = [torch.randn(B, Lq, C, device=dev, requires_grad=True) for _ in mods]
    ref = torch.cat([torch.rand(B, Lq, 2, device=dev), torch.rand(B, Lq, 2, device=dev) * 0.5 + 0.05], -1)
    gos = [torch.randn(B, Lq, C, device=dev) for _ in mods]

    def run(share, probe_first):
        old = ops.share_memory_grad(share)
        try:
            leaf = enc.clone().requires_grad_(True)
            value = TP.value_views(leaf * 1.5, H, shapes)
            outs = [m(q, ref.unsqueeze(2), value, shapes) for m, q in zip(mods, qs)]
            if probe_first:     # layer backwards run, the hub does not
                torch.autograd.grad(outs, qs, gos, retain_graph=True)
                torch.autograd.grad(outs[:2], qs[:2], gos[:2], retain_graph=True)
            torch.autograd.backward(outs, gos)
            return leaf.grad.clone()
        finally:
            ops.share_memory_grad(old)

    want = run(False, False)
    assert_close(run(True, False).cpu().numpy(), want.cpu().numpy(), FP32_RTOL, "hub, plain backward")
    assert_close(run(True, True).cpu().numpy(), want.cpu().numpy(), FP32_RTOL, "hub after autograd.grad(inputs=[query])")


def test_many_queries_fall_back_to_reductions(dev):
    """More sampling points per (image, head, level) than the grad_value kernel can list in
    shared memory (node ids are 16-bit): the library switches to fp32 vector reductions, the
    shared bf16 gradient buffer of the hub still accumulates (one add), results unchanged."""
    import dfine_b200
    from dfine_b200 import ops
    from oracle import torch_port as TP
    torch.manual_seed(4)
    B, Lq, H, c = 1, 2800, 8, 32
    shapes, npts = [[12, 10], [6, 5]], [6, 6]          # 4 * 6 * 2800 = 67200 > 65535 nodes
    spec = ops.level_spec(shapes, npts)
    P = spec.P
    ref = torch.cat([torch.rand(B, Lq, 2, device=dev), torch.rand(B, Lq, 2, device=dev) * 0.5 + 0.05], -1)
    nps = torch.tensor([1.0 / n for n in npts for _ in range(n)], device=dev)
    raws = [torch.randn(B, Lq, 3 * H * P, device=dev).to(torch.bfloat16) for _ in range(2)]
    gos = [torch.randn(B, Lq, H * c, device=dev) for _ in range(2)]
    mem0 = torch.randn(B, spec.L, H * c, device=dev).to(torch.bfloat16)

    def run(share):
        old = ops.share_memory_grad(share)
        try:
            mem = mem0.clone().requires_grad_(True)
            value = TP.value_views(mem, H, shapes)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                outs = [dfine_b200.ops.msda_fused_packed(value, shapes, r, ref.unsqueeze(2), nps, npts) for r in raws]
            torch.autograd.backward(outs, gos)
            return outs[0].detach(), mem.grad.float()
        finally:
            ops.share_memory_grad(old)

    (o1, g1), (o0, g0) = run(True), run(False)
    assert torch.equal(o1, o0)
    assert_close(g1.cpu().numpy(), g0.cpu().numpy(), BF16_RTOL, "grad_memory through the fallback")
    m32 = mem0.float().requires_grad_(True)
    want = sum((TP.msda_from_raw(r.float()[..., :2 * H * P].reshape(B, Lq, H, P, 2),
                                 r.float()[..., 2 * H * P:].reshape(B, Lq, H, P), ref.unsqueeze(2),
                                 TP.value_views(m32, H, shapes), shapes, nps, npts) * g).sum()
               for r, g in zip(raws, gos))
    want.backward()
    assert_close(g1.cpu().numpy(), m32.grad.cpu().numpy(), BF16_RTOL, "grad_memory vs eager reference")


def test_detection_f1_and_mask_iou_unchanged(dev):
    """North-star criterion: "detection F1 / mask IoU must be unchanged on a fixed synthetic set".
    A small detector is assembled from the hot path -- `memory` -> value_op views -> 3 chained
    cross-attention layers -> class head, FDR box decode, mask-prototype contraction -- and run
    on a fixed synthetic set (8 images x 10 ground-truth rectangles, planted so that the
    detector finds most of them) once through the reference call sequence (oracle/torch_port.py,
    eager PyTorch) and once through the CUDA path, both under bf16 autocast.  The reference's
    metric (restated in oracle/det_metrics.py: top-300 postprocess, greedy IoU matching, F1,
    mask IoU) must come out identical, and the pre-threshold top-300 lists must agree."""
    import dfine_b200
    from oracle import det_metrics as DM
    from oracle import torch_port as TP
    torch.manual_seed(123)
    torch.backends.cuda.matmul.allow_tf32 = False
    B, Q, C, H, NC, n_gt = 8, 300, 256, 8, 80, 10
    shapes, npts, n_layers = [[40, 40], [20, 20], [10, 10]], [3, 6, 3], 3
    Hm = Wm = 80
    L = sum(h * w for h, w in shapes)
    # ground truth + planted queries: query i < n_gt looks at object i
    gt_xy = torch.rand(B, n_gt, 2, device=dev) * 0.6 + 0.2
    gt_wh = torch.rand(B, n_gt, 2, device=dev) * 0.2 + 0.08
    gt_boxes = torch.cat([gt_xy, gt_wh], -1)
    gt_labels = torch.randint(0, NC, (B, n_gt), device=dev)
    ref = torch.cat([torch.rand(B, Q, 2, device=dev), torch.rand(B, Q, 2, device=dev) * 0.3 + 0.05], -1)
    ref[:, :n_gt] = gt_boxes * (1 + 0.03 * torch.randn(B, n_gt, 4, device=dev))
    # class head = planted logits + a small data-dependent term: objects 0-7 are found (+3), objects
    # 8-9 are missed (-1 -> false negatives), queries 10-14 fire on random boxes (+1 -> false
    # positives); margins of >= 4 sigma of the data term keep every decision off the threshold
    planted = torch.full((B, Q, NC), -4.0, device=dev)
    strength = torch.tensor([3.0] * 8 + [-1.0] * 2, device=dev).view(1, n_gt, 1).expand(B, n_gt, 1)
    planted[:, :n_gt].scatter_(2, gt_labels.unsqueeze(-1), strength.contiguous())
    planted[:, n_gt:n_gt + 5, 7] = 1.0
    memory = torch.randn(B, L, C, device=dev)
    query = torch.randn(B, Q, C, device=dev)
    mask_feat = torch.randn(B, C, Hm, Wm, device=dev)
    mods = []
    for _ in range(n_layers):
        m = dfine_b200.MSDeformableAttention(C, H, len(shapes), npts).to(dev)
        with torch.no_grad():
            m.sampling_offsets.weight.normal_(0, 0.02)
            m.attention_weights.weight.normal_(0, 0.05)
        mods.append(m)
    w_cls = torch.randn(NC, C, device=dev) * 0.01
    w_cor = torch.randn(132, C, device=dev) * 0.05
    w_msk = torch.randn(C, C, device=dev) * 0.06
    up, rs = torch.tensor([0.5], device=dev), torch.tensor([4.0], device=dev)

    def detector(ours: bool):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            mem = memory.to(torch.bfloat16)
            value = TP.value_views(mem, H, shapes)
            project = dfine_b200.fdr_project(up, rs) if ours else TP.weighting_function(32, up, rs)
            x = query
            for m in mods:
                if ours:
                    y = m(x, ref.unsqueeze(2), value, shapes)
                else:
                    y = TP.msda_module(x, ref.unsqueeze(2), value, shapes, m.sampling_offsets.weight,
                                       m.sampling_offsets.bias, m.attention_weights.weight,
                                       m.attention_weights.bias, m.num_points_scale, npts, H)
                x = x + torch.tanh(y.float())
            logits = planted + torch.nn.functional.linear(x, w_cls).float()
            corners = torch.nn.functional.linear(x, w_cor)
            coef = torch.nn.functional.linear(x, w_msk)
            if ours:
                boxes = dfine_b200.fdr_decode(corners, ref, project, rs)
                masks = dfine_b200.mask_logits(coef, mask_feat.to(torch.bfloat16), apply_sigmoid=True)
            else:
                boxes = TP.distance2bbox(ref, TP.integral(corners, project), rs)
                masks = torch.sigmoid(TP.mask_logits(coef, mask_feat.to(torch.bfloat16)))
        return logits.float(), boxes.float(), masks.float()

    gts = [{"boxes": DM.box_cxcywh_to_xyxy(gt_boxes[b]).cpu(), "labels": gt_labels[b].cpu()} for b in range(B)]
    ys, xs = torch.meshgrid(torch.arange(Hm, device=dev), torch.arange(Wm, device=dev), indexing="ij")

    def evaluate(out):
        logits, boxes, masks = out
        preds = DM.postprocess(logits, boxes, conf_thresh=0.5, num_top_queries=300)
        tps, fps, fns, ious, matches = DM.f1_counts(preds, gts, 0.5)
        miou = []
        for b, pairs in enumerate(matches):
            for p_, g_ in pairs:
                x1, y1, x2, y2 = (DM.box_cxcywh_to_xyxy(gt_boxes[b, g_]) * Wm).tolist()
                gt_mask = ((xs >= x1) & (xs < x2) & (ys >= y1) & (ys < y2)).cpu().numpy()
                pm = (masks[b, int(preds[b]["queries"][p_])] >= 0.5).cpu().numpy()
                miou.append(DM.mask_iou(pm, gt_mask))
        return preds, (tps, fps, fns), DM.f1_score(tps, fps, fns), ious, miou

    p_ref, cnt_ref, f1_ref, iou_ref, miou_ref = evaluate(detector(False))
    p_our, cnt_our, f1_our, iou_our, miou_our = evaluate(detector(True))
    assert cnt_ref[0] >= 0.5 * B * n_gt, f"the synthetic set should be mostly detected, got {cnt_ref}"
    assert cnt_ref[1] > 0 and cnt_ref[2] > 0, "the set should also contain false positives / negatives"
    assert cnt_our == cnt_ref, (cnt_our, cnt_ref)
    assert f1_our == f1_ref
    # (mask logits within bf16 rounding of zero may land on either side of the 0.5 threshold)
    assert len(miou_our) == len(miou_ref) and np.allclose(miou_our, miou_ref, atol=1e-2)
    assert abs(np.mean(miou_our) - np.mean(miou_ref)) <= 3e-3
    assert np.allclose(iou_our, iou_ref, atol=5e-3)
    for a, b_ in zip(p_our, p_ref):   # pre-threshold top-300 lists
        same = (a["all_queries"] == b_["all_queries"]) & (a["all_labels"] == b_["all_labels"])
        assert same.float().mean() >= 0.97          # ties between near-equal scores may swap places
        assert (a["all_scores"] - b_["all_scores"]).abs().max() <= 2e-2
        assert torch.equal(a["labels"], b_["labels"]) and torch.equal(a["queries"], b_["queries"])


def test_no_out_of_bounds_writes(dev):
    """compute-sanitizer is closed on this pool, so writes are checked with canaries: every
    output / workspace of the C-ABI calls sits between guard zones that must stay intact."""
    import dfine_b200.ops as ops
    from dfine_b200 import _lib
    torch.manual_seed(3)
    B, Lq, H, c = 2, 37, 8, 32
    shapes, npts = [[13, 17], [7, 9], [4, 5]], [3, 6, 3]
    spec = ops.level_spec(shapes, npts)
    P = spec.P
    GUARD = 4096

    def guarded(nbytes, dtype):
        raw = torch.full((nbytes + 2 * GUARD,), 0xA5, dtype=torch.uint8, device=dev)
        view = raw[GUARD:GUARD + nbytes].view(dtype)
        return raw, view

    def intact(raw, nbytes):
        return bool((raw[:GUARD] == 0xA5).all() and (raw[GUARD + nbytes:] == 0xA5).all())

    mem = torch.randn(B, spec.L, H * c, device=dev).to(torch.bfloat16)
    rawin = torch.randn(B, Lq, 3 * H * P, device=dev).to(torch.bfloat16)
    ref = torch.cat([torch.rand(B, Lq, 2, device=dev) * 1.4 - 0.2, torch.rand(B, Lq, 2, device=dev)], -1)
    nps = torch.tensor([1.0 / n for n in npts for _ in range(n)], device=dev)
    go = torch.randn(B, Lq, H * c, device=dev)
    lib = _lib.lib()
    n_out, n_rec = B * Lq * H * c * 4, lib.dfine_msda_bwd_workspace_bytes(B, Lq, H, P)
    n_gv, n_graw = B * spec.L * H * c * 2, B * Lq * 3 * H * P * 2
    r_out, out = guarded(n_out, torch.float32)
    r_rec, rec = guarded(n_rec, torch.uint8)
    r_gv, gv = guarded(n_gv, torch.bfloat16)
    r_graw, graw = guarded(n_graw, torch.bfloat16)
    attn_view = rawin.reshape(-1)[2 * H * P:]
    s = torch.cuda.current_stream().cuda_stream
    rs = 3 * H * P
    rc = lib.dfine_msda_fwd(mem.data_ptr(), mem.stride(0), mem.stride(1), spec.hw_c, spec.start_c,
                            spec.npts_c, spec.n_lvl, rawin.data_ptr(), attn_view.data_ptr(), ref.data_ptr(),
                            nps.data_ptr(), 0.5, out.data_ptr(), None, B, Lq, H, c, _lib.BF16, _lib.BF16,
                            _lib.F32, _lib.MSDA_FUSED_INPUTS, rs, rs, rec.data_ptr(), s)
    _lib.check(rc, "fwd")
    flags = (_lib.MSDA_FUSED_INPUTS | _lib.MSDA_GRAD_VALUE_BF16 | _lib.MSDA_GRAD_SAMP_BF16 |
             _lib.MSDA_RECORDS_VALID)
    graw_attn = graw.reshape(-1)[2 * H * P:]
    rc = lib.dfine_msda_bwd(mem.data_ptr(), mem.stride(0), mem.stride(1), spec.hw_c, spec.start_c,
                            spec.npts_c, spec.n_lvl, rawin.data_ptr(), attn_view.data_ptr(), ref.data_ptr(),
                            nps.data_ptr(), 0.5, go.data_ptr(), gv.data_ptr(), graw.data_ptr(),
                            graw_attn.data_ptr(), B, Lq, H, c, _lib.BF16, _lib.BF16, _lib.F32, flags,
                            rs, rs, rs, rs, rec.data_ptr(), n_rec, s)
    _lib.check(rc, "bwd")
    torch.cuda.synchronize()
    assert intact(r_out, n_out) and intact(r_rec, n_rec) and intact(r_gv, n_gv) and intact(r_graw, n_graw)
    assert torch.isfinite(out).all() and torch.isfinite(gv.float()).all() and torch.isfinite(graw.float()).all()
    # FDR and mask GEMM outputs
    N = 61
    corners = torch.randn(N, 132, device=dev)
    refi = torch.rand(N, 4, device=dev)
    proj = ops.fdr_project(torch.tensor([0.5], device=dev), torch.tensor([4.0], device=dev))
    rsd = torch.tensor([4.0], device=dev)
    r_b, boxes = guarded(N * 16, torch.float32)
    r_gc, gc = guarded(N * 132 * 4, torch.float32)
    _lib.check(lib.dfine_fdr_fwd(corners.data_ptr(), _lib.F32, refi.data_ptr(), proj.data_ptr(), rsd.data_ptr(),
                                 None, boxes.data_ptr(), N, 32, s), "fdr_fwd")
    gb = torch.randn(N, 4, device=dev)
    _lib.check(lib.dfine_fdr_bwd(corners.data_ptr(), _lib.F32, refi.data_ptr(), proj.data_ptr(), rsd.data_ptr(),
                                 gb.data_ptr(), None, gc.data_ptr(), _lib.F32, N, 32, s), "fdr_bwd")
    Bm, M, K, Nn = 2, 77, 64, 264
    a = torch.randn(Bm, M, K, device=dev).to(torch.bfloat16)
    bmat = torch.randn(Bm, K, Nn, device=dev).to(torch.bfloat16)
    r_mo, mo = guarded(Bm * M * Nn * 2, torch.bfloat16)
    _lib.check(lib.dfine_mask_gemm_fwd(a.data_ptr(), bmat.data_ptr(), mo.data_ptr(), Bm, M, K, Nn, _lib.BF16, 0, s),
               "mask")
    torch.cuda.synchronize()
    assert intact(r_b, N * 16) and intact(r_gc, N * 132 * 4) and intact(r_mo, Bm * M * Nn * 2)
    want = torch.bmm(a.float(), bmat.float())
    assert rel_err(mo.view(Bm, M, Nn).float().cpu().numpy(), want.cpu().numpy()) <= BF16_RTOL


@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_mask_loss_kernel(xdt, dev):
    """dfine_mask_loss_fwd / _bwd (fused focal-BCE + dice over matched mask rows) against the reference's
    own losses and gradients (golden: DFINECriterion._focal_loss_mask / _dice_loss, dfine_criterion.py:273-312,
    binary / empty / full / soft targets, saturated logits) and against the oracle at the mask resolution of
    config 4 (rows of 160 x 160)."""
    import dfine_b200.ops as ops
    from oracle import cpu_oracle as O
    g = golden("mask_loss")
    # the kernel computes in fp32 on the logits as stored: for bf16 storage the reference values are
    # recomputed by the oracle on the rounded logits (the golden pins the oracle, tests/test_oracle_golden.py)
    pred = _t(g["pred"], dev).to(xdt)
    tgt = _t(g["tgt"], dev)
    o_bce, o_dice, o_gb, o_gd, o_stats = O.mask_loss(pred.float().cpu().numpy(), g["tgt"])
    for which, want_loss, want_grad in (("bce", o_bce, o_gb), ("dice", o_dice, o_gd)):
        p = pred.clone().requires_grad_(True)
        bce, dice = ops.mask_losses(p, tgt)
        loss = bce if which == "bce" else dice
        loss.backward()
        assert abs(float(loss) - want_loss) <= 1e-5 * abs(want_loss), (which, float(loss), want_loss)
        assert p.grad.dtype == xdt
        assert_close(p.grad.float().cpu().numpy(), want_grad, 1e-5 if xdt == torch.float32 else BF16_RTOL,
                     f"d loss_mask_{which} / d logits")
        if xdt == torch.float32:
            assert abs(float(loss) - float(g["loss_" + which])) <= 1e-5 * abs(float(g["loss_" + which]))
            assert_close(p.grad.cpu().numpy(), g["grad_" + which], 1e-5, f"golden grad_{which}")
    st = ops.mask_loss_stats(pred.flatten(1), tgt.flatten(1))
    assert_close(st.cpu().numpy(), o_stats, 1e-5, "row statistics")
    # config-4 rows: 37 matched masks of 160 x 160, a strided logits view, rectangle targets
    torch.manual_seed(5)
    M, N = 37, 160 * 160
    big = (torch.randn(M, N + 64, device=dev) * 4).to(xdt)
    x = big[:, :N]
    t = torch.zeros(M, 160, 160, device=dev)
    for m in range(M):
        t[m, 10 + m:60 + 2 * m, 20:40 + 3 * m] = 1
    t = t.flatten(1)
    xg = x.detach().clone().requires_grad_(True)     # (contiguous copy keeps .grad simple)
    bce, dice = ops.mask_losses(xg.view(M, 160, 160), t.view(M, 160, 160))
    (bce + 2 * dice).backward()
    ob, od, ogb, ogd, ost = O.mask_loss(x.float().cpu().numpy(), t.cpu().numpy())
    assert abs(float(bce) - ob) <= 1e-5 * ob and abs(float(dice) - od) <= 1e-5 * od
    assert_close(ops.mask_loss_stats(x, t).cpu().numpy(), ost, 1e-5, "row statistics (strided rows)")
    assert_close(xg.grad.float().cpu().numpy(), ogb + 2 * ogd, 1e-5 if xdt == torch.float32 else BF16_RTOL, "grad")
    with pytest.raises(ValueError):
        ops.mask_loss_stats(x, t[:, :100])
