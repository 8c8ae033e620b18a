"""Pins the CPU oracle (oracle/dfine_oracle.c) to outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by oracle/make_golden.py, which imports
the reference (src/d_fine/arch/{utils,dfine_decoder}.py) and runs it on seeded inputs.
"""
import numpy as np
import pytest

from oracle import cpu_oracle as O
from util import FP32_RTOL, assert_close, bf16_bits_to_f32, golden

CORE_CASES = ["core_m_small", "core_n_small", "core_x444", "core_edge", "core_near_centre"]


def _core_inputs(g):
    H, c = int(g["H"]), int(g["c"])
    mem = bf16_bits_to_f32(g["memory_bf16"])
    B, L, C = mem.shape
    return mem.reshape(B, L, H, c), bf16_bits_to_f32(g["grad_out_bf16"])


@pytest.mark.parametrize("name", CORE_CASES)
def test_core_forward_matches_reference(name):
    # reference: deformable_attention_core_func_v2, arch/utils.py:191-264
    g = golden(name)
    value, _ = _core_inputs(g)
    out = O.msda_fwd(value, g["shapes"], g["npts"], g["loc"], g["attn"])
    assert_close(out, g["out"], FP32_RTOL, "out")


@pytest.mark.parametrize("name", CORE_CASES)
def test_core_backward_matches_reference(name):
    g = golden(name)
    value, go = _core_inputs(g)
    gv, gl, ga = O.msda_bwd(value, g["shapes"], g["npts"], g["loc"], g["attn"], go)
    assert_close(gv.reshape(g["grad_memory"].shape), g["grad_memory"], FP32_RTOL, "grad_value")
    assert_close(gl, g["grad_loc"], FP32_RTOL, "grad_loc")
    assert_close(ga, g["grad_attn"], FP32_RTOL, "grad_attn")


def test_corner_indices_match_aten_probe():
    """The pixels aten::grid_sampler_2d touches (non-zeros of grad_input for an all-ones
    map, one sample per image) must be exactly the oracle's in-bounds corners, with the
    same weights.  Zero-weight corners are invisible to the probe and are skipped."""
    g = golden("aten_corner_probe")
    h, w = [int(v) for v in g["hw"]]
    loc = g["loc"]
    n = loc.shape[0]
    value = np.ones((1, h * w, 1, 1), np.float32)
    attn = np.ones((1, n, 1, 1), np.float32)
    out, idx, wts = O.msda_fwd(value, [[h, w]], [1], loc.reshape(1, n, 1, 1, 2), attn, want_idx=True)
    idx, wts = idx.reshape(n, 4), wts.reshape(n, 4)
    dense = np.zeros((n, h * w), np.float32)
    for j in range(4):
        ok = idx[:, j] >= 0
        np.add.at(dense, (np.nonzero(ok)[0], idx[ok, j]), wts[ok, j])
    ref = g["grad_input"]
    assert np.array_equal(dense != 0, ref != 0), "touched-pixel sets differ from ATen"
    assert np.abs(dense - ref).max() <= 1e-6
    assert_close(out.reshape(n), g["out"], FP32_RTOL, "probe out")


@pytest.mark.parametrize("name", ["module_m_small", "module_n_small"])
def test_fused_module_matches_reference(name):
    # reference: MSDeformableAttention.forward, dfine_decoder.py:119-178
    g = golden(name)
    H = int(g["H"])
    mem = bf16_bits_to_f32(g["memory_bf16"])
    B, L, C = mem.shape
    value = mem.reshape(B, L, H, C // H)
    out = O.msda_fused_fwd(value, g["shapes"], g["npts"], g["raw_off"], g["raw_logit"],
                           g["ref_points"], g["num_points_scale"], float(g["offset_scale"]))
    assert_close(out, g["out"], FP32_RTOL, "module out")
    go = bf16_bits_to_f32(g["grad_out_bf16"])
    gv, g_off, g_logit = O.msda_fused_bwd(value, g["shapes"], g["npts"], g["raw_off"],
                                          g["raw_logit"], g["ref_points"],
                                          g["num_points_scale"], go, float(g["offset_scale"]))
    assert_close(gv.reshape(B, L, C), g["grad_memory"], FP32_RTOL, "grad_memory")
    # push the raw-output gradients through the two Linears (plain matmuls) and compare
    # with the parameter / query gradients autograd produced in the reference
    q = g["query"].reshape(-1, C).astype(np.float64)
    g_off2 = g_off.reshape(q.shape[0], -1).astype(np.float64)
    g_log2 = g_logit.reshape(q.shape[0], -1).astype(np.float64)
    assert_close(g_off2.T @ q, g["g_so_w"], 2e-5, "grad sampling_offsets.weight")
    assert_close(g_off2.sum(0), g["g_so_b"], 2e-5, "grad sampling_offsets.bias")
    assert_close(g_log2.T @ q, g["g_aw_w"], 2e-5, "grad attention_weights.weight")
    assert_close(g_log2.sum(0), g["g_aw_b"], 2e-5, "grad attention_weights.bias")
    gq = g_off2 @ g["so_w"].astype(np.float64) + g_log2 @ g["aw_w"].astype(np.float64)
    assert_close(gq.reshape(g["grad_query"].shape), g["grad_query"], 2e-5, "grad_query")


@pytest.mark.parametrize("tag", ["m", "x", "odd"])
def test_weighting_function(tag):
    # reference: weighting_function, arch/utils.py:145-188 (both deploy modes)
    g = golden("fdr")
    up, rs = [float(v) for v in g[f"up_rs_{tag}"]]
    proj = O.fdr_project(up, rs, 32)
    assert_close(proj, g[f"project_{tag}"], FP32_RTOL, "project")
    assert_close(proj, g[f"project_deploy_{tag}"], FP32_RTOL, "project deploy")
    assert proj[16] == 0.0 and proj.shape == (33,)


def test_fdr_forward_backward():
    # reference: Integral.forward dfine_decoder.py:291-295, distance2bbox arch/utils.py:119-142
    g = golden("fdr")
    dist, boxes = O.fdr_fwd(g["corners"], g["ref_init"], g["project"], float(g["reg_scale"]))
    assert_close(dist, g["dist"], FP32_RTOL, "dist")
    assert_close(boxes, g["boxes"], FP32_RTOL, "boxes")
    gc = O.fdr_bwd(g["corners"], g["ref_init"], g["project"], float(g["reg_scale"]), g["grad_boxes"])
    assert_close(gc.reshape(g["corners"].shape), g["grad_corners_from_boxes"], FP32_RTOL, "gc boxes")
    gc = O.fdr_bwd(g["corners"], g["ref_init"], g["project"], float(g["reg_scale"]),
                   g["grad_boxes"], g["grad_dist"])
    assert_close(gc.reshape(g["corners"].shape), g["grad_corners_from_both"], FP32_RTOL, "gc both")


def test_mask_assembly():
    # reference: DFINETransformer._mask_logits_from_h, dfine_decoder.py:937-940 (+ :1041)
    g = golden("mask")
    assert_close(O.mask_gemm(g["coef"], g["proto"]), g["logits"], FP32_RTOL, "logits")
    assert_close(O.mask_gemm(g["coef"], g["proto"], True), g["probs"], FP32_RTOL, "probs")
    # backward of the contraction (autograd of dfine_decoder.py:940)
    g = golden("mask_bwd")
    assert_close(O.mask_gemm(g["coef"], g["proto"]), g["logits"], FP32_RTOL, "logits (K=128)")
    gc, gp = O.mask_gemm_bwd(g["coef"], g["proto"], g["grad_out"])
    assert_close(gc, g["grad_coef"], FP32_RTOL, "grad_coef")
    assert_close(gp, g["grad_proto"], FP32_RTOL, "grad_proto")


def test_linear_wgrad_restatement_matches_torch_autograd():
    """oracle.linear_wgrad (checker of dfine_linear_wgrad) == autograd of torch.nn.functional.linear."""
    import torch
    from oracle import cpu_oracle as O
    torch.manual_seed(0)
    x = torch.randn(37, 24, dtype=torch.float64)
    w = torch.randn(16, 24, dtype=torch.float64, requires_grad=True)
    b = torch.randn(16, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(37, 16, dtype=torch.float64)
    torch.nn.functional.linear(x, w, b).backward(gy)
    dw, db = O.linear_wgrad(gy.numpy(), x.numpy())
    np.testing.assert_allclose(dw, w.grad.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(db, b.grad.numpy(), rtol=1e-12, atol=1e-12)


def test_mask_loss():
    # reference: DFINECriterion._focal_loss_mask / _dice_loss, dfine_criterion.py:273-312
    g = golden("mask_loss")
    bce, dice, g_bce, g_dice, _ = O.mask_loss(g["pred"], g["tgt"])
    assert abs(bce - float(g["loss_bce"])) <= FP32_RTOL * abs(float(g["loss_bce"]))
    assert abs(dice - float(g["loss_dice"])) <= FP32_RTOL * abs(float(g["loss_dice"]))
    assert_close(g_bce, g["grad_bce"], FP32_RTOL, "d loss_mask_bce / d logits")
    assert_close(g_dice, g["grad_dice"], FP32_RTOL, "d loss_mask_dice / d logits")


def _layer_golden():
    g = golden("layer")
    d = {k: g[k] for k in g.files if not k.endswith("_bf16")}
    d.update({k[:-5]: bf16_bits_to_f32(g[k]) for k in g.files if k.endswith("_bf16")})
    return d


def _one_bf16_ulp(got, want, what, min_identical=0.999):
    """Linear outputs are bf16 values: the summation order may move one by a single ulp (2^-8 relative)."""
    rms = float(np.sqrt(np.mean(np.square(want, dtype=np.float64))))
    bad = np.abs(got - want) > 2.0 ** -7 * np.abs(want) + 2.0 ** -8 * rms
    assert not bad.any(), f"{what}: {int(bad.sum())} elements beyond one bf16 ulp"
    same = float((got == want).mean())
    assert same >= min_identical, f"{what}: only {same:.5f} bit-identical"


def test_layer_linears_match_reference_autocast():
    # reference: MSDeformableAttention's Linears on with_pos_embed(target, pos) (dfine_decoder.py:139-147, :245),
    # linear1 + ReLU (:229-230), torch CPU autocast(bfloat16)
    g = _layer_golden()
    raw = O.linear_bf16(g["target"], np.concatenate([g["so_w"], g["aw_w"]]), np.concatenate([g["so_b"], g["aw_b"]]),
                        x_add=g["pos"])
    _one_bf16_ulp(raw, g["raw"], "packed Linear")
    _one_bf16_ulp(O.linear_bf16(g["target"], g["w1"], g["b1"], relu=True), g["hidden"], "linear1 + relu")


def test_gate_and_ffn_tail_match_reference_autocast():
    # reference: Gate.forward (dfine_decoder.py:265-271), TransformerDecoderLayer.forward :251-253
    g = _layer_golden()
    got = O.gate_fwd(g["target"], g["x2"], g["gate_w"], g["gate_b"], g["gate_ln_w"], g["gate_ln_b"], float(g["gate_eps"]))
    # a gate that lands one bf16 ulp away (exp / summation order) moves its row by ~2^-8 of the row's scale
    assert_close(got, g["gate_out"], 1e-3, "gate")
    assert (np.abs(got - g["gate_out"]) <= 1e-5 * np.abs(g["gate_out"]).max()).mean() >= 0.999
    got = O.ffn_tail(g["hidden"], g["w2"], g["b2"], g["target"], g["ln3_w"], g["ln3_b"], float(g["ln3_eps"]))
    assert_close(got, g["ffn_out"], FP32_RTOL, "ffn tail")       # row 0 went through the clamp of :253


def test_lqe_matches_reference():
    # reference: LQE.forward, dfine_decoder.py:307-313 (float32, and CPU autocast(bfloat16))
    g = golden("lqe")
    args = (g["scores"], g["corners"], g["w1"], g["b1"], g["w2"], g["b2"])
    assert_close(O.lqe_fwd(*args), g["out_f32"], FP32_RTOL, "lqe float32")
    # torch's CPU autocast keeps softmax / topk / mean in bf16 (CUDA autocast runs softmax in float32, which is what
    # the oracle's emulate_bf16 and the kernel restate): the CPU fixture pins the bf16 route to the north star's
    # bf16 tolerance only; the CUDA policy is checked against the reference's ops on the GPU itself
    # (tests/test_gpu_layer.py::test_lqe_fwd_matches_reference_ops)
    assert_close(O.lqe_fwd(*args, emulate_bf16=True), g["out_bf16"], 1e-2, "lqe autocast (CPU policy)")
