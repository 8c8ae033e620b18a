"""CPU-side tests: the C-ABI symbol table, argument validation (no compute calls), the
zero-copy recovery of `memory` behind value_op's views, model patching, and the
world_size-2 (gloo) logic of bench.py."""
import copy
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


def test_cabi_exports_every_declared_symbol():
    import dfine_b200
    from dfine_b200 import build
    build.build()
    header = open(os.path.join(ROOT, "include", "dfine_b200.h")).read()
    declared = re.findall(r"^DFINE_API\s+[\w\s\*]+?\b(dfine_\w+)\s*\(", header, flags=re.M)
    assert len(declared) >= 9
    lib = ctypes.CDLL(dfine_b200.library_path())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/dfine_b200.h but not exported"
    assert sorted(declared) == dfine_b200._lib.exported_symbols()
    m = re.search(r"#define DFINE_B200_VERSION (\d+)", header)
    assert dfine_b200._lib.lib().dfine_version() == int(m.group(1)) >= 101


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call: negative return + message."""
    from dfine_b200 import _lib
    lib = _lib.lib()
    hw, st, npts = _lib.i32_array([4, 4]), _lib.i32_array([0]), _lib.i32_array([2])
    rc = lib.dfine_msda_fwd(None, 0, 0, hw, st, npts, 1, None, None, None, None, 0.5, None, None,
                            0, 1, 1, 32, 0, 0, 0, 0, 0, 0, None, None)
    assert rc == -2 and b"positive" in lib.dfine_last_error()
    rc = lib.dfine_msda_fwd(None, 0, 0, hw, st, npts, 9, None, None, None, None, 0.5, None, None,
                            1, 1, 1, 32, 0, 0, 0, 0, 0, 0, None, None)
    assert rc == -3 and b"n_lvl" in lib.dfine_last_error()
    rc = lib.dfine_msda_fwd(None, 0, 0, hw, st, _lib.i32_array([40]), 1, None, None, None, None, 0.5,
                            None, None, 1, 1, 1, 32, 0, 0, 0, 0, 0, 0, None, None)
    assert rc == -3 and b"sampling points" in lib.dfine_last_error()
    rc = lib.dfine_msda_fwd(None, 0, 0, hw, st, npts, 1, None, None, None, None, 0.5, None, None,
                            1, 1, 1, 32, 0, 0, 0, 0, 0, 0, None, None)
    assert rc == -1 and b"NULL" in lib.dfine_last_error()
    assert lib.dfine_mask_gemm_fwd(None, None, None, 1, 8, 100, 64, 1, 0, None) == -3
    assert lib.dfine_fdr_project(None, None, None, 31, None) == -2
    # dfine_linear_wgrad: shapes first (K > 256 is "use a library GEMM", the rest bad shapes), then pointers
    assert lib.dfine_linear_wgrad(None, 0, None, 0, 16, 288, 512, None, None) == -3
    assert b"K <= 256" in lib.dfine_last_error()
    assert lib.dfine_linear_wgrad(None, 0, None, 0, 16, 287, 256, None, None) == -2
    assert lib.dfine_linear_wgrad(None, 0, None, 0, 0, 288, 256, None, None) == -2
    assert lib.dfine_linear_wgrad(None, 100, None, 0, 16, 288, 256, None, None) == -2
    assert b"row strides" in lib.dfine_last_error()
    assert lib.dfine_linear_wgrad(None, 0, None, 0, 16, 288, 256, None, None) == -1
    with pytest.raises(_lib.DfineB200Error):
        _lib.check(rc, "probe")
    # the fused decoder-layer entry points: shapes / dtypes before pointers
    L = lib.dfine_linear_fwd
    assert L(None, 0, 0, None, 0, 0, None, None, 1, None, 1, 0, None, 0, 288, 256, 0, None) == -2          # M = 0
    assert L(None, 7, 0, None, 0, 0, None, None, 1, None, 1, 0, None, 16, 288, 256, 0, None) == -3         # dtype
    assert L(None, 0, 100, None, 0, 0, None, None, 1, None, 1, 0, None, 16, 288, 256, 0, None) == -2       # stride < K
    assert b"row strides" in lib.dfine_last_error()
    assert L(None, 0, 0, None, 0, 0, None, None, 1, None, 1, 0, None, 16, 288, 256, 0, None) == -1         # NULL x
    assert lib.dfine_gate_fwd(None, 0, None, 0, None, None, 1, None, None, 1e-5, None, 0, 16, 100, None) == -3
    assert b"multiple of 64" in lib.dfine_last_error()
    assert lib.dfine_gate_fwd(None, 0, None, 0, None, None, 1, None, None, 1e-5, None, 0, 16, 512, None) == -3
    assert lib.dfine_ffn_out_fwd(None, 0, None, None, 1, None, 0, None, None, 1e-5, None, 0, 16, 256, 100, None) == -1 \
        or lib.dfine_ffn_out_fwd(None, 0, None, None, 1, None, 0, None, None, 1e-5, None, 0, 16, 256, 100, None) == -3
    assert lib.dfine_lqe_fwd(None, 0, None, 0, None, None, None, None, None, 10, 80, 3, 64, 32, 0, None) == -3
    assert b"k = 4" in lib.dfine_last_error()
    assert lib.dfine_lqe_fwd(None, 0, None, 0, None, None, None, None, None, 10, 80, 4, 64, 32, 0, None) == -1
    assert lib.dfine_lqe_fwd(None, 0, None, 0, None, None, None, None, None, 0, 80, 4, 64, 32, 0, None) == 0   # empty


def test_parameter_caches_live_and_die_with_the_parameter():
    """bf16 copies of parameters (inference) are cached on the parameter OBJECT and invalidated by in-place
    modification; a new tensor at the same address, with the same shape and version, never sees them."""
    from dfine_b200 import ops
    p = torch.nn.Parameter(torch.randn(8, 8))
    t = ops.bf16_param(p)
    assert t.dtype == torch.bfloat16 and ops.bf16_param(p) is t
    with torch.no_grad():
        p.add_(1.0)
    t2 = ops.bf16_param(p)
    assert t2 is not t and torch.equal(t2, p.detach().bfloat16())
    q = copy.deepcopy(p)
    assert "_dfine_bf16" not in q.__dict__
    with torch.no_grad():
        q.mul_(2.0)
    assert torch.equal(ops.bf16_param(q), q.detach().bfloat16())
    b = torch.randn(4).bfloat16()
    assert ops.bf16_param(b) is b


def test_cpu_tensors_are_rejected():
    import dfine_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dfine_b200.fdr_integral(torch.zeros(3, 132), torch.zeros(33))
    m = dfine_b200.MSDeformableAttention(64, 4, 2, [2, 2])
    with pytest.raises(ValueError, match="Last dim of reference_points"):
        m(torch.zeros(1, 3, 64), torch.zeros(1, 3, 1, 3), torch.zeros(1, 20, 4, 16), [[4, 4], [2, 2]])
    with pytest.raises(AssertionError):
        dfine_b200.MSDeformableAttention(64, 4, 3, [2, 2])


def test_memory_recovered_zero_copy_from_value_views():
    from dfine_b200 import ops
    B, H, c = 2, 4, 16
    shapes, npts = [[3, 5], [2, 2]], [2, 1]
    spec = ops.level_spec(shapes, npts)
    assert (spec.L, spec.P, spec.starts) == (19, 3, [0, 15])
    mem = torch.randn(B, spec.L, H * c, requires_grad=True)
    views = mem.reshape(B, spec.L, H, c).permute(0, 2, 3, 1).split(spec.sizes, dim=-1)
    base, h2, c2, zero_copy = ops.memory_from_value(views, spec)
    assert zero_copy and base is mem and (h2, c2) == (H, c)
    # anything else is packed into the same layout
    clones = tuple(v.clone() for v in views)
    packed, _, _, zero_copy = ops.memory_from_value(clones, spec)
    assert not zero_copy and torch.equal(packed, mem.detach())
    sliced = torch.randn(B, spec.L + 3, H * c)[:, :spec.L]
    v2 = sliced.reshape(B, spec.L, H, c).permute(0, 2, 3, 1).split(spec.sizes, dim=-1)
    packed, _, _, zero_copy = ops.memory_from_value(v2, spec)
    assert not zero_copy and torch.equal(packed, sliced)
    with pytest.raises(ValueError):
        ops.memory_from_value(views[:1], spec)


def test_module_mirror_has_reference_parameters():
    import dfine_b200
    m = dfine_b200.MSDeformableAttention(256, 8, 3, [3, 6, 3])
    sd = m.state_dict()
    assert sorted(sd) == ["attention_weights.bias", "attention_weights.weight", "num_points_scale",
                          "sampling_offsets.bias", "sampling_offsets.weight"]
    assert sd["sampling_offsets.weight"].shape == (8 * 12 * 2, 256)
    assert sd["attention_weights.weight"].shape == (8 * 12, 256)
    assert torch.allclose(sd["num_points_scale"], torch.tensor([1 / 3] * 3 + [1 / 6] * 6 + [1 / 3] * 3))
    assert float(sd["sampling_offsets.weight"].abs().max()) == 0.0
    if os.path.isdir(REFERENCE):
        sys.path.insert(0, REFERENCE)
        from src.d_fine.arch.dfine_decoder import MSDeformableAttention as Ref
        torch.manual_seed(0)
        r = Ref(256, 8, 3, [3, 6, 3])
        for k, v in r.state_dict().items():
            assert torch.allclose(sd[k], v, atol=1e-6), k
        m.load_state_dict(r.state_dict(), strict=True)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not present")
def test_patch_model_keeps_reference_checkpoint_format():
    import dfine_b200
    sys.path.insert(0, REFERENCE)
    from src.d_fine.dfine import build_model
    torch.manual_seed(0)
    model = build_model("n", 80, True, "cpu", img_size=[640, 640])
    keys = list(model.state_dict().keys())
    n = dfine_b200.patch_model(model)
    assert n["msda"] == 3 and n["integral"] == 1 and n["mask"] == 1
    assert list(model.state_dict().keys()) == keys
    assert "decoder.decoder.layers.0.cross_attn.num_points_scale" in keys
    layer = model.decoder.decoder.layers[0].cross_attn
    assert layer.ms_deformable_attn_core.func is dfine_b200.ops.msda_core
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.eval()(torch.rand(1, 3, 640, 640))  # the patched path refuses CPU tensors
    dfine_b200.unpatch_model(model)
    with torch.no_grad():
        out = model.eval()(torch.rand(1, 3, 640, 640))  # reference path restored
    assert out["pred_boxes"].shape == (1, 300, 4)
    import copy
    dfine_b200.patch_model(model)
    copy.deepcopy(model)  # EMA deep-copies the model (reference src/dl/train.py:56)


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms = bench.max_over_ranks(10.0 * (rank + 1), torch.device("cpu"), True)
    wl = dict(bench.WORKLOADS[bench.DEFAULT_WORKLOAD], B=1, Lq=3, layers=1, shapes=[[2, 2]], npts=[1])
    inp = bench.make_inputs(wl, 1, bench.rank_seed(rank), "cpu")
    q.put((rank, ms, float(inp["memory"].sum()), bench.job_throughput(32, world, 4, ms)))
    dist.destroy_process_group()


def test_bench_multi_rank_logic_gloo():
    """world_size 2 over gloo: ranks draw different shards, the step time is the max over
    ranks and the reported throughput is the whole-job aggregate."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, ms0, sum0, thr0), (r1, ms1, sum1, thr1) = res
    assert ms0 == ms1 == 20.0
    assert sum0 != sum1
    assert thr0 == thr1 == 32 * 2 * 4 / 0.020


def _bucket_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "d-fine-seg_b200"))
    from dfine_b200 import grad_sync
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                       # same parameters on every rank
    mods = [torch.nn.ModuleDict(dict(sampling_offsets=torch.nn.Linear(8, 12),
                                     attention_weights=torch.nn.Linear(8, 6))) for _ in range(2)]
    params = grad_sync.path_parameters(mods)
    bucket = grad_sync.GradBucket(params)
    g = torch.Generator().manual_seed(100 + rank)   # different shards -> different gradients
    local = [torch.randn(p.shape, generator=g) for p in params]
    for p, t in zip(params, local):
        p.grad = t.clone()
    flat = bucket.reduce()
    # explicit gradient list (what a CUDA-graph replay hands over) gives the same result
    again = [t.clone() for t in local]
    bucket.reduce(again)
    q.put((rank, [t.numpy() for t in local], [p.grad.numpy() for p in params],
           [t.numpy() for t in again], flat.numel(), bucket.nbytes))
    # a rank that lacks a gradient must fail loudly, before any collective is entered
    params[0].grad = None
    try:
        bucket.reduce()
        q.put((rank, "no error"))
    except RuntimeError as exc:
        q.put((rank, str(exc)))
    dist.destroy_process_group()


def test_grad_bucket_averages_linear_gradients_gloo():
    """world_size 2 over gloo: the flat bucket all-reduce leaves the rank average of every
    Linear gradient in param.grad on both ranks (DDP's step for the path's own parameters)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=120) for _ in range(4)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    res = sorted((g for g in got if len(g) == 6), key=lambda g: g[0])
    errs = [g for g in got if len(g) == 2]
    assert len(res) == 2 and len(errs) == 2
    assert all("no gradient" in e[1] for e in errs)
    (_, l0, r0, a0, n0, b0), (_, l1, r1, a1, n1, b1) = res
    assert n0 == n1 == 2 * (8 * 12 + 12 + 8 * 6 + 6) and b0 == 4 * n0
    for x0, x1, y0, y1, z0, z1 in zip(l0, l1, r0, r1, a0, a1):
        want = (x0 + x1) / 2
        np.testing.assert_allclose(y0, want, rtol=0, atol=1e-7)
        np.testing.assert_array_equal(y0, y1)
        np.testing.assert_array_equal(z0, y0)
        np.testing.assert_array_equal(z1, y1)


def _layerwise_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "d-fine-seg_b200"))
    from dfine_b200 import grad_sync
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                       # same parameters on every rank
    mods = [torch.nn.ModuleDict(dict(sampling_offsets=torch.nn.Linear(8, 12),
                                     attention_weights=torch.nn.Linear(8, 6))) for _ in range(3)]
    sync = grad_sync.LayerwiseGradSync(mods)
    g = torch.Generator().manual_seed(100 + rank)   # different shards -> different gradients
    x = torch.randn(5, 8, generator=g)

    def loss(ms):
        return sum((m["sampling_offsets"](x).square().sum() + m["attention_weights"](x).sum()) * (i + 1)
                   for i, m in enumerate(ms))

    local = torch.autograd.grad(loss(mods), grad_sync.path_parameters(mods))   # no hooks fire here
    out = []
    for step in range(2):                      # two passes: the pending counters re-arm in finish()
        for m in mods:
            m.zero_grad(set_to_none=True)
        loss(mods).backward()                  # hooks: one all-reduce per module, last layer first
        sync.finish()
        out.append([p.grad.numpy().copy() for p in grad_sync.path_parameters(mods)])
    q.put((rank, [t.numpy() for t in local], out, sync.launched, sync.nbytes))
    # a bucket that did not get all its gradients is reported at the join
    for m in mods:
        m.zero_grad(set_to_none=True)
    mods[0]["sampling_offsets"](x).sum().backward()
    try:
        sync.finish()
        q.put((rank, "no error"))
    except RuntimeError as exc:
        q.put((rank, str(exc)))
    dist.destroy_process_group()


def test_layerwise_grad_sync_gloo():
    """world_size 2 over gloo: per-module buckets reduced from autograd hooks leave the rank average in
    every param.grad, on every step; an incomplete bucket fails loudly at finish()."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_layerwise_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=120) for _ in range(4)]
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    res = sorted((g for g in got if len(g) == 5), key=lambda g: g[0])
    errs = [g for g in got if len(g) == 2]
    assert len(res) == 2 and len(errs) == 2
    assert all("did not receive" in e[1] for e in errs)
    (_, l0, o0, n0, b0), (_, l1, o1, n1, b1) = res
    assert n0 == n1 == 6 and b0 == b1 == 3 * 4 * (8 * 12 + 12 + 8 * 6 + 6)
    for step in range(2):
        for x0, x1, y0, y1 in zip(l0, l1, o0[step], o1[step]):
            np.testing.assert_allclose(y0, (x0 + x1) / 2, rtol=1e-6, atol=1e-6)
            np.testing.assert_array_equal(y0, y1)


def test_backward_point_limit_is_refused_before_any_work():
    """The forward takes up to 32 sampling points per head, the backward kernels 16: a differentiated forward with
    17-32 points must fail when it allocates its records, not at loss.backward() (no GPU needed: the check
    precedes the library call)."""
    from dfine_b200 import ops
    spec = ops.level_spec([[8, 8], [4, 4]], [9, 9])
    with pytest.raises(ValueError, match="at most 16"):
        ops.new_records(torch.zeros(1, 80, 256), spec, 8, 10)


def test_reference_points_last_dim_2_branch_is_dead_in_the_reference_and_fails_alike():
    """MSDeformableAttention.forward's `reference_points.shape[-1] == 2` branch (reference dfine_decoder.py:149-155,
    inherited from RT-DETR) divides the [bs, Lq, H, P, 2] offsets by a [1, 1, 1, n_levels, 1, 2] normaliser: the
    broadcast fails for every head / level count (probed: RuntimeError "The size of tensor a ... must match"), so
    no configuration can reach the sampling core through it.  The mirror keeps the branch and its error behaviour
    (same exception type, raised by the same arithmetic, before any kernel is launched); last dim 3 raises the
    reference's ValueError."""
    import dfine_b200
    H, shapes, npts = 8, [[8, 8], [4, 4], [2, 2]], [3, 6, 3]
    C, B, Lq = 128, 2, 5
    L = sum(h * w for h, w in shapes)
    mem, q = torch.randn(B, L, C), torch.randn(B, Lq, C)
    value = mem.reshape(B, L, H, C // H).permute(0, 2, 3, 1).split([h * w for h, w in shapes], dim=-1)
    ref2 = torch.rand(B, Lq, len(shapes), 2)
    mods = [dfine_b200.MSDeformableAttention(C, H, len(shapes), npts)]
    try:
        from baseline import ref_install
        if ref_install.installed():
            ref_install.import_reference()
            from src.d_fine.arch.dfine_decoder import MSDeformableAttention as RefMSDA
            mods.append(RefMSDA(C, H, len(shapes), npts))
    except ImportError:
        pass
    for m in mods:
        with pytest.raises(RuntimeError, match="must match the size"):
            m(q, ref2, value, shapes)
        with pytest.raises(ValueError, match="Last dim of reference_points must be 2 or 4"):
            m(q, torch.rand(B, Lq, 1, 3), value, shapes)
