"""Breakdown of the data-parallel gradient exchange (run under torchrun, N >= 2): device time of the multicast add,
of torch's symmetric-memory barrier, and of both, each as 50 back-to-back repetitions inside one CUDA graph."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/d-fine-seg_b200")
from dfine_b200 import grad_sync, _lib
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
n = 288 * 256 + 288
ex = grad_sync.NvlsGradExchange([n] * 4, dev)
lib = _lib.lib()
def add():
    rc = lib.dfine_multicast_add(ex.local.data_ptr(), ex.mc_ptr, ex.total, ex.scale, torch.cuda.current_stream(dev).cuda_stream)
    assert rc == 0
def bar():
    ex.handle.barrier(channel=1)
def zero():
    ex.buf.zero_()
def timed(fn, reps=50):
    s = torch.cuda.Stream(dev); s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream(dev).wait_stream(s); torch.cuda.synchronize(dev); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize(dev); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize(dev)
    t = torch.tensor([a.elapsed_time(b) / reps * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del g
    return float(t)
res = {"multicast_add_us": timed(add), "barrier_us": timed(bar), "zero_us": timed(zero),
       "add_plus_barrier_us": timed(lambda: (add(), bar()))}
if rank == 0: print(world, res)
torch.cuda.synchronize(dev); dist.barrier(); dist.destroy_process_group()
