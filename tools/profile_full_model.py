"""Where does the full-model train step go?  torch.profiler over a few steps of the reference model
(patched / unpatched): GPU busy time, top kernels, host time per phase."""
import copy, os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-fine-seg_b200"))
from loguru import logger
logger.remove()
from baseline import model_harness as H
import dfine_b200
dev = torch.device("cuda:0")
name, B = (sys.argv[1] if len(sys.argv) > 1 else "m"), int(sys.argv[2]) if len(sys.argv) > 2 else 32
patch = (sys.argv[3] if len(sys.argv) > 3 else "patched") == "patched"
model, loss_fn = H.build(name, dev, 640, False)
model.train(); loss_fn.train()
if patch:
    dfine_b200.patch_model(model)
opt = H.build_optimizer(model, name)
images, targets = H.synthetic_batch(B, 640, dev, seed=42)

def phase_times():
    t = {}
    def tic(): torch.cuda.synchronize(); return time.perf_counter()
    t0 = tic()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        x = model.backbone(images); t1 = tic()
        x = model.encoder(x); t2 = tic()
        out = model.decoder(x, targets); t3 = tic()
    with torch.autocast("cuda", enabled=False):
        ld = loss_fn(out, targets)
    loss = sum(ld.values()); t4 = tic()
    loss.backward(); t5 = tic()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1); opt.step(); opt.zero_grad(); t6 = tic()
    return dict(backbone=t1-t0, encoder=t2-t1, decoder=t3-t2, criterion=t4-t3, backward=t5-t4, optimizer=t6-t5, total=t6-t0)

for _ in range(3):
    H.train_step(model, loss_fn, images, targets, torch.bfloat16, optimizer=opt)
print("phases (ms, synchronised):", {k: round(v*1e3, 2) for k, v in phase_times().items()}, flush=True)
print("phases (ms, synchronised):", {k: round(v*1e3, 2) for k, v in phase_times().items()}, flush=True)
from torch.profiler import profile, ProfilerActivity
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        H.train_step(model, loss_fn, images, targets, torch.bfloat16, optimizer=opt)
    torch.cuda.synchronize()
ev = prof.key_averages()
tot_cuda = sum(e.self_device_time_total for e in ev) / 2 / 1e3
print("GPU kernel time per step (ms):", round(tot_cuda, 2))
rows = sorted(ev, key=lambda e: -e.self_device_time_total)[:40]
for e in rows:
    print(f"{e.self_device_time_total/2/1e3:9.3f} ms  x{e.count//2:5d}  {e.key[:110]}")
print("--- top CPU self time")
rows = sorted(ev, key=lambda e: -e.self_cpu_time_total)[:25]
for e in rows:
    print(f"{e.self_cpu_time_total/2/1e3:9.3f} ms  x{e.count//2:5d}  {e.key[:110]}")
n_launch = sum(e.count for e in ev if e.self_device_time_total > 0 and e.device_type is not None) 
print("events with device time per step:", sum(e.count for e in ev if e.self_device_time_total > 0)//2)
