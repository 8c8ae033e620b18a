import sys, json, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/d-fine-seg_b200")
import bench
dev = torch.device("cuda:0")
timed = bench.graph_timed(dev)
r = bench.decoder_layer_legs(dev, 6531.6, timed)
for k, v in r.items():
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a in ("ms", "frac", "tflops", "reference_ops_ms", "vs_reference_ops", "two_kernel_route_ms")})
