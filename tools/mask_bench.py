import sys, json, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/d-fine-seg_b200")
import bench
dev = torch.device("cuda:0")
r = bench.secondary_kernel_legs(dev, 6531.6)
print(json.dumps(r["config4_mask_assembly_bf16"], indent=1))
