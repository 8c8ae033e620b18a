import copy, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "d-fine-seg_b200"))
from loguru import logger
logger.remove()
from baseline import model_harness as H
import dfine_b200
dev = torch.device("cuda:0")
for name, seg in (("s", False), ("m", True)):
    model, loss_fn = H.build(name, dev, 640, seg)
    model.train()
    patched = copy.deepcopy(model); dfine_b200.patch_model(patched)
    images, targets = H.synthetic_batch(2, 640, dev, seed=42, seg=seg)
    # capture decoder inputs via a forward pre-hook on the decoder (cheap: clone only)
    for seed in range(5):
        for tag, m in (("ref", model), ("patched", patched)):
            cap = {}
            h1 = m.encoder.register_forward_hook(lambda mod, i, o: cap.__setitem__("enc", o))
            torch.manual_seed(seed)
            with torch.autocast("cuda", dtype=torch.float16):
                out = m(images, targets=targets)
            h1.remove()
            torch.cuda.synchronize()
            flat = H.flat_outputs(out)
            bad = sorted(k for k, v in flat.items() if v.is_floating_point() and not torch.isfinite(v).all())
            enc = cap["enc"]
            enc_list = enc[0] if isinstance(enc, tuple) and isinstance(enc[0], (list, tuple)) else enc
            encbad = [bool(torch.isfinite(t).all()) for t in (enc_list if isinstance(enc_list, (list, tuple)) else [enc_list]) if isinstance(t, torch.Tensor)]
            print(name, seed, tag, "enc finite:", encbad, "n_bad", len(bad), [b for b in bad if "enc" in b or "pre" in b][:6], bad[:3], flush=True)
