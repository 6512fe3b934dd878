"""In-process A/B of the attention variants inside the full forward (needs a -DLDIT_EXPERIMENTAL library):
one model, one captured graph per variant, alternating timed blocks.  usage: attn_ab_step.py [workload] [rounds]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS
from layoutdit_b200 import DiTBackbone, _lib, config as cfgmod
from layoutdit_b200.synth import make_state_dict, synthetic_pages
wl = sys.argv[1] if len(sys.argv) > 1 else "base224"
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 4
fac, B, H, W = WORKLOADS[wl]
cfg = getattr(cfgmod, fac)()
dev = torch.device("cuda", 0)
lib = _lib.load()
model = DiTBackbone(pretrained=False, config=cfg, state_dict=make_state_dict(cfg, 0, False)).to(dev).eval()
eng = model._get_engine()
x = synthetic_pages(B, H, W, 1234).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
variants = {"v3": (0, 0), "old": (4, 1)} if lib.ldit_has_experimental() else {"v3": (0, 0)}
for name, (impl, slot) in variants.items():
    lib.ldit_set_attention_impl(impl)
    eng.forward_graphed(x, slot)          # captures the graph of this slot with this variant
torch.cuda.synchronize()
res = {k: [] for k in variants}
for r in range(rounds):
    for name, (impl, slot) in variants.items():
        for _ in range(3): eng.forward_graphed(x, slot)
        ts = []
        for _ in range(20):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); eng.forward_graphed(x, slot); b.record()
            ts.append((a, b))
        torch.cuda.synchronize()
        res[name].append(sum(a.elapsed_time(b) for a, b in ts) / len(ts))
for k, v in res.items():
    print(f"{wl} attention {k}: " + " ".join(f"{t:.4f}" for t in v) + f"  | mean {sum(v)/len(v):.4f} ms/step")
