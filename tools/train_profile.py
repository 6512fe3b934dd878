"""Kernel-time table of one training step of the backbone (TrainableBackbone forward + loss + backward).
usage: train_profile.py [batch side]   (default 64 224; 32 512 = the base512 geometry, flash attention backward)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import config as cfgmod
from layoutdit_b200.dit_params import DiTParameters
from layoutdit_b200.train import TrainableBackbone
from layoutdit_b200.synth import make_state_dict, synthetic_pages
from torch.profiler import profile, ProfilerActivity

cfg = cfgmod.dit_base()
dev = torch.device("cuda", 0)
tree = DiTParameters(cfg)
tree.load_state_dict(make_state_dict(cfg, 0, False), strict=False)
tree = tree.to(dev)
model = TrainableBackbone(tree, cfg)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 224
pages = synthetic_pages(B, S, S, 1).to(dev)
def step():
    for p in tree.parameters(): p.grad = None
    sum(f.square().mean(dtype=torch.float32) for f in model(pages).values()).backward()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
