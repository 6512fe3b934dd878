"""A few DiTWithFPN forwards at base224 (for the ncu launch list of the FPN head)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import DiTWithFPN, config as cfgmod
from layoutdit_b200.synth import make_fpn_state_dict, make_state_dict, synthetic_pages
cfg = cfgmod.dit_base()
m = DiTWithFPN(pretrained=False, config=cfg, state_dict=make_state_dict(cfg, 0, False),
               fpn_state_dict=make_fpn_state_dict(768, 256, 0, False)).cuda().eval()
x = synthetic_pages(64, 224, 224, 1234).cuda()
for _ in range(3):
    m(x)
torch.cuda.synchronize()
print("ok")
