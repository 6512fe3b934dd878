import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import DiTBackbone, _lib, config as cfgmod
from layoutdit_b200.synth import make_state_dict, synthetic_pages
cfg = cfgmod.DiTConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, image_size=64)
dev = torch.device("cuda", 0)
model = DiTBackbone(pretrained=False, config=cfg, state_dict=make_state_dict(cfg, 0, False)).to(dev).eval()
eng = model._get_engine()
x = synthetic_pages(2, 64, 64, 1).to(dev)
model(x); torch.cuda.synchronize()
geo = eng._geometry(2, 64, 64); outs = eng._alloc_outputs(geo)
g = torch.cuda.CUDAGraph()
g.enable_debug_mode()
with torch.cuda.graph(g):
    eng._enqueue(geo, x, outs, torch.cuda.current_stream().cuda_stream)
g.debug_dump("gpurun_out/graph.dot")
print(open("gpurun_out/graph.dot").read()[:6000])
