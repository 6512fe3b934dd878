"""Whole-forward device time, eager stream vs CUDA graph, with PDL on / off (launches queued behind a sleep)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import DiTBackbone, _lib, config as cfgmod
from layoutdit_b200.synth import make_state_dict, synthetic_pages
cfg = cfgmod.dit_base(); dev = torch.device("cuda", 0)
lib = _lib.load()
x = synthetic_pages(64, 224, 224, 1234).to(dev)
for pdl in (0, 1, 0, 1):
    lib.ldit_set_pdl(pdl)
    model = DiTBackbone(pretrained=False, config=cfg, state_dict=make_state_dict(cfg, 0, False)).to(dev).eval()
    eng = model._get_engine()
    model(x); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        torch.cuda._sleep(30_000_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); model(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    eager = sorted(ts)[2]
    model.use_cuda_graph = True
    model(x); model(x); torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); model(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"pdl={pdl}: eager {eager:.3f} ms, graph {sorted(ts)[4]:.3f} ms")
