import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import DiTBackbone, _lib, config as cfgmod
from layoutdit_b200.synth import make_state_dict, synthetic_pages
cfg = cfgmod.dit_base(); dev = torch.device("cuda", 0)
x = synthetic_pages(64, 224, 224, 1234).to(dev)
model = DiTBackbone(pretrained=False, config=cfg, state_dict=make_state_dict(cfg, 0, False)).to(dev).eval()
for _ in range(3): model(x)
torch.cuda.synchronize()
torch.cuda._sleep(200_000_000)
t0 = time.perf_counter()
for _ in range(10): model(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host enqueue time per forward: {(t1 - t0) / 10 * 1e3:.3f} ms")
eng = model._get_engine(); geo = eng._geometry(64, 224, 224); outs = eng._alloc_outputs(geo)
st = torch.cuda.current_stream().cuda_stream
torch.cuda._sleep(200_000_000)
t0 = time.perf_counter()
for _ in range(10): plan = eng._plan(geo, x, outs, st)
t1 = time.perf_counter()
for _ in range(10):
    for name, fn, args in plan: fn(*args)
t2 = time.perf_counter()
torch.cuda.synchronize()
print(f"plan build {(t1 - t0) / 10 * 1e3:.3f} ms; 91 ctypes launches {(t2 - t1) / 10 * 1e3:.3f} ms")
