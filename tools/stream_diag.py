import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import build_case
from layoutdit_b200 import DiTBackbone, _lib

cfg, sd, x, _, _ = build_case("tiny_abs_interp")
xe = x.cuda()
m = DiTBackbone(pretrained=False, config=cfg, state_dict=sd).cuda().eval()
eng = m._get_engine(); eng.refresh_weights()
xx = eng.prepare_input(xe)
geo = eng._geometry(1, 96, 64)
outs = eng._alloc_outputs(geo)
torch.cuda.synchronize()
side = torch.cuda.Stream()
print("side handle", hex(side.cuda_stream))
for use_side in [False, True]:
    st = side.cuda_stream if use_side else 0
    plan = eng._plan(geo, xx, outs, st)
    for i, (name, fn, args) in enumerate(plan[:8]):
        rc = fn(*args)
        try:
            torch.cuda.synchronize()
            print(use_side, i, name, "rc", rc, "ok")
        except Exception as e:
            print(use_side, i, name, "rc", rc, "FAILED", str(e)[:100])
            sys.exit(1)
