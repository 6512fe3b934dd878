import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import build_case
from layoutdit_b200 import DiTBackbone
from layoutdit_b200.engine import Engine

cfg, sd, x, _, _ = build_case("tiny_abs_interp")
xe = x.cuda()
orig_plan = Engine._plan
names = ["x", "a", "big"]
for limit in [1, 2, 3, 4, 5, 6, 9, 16, 47]:
    Engine._plan = lambda self, *a, _l=limit, **k: orig_plan(self, *a, **k)[:_l]
    m1 = DiTBackbone(pretrained=False, config=cfg, state_dict=sd).cuda().eval()
    m2 = DiTBackbone(pretrained=False, config=cfg, state_dict=sd, use_cuda_graph=True).cuda().eval()
    o1 = m1(xe); torch.cuda.synchronize()
    o2 = m2(xe); torch.cuda.synchronize()
    g1 = m1._engine._geoms[(1, 96, 64)]; g2 = m2._engine._geoms[(1, 96, 64)]
    bad = []
    for n in names:
        t1, t2 = getattr(g1, n), getattr(g2, n)
        if not torch.equal(t1, t2):
            d = (t1.float() - t2.float()).abs()
            bad.append((n, float(d.max()), int((d > 0).sum()), d.numel()))
    for k in o1:
        if not torch.equal(o1[k], o2[k]):
            bad.append((k, float((o1[k].float() - o2[k].float()).abs().max())))
    print(limit, "graph_in==x:", torch.equal(g2.graph_in, xe), "pos_bias eq:", torch.equal(g1.pos_bias, g2.pos_bias), bad)
