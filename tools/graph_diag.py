"""Bisect eager vs CUDA-graph replay: run the first `limit` library calls both ways and
compare every workspace buffer."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import build_case
from layoutdit_b200 import DiTBackbone

case = sys.argv[1] if len(sys.argv) > 1 else "tiny_abs_interp"
cfg, sd, x, _, _ = build_case(case)
m = DiTBackbone(pretrained=False, config=cfg, state_dict=sd).cuda().eval()
eng = m._get_engine()
eng.refresh_weights()
xx = eng.prepare_input(x.cuda())
geo = eng._geometry(xx.shape[0], xx.shape[2], xx.shape[3])
outs = eng._alloc_outputs(geo)
plan = eng._plan(geo, xx, outs, 0)
print("plan length", len(plan))
cur = torch.cuda.current_stream()

def snapshot():
    torch.cuda.synchronize()
    return [geo.x.clone(), geo.a.clone(), geo.big.clone()] + [o.clone() for o in outs]

def clear():
    geo.x.zero_(); geo.a.zero_(); geo.big.zero_()
    for o in outs: o.zero_()
    torch.cuda.synchronize()

names = ["x", "a", "big", "p2", "p3", "p4", "p5"]
for limit in list(range(1, len(plan) + 1)):
    clear()
    eng._enqueue(geo, xx, outs, cur.cuda_stream, limit)
    ref = snapshot()
    clear()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        eng._enqueue(geo, xx, outs, torch.cuda.current_stream().cuda_stream, limit)
    clear()
    g.replay()
    got = snapshot()
    bad = [n for n, r, t in zip(names, ref, got) if not torch.equal(r, t)]
    # repeat eager to see whether eager itself is stable
    clear()
    eng._enqueue(geo, xx, outs, cur.cuda_stream, limit)
    ref2 = snapshot()
    bad2 = [n for n, r, t in zip(names, ref, ref2) if not torch.equal(r, t)]
    print(limit, plan[limit - 1][0], "graph-vs-eager mismatch:", bad, " eager-vs-eager mismatch:", bad2)
    if bad or bad2:
        for n, r, t in zip(names, ref, got):
            if n in bad:
                d = (r.float() - t.float()).abs()
                print("   ", n, "max diff", float(d.max()), "n diff", int((d > 0).sum()), "of", d.numel(),
                      "first idx", (d.flatten() > 0).nonzero()[:5].flatten().tolist())
        break
