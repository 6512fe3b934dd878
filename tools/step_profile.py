"""In-situ duration of every launch of one backbone forward (warm L2, real predecessor/successor):
CUDA events around each library call, enqueued behind a long sleep kernel so the host is never
the bottleneck.  usage: step_profile.py [workload] [reps]"""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS
from layoutdit_b200 import DiTBackbone, _lib, config as cfgmod
from layoutdit_b200.synth import make_state_dict, synthetic_pages

wl = sys.argv[1] if len(sys.argv) > 1 else "base224"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
fac, B, H, W = WORKLOADS[wl]
cfg = getattr(cfgmod, fac)()
dev = torch.device("cuda", 0)
model = DiTBackbone(pretrained=False, config=cfg, state_dict=make_state_dict(cfg, 0, False)).to(dev).eval()
eng = model._get_engine()
x = synthetic_pages(B, H, W, 1234).to(dev)
model(x); torch.cuda.synchronize()
geo = eng._geometry(B, H, W)
outs = eng._alloc_outputs(geo)
stream = torch.cuda.current_stream(dev)
plan = eng._plan(geo, x, outs, stream.cuda_stream)
agg = collections.OrderedDict()
total = []
for rep in range(reps):
    evs = []
    torch.cuda._sleep(40_000_000)
    t0 = torch.cuda.Event(enable_timing=True); t0.record(stream)
    for i, (name, fn, args) in enumerate(plan):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); _lib.check(fn(*args), name); b.record(stream)
        evs.append((i, name, args, a, b))
    t1 = torch.cuda.Event(enable_timing=True); t1.record(stream)
    torch.cuda.synchronize()
    total.append(t0.elapsed_time(t1))
    if rep == 0:
        continue
    for i, name, args, a, b in evs:
        key = name
        if name.startswith("ldit_gemm"):
            key = f"{name} N={args[-4] if name != 'ldit_gemm_bias_scale_residual' else args[-3]} K={args[-3] if name != 'ldit_gemm_bias_scale_residual' else args[-2]}"
        if name == "ldit_resample_taps":
            key = f"{name} x{args[-2]}"
        e = agg.setdefault(key, [0, 0.0]); e[0] += 1; e[1] += a.elapsed_time(b)
n = reps - 1
tot = sum(v for _, v in agg.values()) / n
print(f"{wl}: eager step with events {sum(total[1:]) / n:.3f} ms; sum of per-launch durations {tot:.3f} ms")
for k, (c, v) in agg.items():
    print(f"  {k:55s} x{c // n:3d}  avg {1e3 * v / c:7.1f} us   total {v / n:7.3f} ms  {100 * v / n / tot:5.1f}%")
