"""DiTWithFPN forward (graph replay) next to the backbone-only forward, and the per-launch times of the FPN kernels
in situ.  usage: fpn_bench.py [workload]"""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS
from layoutdit_b200 import DiTBackbone, DiTWithFPN, _lib, config as cfgmod
from layoutdit_b200.synth import make_fpn_state_dict, make_state_dict, synthetic_pages

wl = sys.argv[1] if len(sys.argv) > 1 else "base224"
fac, B, H, W = WORKLOADS[wl]
cfg = getattr(cfgmod, fac)()
dev = torch.device("cuda", 0)
sd, fsd = make_state_dict(cfg, 0, False), make_fpn_state_dict(cfg.hidden_size, 256, 0, False)
x = synthetic_pages(B, H, W, 1234).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def time_graphed(model, n=20):
    for _ in range(5): model(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); model(x); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

bb = DiTBackbone(pretrained=False, config=cfg, state_dict=sd, use_cuda_graph=True).to(dev).eval()
fp = DiTWithFPN(pretrained=False, config=cfg, state_dict=sd, fpn_state_dict=fsd, use_cuda_graph=True).to(dev).eval()
t_bb, t_fp = time_graphed(bb), time_graphed(fp)
print(f"{wl}: backbone (4 D-channel taps) {t_bb:.3f} ms = {B / t_bb * 1e3:.0f} img/s | backbone + FPN (p2..p5, pool) {t_fp:.3f} ms = {B / t_fp * 1e3:.0f} img/s")

eng = fp.backbone._get_engine()
geo = eng._geometry(B, H, W, 0, "fpn")
outs = eng._alloc_outputs(geo)
stream = torch.cuda.current_stream(dev)
plan = eng._plan(geo, x, outs, stream.cuda_stream)
first = next(i for i, (n, _, a) in enumerate(plan) if n == "ldit_fpn_merge")
agg = collections.OrderedDict()
reps = 4
for rep in range(reps + 1):
    torch.cuda._sleep(40_000_000)
    evs = []
    for i, (name, fn, args) in enumerate(plan):
        fpn_launch = i >= first or (name == "ldit_gemm_bias" and args[5] == 256) or (name == "ldit_resample_taps" and plan[i + 1][0] == "ldit_gemm_bias" and plan[i + 1][2][5] == 256)
        if fpn_launch:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); _lib.check(fn(*args), name); b.record(stream)
            key = name
            if name == "ldit_conv3x3_bias": key += f" {args[5]}x{args[6]}"
            if name == "ldit_fpn_merge": key += f" x{args[7]}"
            evs.append((key, a, b))
        else:
            _lib.check(fn(*args), name)
    torch.cuda.synchronize()
    if rep:
        for key, a, b in evs:
            e = agg.setdefault(key, [0, 0.0]); e[0] += 1; e[1] += a.elapsed_time(b)
tot = 0.0
for k, (c, v) in agg.items():
    print(f"  {k:40s} x{c // reps:2d}  avg {1e3 * v / c:7.1f} us")
    tot += v / reps
print(f"  FPN launches total {tot:.3f} ms (events add ~2-3 us per launch)")
for slot, s in enumerate((4.0, 2.0, 1.0, 0.5)):
    h, w = int(geo.Gh * s), int(geo.Gw * s)
    fl = 2.0 * B * h * w * 9 * 256 * 256
    k = f"ldit_conv3x3_bias {h}x{w}"
    if k in agg:
        t = agg[k][1] / agg[k][0]
        print(f"  conv {h}x{w}: {fl / 1e9:.1f} GFLOP algorithmic -> {fl / t / 1e9:.0f} TFLOP/s")
