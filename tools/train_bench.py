"""Forward + backward of encoder layers at the bench shape: layoutdit_b200.train (hand-written backward over the C ABI)
against HF BeitLayer under torch autocast bf16 (cuBLASLt / SDPA / ATen autograd).  usage: train_bench.py [layers]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import config as cfgmod
from layoutdit_b200.dit_params import DiTParameters
from layoutdit_b200.train import TrainableEncoder
from layoutdit_b200.synth import make_state_dict

L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, G = 64, 14
cfg = cfgmod.dit_base()
cfg = cfg.__class__(**{**cfg.to_dict(), "num_hidden_layers": L})
dev = torch.device("cuda", 0)
params = DiTParameters(cfg).to(dev)
params.load_state_dict(make_state_dict(cfg, 0, False), strict=False)
enc = TrainableEncoder(params, cfg)
N, D = G * G + 1, cfg.hidden_size
x = torch.randn(B, N, D, device=dev, requires_grad=True)
g = torch.randn(B, N, D, device=dev)

def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def ours():
    for p in params.parameters(): p.grad = None
    x.grad = None
    y = enc(x, G, G)
    y.backward(g)
def ours_fwd():
    with torch.no_grad():
        enc(x, G, G)
t_all, t_f = timeit(ours), timeit(ours_fwd)
print(f"ours : {L} layers fwd+bwd {t_all:.3f} ms ({t_all / L:.3f} per layer), fwd only {t_f:.3f} ms")

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ours(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))

from transformers.models.beit.modeling_beit import BeitLayer
from transformers import BeitConfig
hc = BeitConfig(**{k: v for k, v in cfg.to_dict().items() if k in BeitConfig().to_dict()})
hc._attn_implementation = "sdpa"
layers = torch.nn.ModuleList([BeitLayer(hc, drop_path_rate=0.0) for _ in range(L)]).to(dev).train()
xh = torch.randn(B, N, D, device=dev, requires_grad=True)
def hf():
    for p in layers.parameters(): p.grad = None
    xh.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        h = xh
        for l in layers:
            o = l(h)
            h = o[0] if isinstance(o, tuple) else o
    h.backward(g)
t_h = timeit(hf)
print(f"HF   : {L} layers fwd+bwd {t_h:.3f} ms ({t_h / L:.3f} per layer)  [autocast bf16, attn={hc._attn_implementation}]")
