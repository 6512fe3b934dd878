"""Device time of the HBM-bound kernels (LayerNorm, taps, patch embed) at base224 sizes: each kernel is
captured 16x into a CUDA graph rotating over 4 input buffers (155 MB > L2), time = replay / 16."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
B, G, D = 64, 14, 768
N = G * G + 1; M = B * N
xs = [torch.randn(M, D, device="cuda") for _ in range(4)]
g = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
ys = [torch.empty(M, D, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
def timeit(name, fn, nbytes, reps=16):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(4): fn(i, s.cuda_stream)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        st = torch.cuda.current_stream().cuda_stream
        for i in range(reps): fn(i % 4, st)
    ts = []
    for _ in range(7):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e) / reps)
    t = sorted(ts)[3]
    print(f"{name:30s} {t*1e3:7.1f} us  {nbytes/t/1e6:7.0f} GB/s (algorithmic bytes {nbytes/1e6:.1f} MB)")
timeit("layernorm", lambda i, st: lib.ldit_layernorm(xs[i].data_ptr(), g.data_ptr(), b.data_ptr(), ys[i].data_ptr(), M, D, 1e-12, st), M * D * 6)
for s in (4.0, 2.0, 1.0, 0.5):
    oh = int(G * s)
    outs = [torch.empty(B, oh, oh, D, device="cuda", dtype=torch.bfloat16) for _ in range(2 if s == 4.0 else 4)]
    timeit(f"resample_taps x{s}", lambda i, st: lib.ldit_resample_taps(xs[i].data_ptr(), outs[i % len(outs)].data_ptr(), B, G, G, D, s, st), B * G * G * D * 4 + outs[0].numel() * 2)
pix = [torch.randn(B, 3, 224, 224, device="cuda") for _ in range(4)]
w = torch.randn(D, 768, device="cuda").to(torch.bfloat16); pb = torch.randn(G * G, D, device="cuda"); cp = torch.randn(D, device="cuda")
scratch = torch.empty(B * G * G * 768, device="cuda", dtype=torch.bfloat16)
timeit("patch_embed (all kernels)", lambda i, st: lib.ldit_patch_embed(pix[i].data_ptr(), 0, w.data_ptr(), pb.data_ptr(), cp.data_ptr(), scratch.data_ptr(), xs[i].data_ptr(), B, 224, 224, D, st), pix[0].numel() * 4 + M * D * 4 + scratch.numel() * 4)
