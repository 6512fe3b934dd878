"""fc1 + fc2 as two launches vs ldit_mlp_fused: per-layer device time (CUDA graph of 12 back-to-back MLPs, each preceded by a
LayerNorm so the surrounding launch pattern is the forward's).  usage: mlp_bench.py [M D I]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
M, D, I = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (12608, 768, 3072)
a = torch.randn(M, D, device="cuda").to(torch.bfloat16)
W1 = (torch.randn(I, D, device="cuda") * 0.05).to(torch.bfloat16); W2 = (torch.randn(D, I, device="cuda") * 0.05).to(torch.bfloat16)
b1, b2, lam = torch.randn(I, device="cuda"), torch.randn(D, device="cuda"), torch.rand(D, device="cuda")
gam, bet = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
x = torch.randn(M, D, device="cuda"); h = torch.empty(M, I, device="cuda", dtype=torch.bfloat16)
stride = lib.ldit_mlp_schedule(M, D, I, None, 0)
host = torch.empty(lib.ldit_mlp_clusters() * stride, dtype=torch.int32); lib.ldit_mlp_schedule(M, D, I, host.data_ptr(), host.numel())
sched = host.cuda(); ready = torch.zeros(2 * ((M + 255) // 256), device="cuda", dtype=torch.int32)
def separate(st):
    lib.ldit_layernorm(x.data_ptr(), gam.data_ptr(), bet.data_ptr(), a.data_ptr(), M, D, 1e-12, st)
    lib.ldit_gemm_bias_gelu(a.data_ptr(), W1.data_ptr(), b1.data_ptr(), h.data_ptr(), M, I, D, st)
    lib.ldit_gemm_bias_scale_residual(h.data_ptr(), W2.data_ptr(), b2.data_ptr(), lam.data_ptr(), x.data_ptr(), M, D, I, st)
def fused(st):
    lib.ldit_layernorm(x.data_ptr(), gam.data_ptr(), bet.data_ptr(), a.data_ptr(), M, D, 1e-12, st)
    lib.ldit_mlp_fused(a.data_ptr(), W1.data_ptr(), b1.data_ptr(), h.data_ptr(), W2.data_ptr(), b2.data_ptr(), lam.data_ptr(), x.data_ptr(),
                       M, D, I, sched.data_ptr(), stride, ready.data_ptr(), st)
def lnonly(st):
    lib.ldit_layernorm(x.data_ptr(), gam.data_ptr(), bet.data_ptr(), a.data_ptr(), M, D, 1e-12, st)
def timeit(fn, reps=12):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn(s.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn(torch.cuda.current_stream().cuda_stream)
    ts = []
    for _ in range(7):
        x.normal_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / reps * 1e3)
    return sorted(ts)[3]
tl, ts_, tf = timeit(lnonly), timeit(separate), timeit(fused)
print(f"M={M} D={D} I={I}: LayerNorm {tl:.1f} us | LN + fc1 + fc2 (two GEMM launches) {ts_:.1f} us | LN + fused MLP {tf:.1f} us  -> MLP {ts_ - tl:.1f} vs {tf - tl:.1f} us")
