"""Run one GEMM configuration a few times (for ncu).  usage: gemm_one.py kind M N K ctas bn [iters]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
kind, m, n, k, ctas, bn = sys.argv[1], *map(int, sys.argv[2:7])
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 5
st = torch.cuda.current_stream().cuda_stream
A = torch.randn(m, k, device="cuda").to(torch.bfloat16)
W = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.randn(n, device="cuda"); scale = torch.rand(n, device="cuda")
out_b = torch.empty(m, n, device="cuda", dtype=torch.bfloat16); x = torch.zeros(m, n, device="cuda")
lib.ldit_set_gemm_cta_pair(ctas); lib.ldit_set_gemm_tile_n(bn)
def call():
    if kind == "bias": return lib.ldit_gemm_bias(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), m, n, k, st)
    if kind == "gelu": return lib.ldit_gemm_bias_gelu(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), m, n, k, st)
    return lib.ldit_gemm_bias_scale_residual(A.data_ptr(), W.data_ptr(), bias.data_ptr(), scale.data_ptr(), x.data_ptr(), m, n, k, st)
for _ in range(3):
    assert call() == 0
torch.cuda.synchronize()
reps = 20
ts = []
for i in range(iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        call()
    b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / reps)
t = sorted(ts)[len(ts) // 2]
print(f"{kind} M={m} N={n} K={k} ctas={ctas} bn={bn} dbg={os.environ.get('LDIT_GEMM_DBG','0')}: {t*1e3:.1f} us  {2.0*m*n*k/t/1e9:.1f} TF/s  (avg of {reps} back-to-back launches)")
