import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
kind, m, n, k, bn = sys.argv[1], *map(int, sys.argv[2:6])
st = torch.cuda.current_stream().cuda_stream
A = torch.randn(m, k, device="cuda").to(torch.bfloat16); W = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
if os.environ.get("ZERO"): A.zero_(); W.zero_()
bias = torch.randn(n, device="cuda"); scale = torch.rand(n, device="cuda")
out_b = torch.empty(m, n, device="cuda", dtype=torch.bfloat16); x = torch.zeros(m, n, device="cuda")
lib.ldit_set_gemm_cta_pair(2); lib.ldit_set_gemm_tile_n(bn)
def call():
    if kind == "bias": return lib.ldit_gemm_bias(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), m, n, k, st)
    if kind == "gelu": return lib.ldit_gemm_bias_gelu(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), m, n, k, st)
    return lib.ldit_gemm_bias_scale_residual(A.data_ptr(), W.data_ptr(), bias.data_ptr(), scale.data_ptr(), x.data_ptr(), m, n, k, st)
for _ in range(3): call()
buf = torch.zeros(74 * 16 * 16, dtype=torch.int64, device="cuda")
lib.ldit_debug_gemm_timeline(buf.data_ptr()); call(); torch.cuda.synchronize(); lib.ldit_debug_gemm_timeline(None)
t = buf.cpu().reshape(74, 16, 16)
for cl in (0, 40):
    base = int(t[cl, 0, 0])
    for i in range(8):
        r = [int(v) - base for v in t[cl, i, :7]]; wf = int(t[cl, i, 7])
        if t[cl, i, 0] == 0: break
        print(f"cl{cl} tile{i}: MMA: start {r[0]:6d} tempty-ok {r[1]:6d} kb0-issued {r[2]:6d} all-issued {r[3]:6d} | EPI: wait-start {r[4]:6d} tfull {r[5]:6d} done {r[6]:6d}  (mma wait tempty {r[1]-r[0]}, first full wait {r[2]-r[1]}, issue span {r[3]-r[1]}, epi wait {r[5]-r[4]}, epi work {r[6]-r[5]}, full-wait total {wf}; producer empty-wait total {int(t[cl,i,9])})")
