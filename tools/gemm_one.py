"""Run one GEMM entry point a few times (for ncu captures): gemm_one.py kind M N K [bn]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
kind, m, n, k = sys.argv[1], *map(int, sys.argv[2:5])
bn = int(sys.argv[5]) if len(sys.argv) > 5 else 0
st = torch.cuda.current_stream().cuda_stream
A = torch.randn(m, k, device="cuda").to(torch.bfloat16); W = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.randn(n, device="cuda"); scale = torch.rand(n, device="cuda")
out_b = torch.empty(m, n, device="cuda", dtype=torch.bfloat16); x = torch.zeros(m, n, device="cuda")
lib.ldit_set_gemm_tile_n(bn)
for _ in range(4):
    if kind == "bias": rc = lib.ldit_gemm_bias(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), m, n, k, st)
    elif kind == "gelu": rc = lib.ldit_gemm_bias_gelu(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), m, n, k, st)
    else: rc = lib.ldit_gemm_bias_scale_residual(A.data_ptr(), W.data_ptr(), bias.data_ptr(), scale.data_ptr(), x.data_ptr(), m, n, k, st)
    _lib.check(rc, kind)
torch.cuda.synchronize()
print("ok")
