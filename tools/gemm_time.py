"""Device time of one GEMM entry point, enqueued behind a sleep kernel (no host overhead in the events):
gemm_time.py kind M N K [bn]"""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
lib.ldit_set_gemm_tile_n(bn)
def call():
    if kind == "bias": rc = lib.ldit_gemm_bias(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), m, n, k, st)
    elif kind == "gelu": rc = lib.ldit_gemm_bias_gelu(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), m, n, k, st)
    else: rc = lib.ldit_gemm_bias_scale_residual(A.data_ptr(), W.data_ptr(), bias.data_ptr(), scale.data_ptr(), x.data_ptr(), m, n, k, st)
    _lib.check(rc, kind)
for _ in range(3): call()
torch.cuda.synchronize()
res = {}
for mode in ("warm", "cold"):
    torch.cuda._sleep(20_000_000)
    evs = []
    for _ in range(10):
        if mode == "cold": flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    res[mode] = ts[len(ts) // 2]
fl = 2.0 * m * n * k
print(f"{kind} M={m} N={n} K={k} BN={bn}: warm {res['warm']*1e3:6.1f} us {fl/res['warm']/1e9:7.1f} TF/s | cold {res['cold']*1e3:6.1f} us {fl/res['cold']/1e9:7.1f} TF/s")
