"""Time the tcgen05 GEMM entry points on the shapes of the forward (CUDA events, L2-warm and L2-cold)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
M = int(os.environ.get("M", 12608))
shapes = [("qkv", "bias", M, 2304, 768), ("proj", "resid", M, 768, 768), ("fc1", "gelu", M, 3072, 768), ("fc2", "resid", M, 768, 3072)]
if len(sys.argv) > 1 and sys.argv[1] == "large":
    shapes = [("qkv", "bias", M, 3072, 1024), ("proj", "resid", M, 1024, 1024), ("fc1", "gelu", M, 4096, 1024), ("fc2", "resid", M, 1024, 4096)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, kind, m, n, k in shapes:
    A = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    W = (torch.randn(n, k, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    scale = torch.rand(n, device="cuda")
    out_b = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    x = torch.zeros(m, n, device="cuda")
    for ctas, bn in [(c, b) for c in (1, 2) for b in (128, 192, 256)]:
        if n % bn:
            continue
        lib.ldit_set_gemm_cta_pair(ctas)
        lib.ldit_set_gemm_tile_n(bn)
        def call():
            if kind == "bias":
                return lib.ldit_gemm_bias(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), m, n, k, st)
            if kind == "gelu":
                return lib.ldit_gemm_bias_gelu(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), m, n, k, st)
            return lib.ldit_gemm_bias_scale_residual(A.data_ptr(), W.data_ptr(), bias.data_ptr(), scale.data_ptr(), x.data_ptr(), m, n, k, st)
        for _ in range(3):
            _lib.check(call(), name)
        torch.cuda.synchronize()
        res = {}
        for mode in ("warm", "cold"):
            ts = []
            for _ in range(10):
                if mode == "cold":
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); call(); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            res[mode] = ts[len(ts) // 2]
        fl = 2.0 * m * n * k
        print(f"{name:5s} {kind:5s} M={m} N={n} K={k} ctas={ctas} BN={bn}: warm {res['warm']*1e3:7.1f} us {fl/res['warm']/1e9:7.1f} TF/s | cold {res['cold']*1e3:7.1f} us {fl/res['cold']/1e9:7.1f} TF/s")
lib.ldit_set_gemm_tile_n(0)
lib.ldit_set_gemm_cta_pair(2)
