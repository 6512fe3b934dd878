"""Fixed cost vs per-round cost of the persistent GEMM: M = 74 row-blocks of 256, N = 192 r -> exactly r schedule
rounds of 256x192 tiles.  Each point is the per-launch time of a CUDA graph of 20 back-to-back launches (no event or
host overhead inside).  usage: gemm_rounds.py [K]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 768
M = 74 * 256
lib.ldit_set_gemm_tile_n(192)
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
for kind in ("bias", "gelu", "resid"):
    pts = []
    for r in (1, 2, 3, 4, 8, 16):
        n = 192 * r
        W = (torch.randn(n, K, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(n, device="cuda"); scale = torch.rand(n, device="cuda")
        out_b = torch.empty(M, n, device="cuda", dtype=torch.bfloat16); x = torch.zeros(M, n, device="cuda")
        def call(st):
            if kind == "bias": rc = lib.ldit_gemm_bias(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), M, n, K, st)
            elif kind == "gelu": rc = lib.ldit_gemm_bias_gelu(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out_b.data_ptr(), M, n, K, st)
            else: rc = lib.ldit_gemm_bias_scale_residual(A.data_ptr(), W.data_ptr(), bias.data_ptr(), scale.data_ptr(), x.data_ptr(), M, n, K, st)
            _lib.check(rc, kind)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3): call(s.cuda_stream)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20): call(torch.cuda.current_stream().cuda_stream)
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / 20 * 1e3)
        pts.append((r, sorted(ts)[2]))
    (r0, t0), (r1, t1) = pts[3], pts[5]
    slope = (t1 - t0) / (r1 - r0)
    print(f"{kind:5s} K={K}: " + "  ".join(f"r={r}: {t:6.1f}us" for r, t in pts) + f"   | per round {slope:5.2f} us, fixed {t0 - slope * r0:5.2f} us")
