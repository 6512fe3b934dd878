"""Diagnostic for the tcgen05 GEMM: prints where (rows / columns / K) the result departs
from an fp32 reference.  Run on a GPU box: python tools/gemm_diag.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib  # noqa: E402

lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream


def run(M, N, K, bn, mode="rand"):
    lib.ldit_set_gemm_tile_n(bn)
    g = torch.Generator(device="cuda").manual_seed(1)
    if mode == "rand":
        A = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    else:  # structured: A[m,k] = (m%7)+1 if k==m%K ; W[n,k] = k+ n*0.001 -> out[m,n] identifies (m, k) routing
        A = torch.zeros(M, K, device="cuda")
        A[torch.arange(M), torch.arange(M) % K] = 1.0
        A = A.to(torch.bfloat16)
        W = (torch.arange(K, device="cuda")[None, :].float() + 256 * (torch.arange(N, device="cuda")[:, None] % 4)).to(torch.bfloat16)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib.ldit_gemm_bias(A.data_ptr(), W.data_ptr(), None, out.data_ptr(), M, N, K, st)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    err = (out.float() - ref).abs()
    nan = torch.isnan(out.float())
    print(f"M={M} N={N} K={K} bn={bn} mode={mode} rc={rc} max_err={float(err[~nan].max()) if (~nan).any() else -1:.4f} "
          f"ref_absmax={float(ref.abs().max()):.2f} nan_frac={float(nan.float().mean()):.4f}")
    if float(err[~nan].max() if (~nan).any() else 1) > 0.05 * float(ref.abs().max()) or nan.any():
        bad = (err > 0.02 * ref.abs().max()) | nan
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        print("  bad rows:", rows[:16].tolist(), "... n=", len(rows), " bad cols:", cols[:16].tolist(), "... n=", len(cols))
        print("  out[0,:8]", out[0, :8].float().tolist())
        print("  ref[0,:8]", ref[0, :8].tolist())
        if mode != "rand":
            print("  out[:16,0]", out[:16, 0].float().tolist())
            print("  ref[:16,0]", ref[:16, 0].tolist())


if __name__ == "__main__":
    lib.ldit_set_gemm_cta_pair(int(os.environ.get("CTAS", "2")))
    for args in [(128, 128, 64, 128, "struct"), (128, 128, 64, 128, "rand"), (128, 256, 64, 256, "rand"),
                 (128, 192, 64, 192, "rand"), (128, 128, 256, 128, "rand"), (256, 256, 768, 128, "rand"),
                 (1000, 768, 768, 192, "rand"), (12608, 2304, 768, 0, "rand")]:
        try:
            run(*args)
        except Exception as e:  # noqa: BLE001
            print("EXC", args, e)
            break
