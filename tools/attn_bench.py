"""ldit_attention alone: back-to-back launches at the bench shapes, with and without a relative-position table."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for B, heads, G in ((64, 12, 14), (32, 12, 32), (64, 16, 14)):
    N, D = G * G + 1, heads * 64
    qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
    ctx = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    T = (2 * G - 1) ** 2 + 3
    table = torch.randn(heads, T, device="cuda")
    for name, tab in (("no bias", None), ("rel-pos bias", table.data_ptr())):
        for _ in range(3): _lib.check(lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), tab, B, N, heads, G, G, st), "attn")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), tab, B, N, heads, G, G, st)
        b.record(); torch.cuda.synchronize()
        us = 1e3 * a.elapsed_time(b) / 20
        print(f"B={B} heads={heads} N={N} {name:13s}: {us:7.1f} us  {4.0 * B * heads * N * N * 64 / us / 1e6:7.1f} TFLOP/s")
