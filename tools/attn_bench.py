"""Time ldit_attention (tcgen05 vs mma.sync variants) on the BASELINE geometries."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, B, heads, G in [("base224 b64", 64, 12, 14), ("base512 b32", 32, 12, 32), ("large224 b64", 64, 16, 14)]:
    N, D = G * G + 1, heads * 64
    qkv = (torch.randn(B * N, 3 * D, device="cuda")).to(torch.bfloat16)
    ctx = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    for impl in ((0, 4) if lib.ldit_has_experimental() else (0,)):
        lib.ldit_set_attention_impl(impl)
        for _ in range(3):
            _lib.check(lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), None, B, N, heads, G, G, st), "attn")
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), None, B, N, heads, G, G, st); b.record()
            torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        t = sorted(ts)[len(ts) // 2]
        fl = 4.0 * B * heads * N * N * 64
        byts = B * N * 4 * D * 2
        print(f"{name} impl={impl}: {t*1e3:8.1f} us  {fl/t/1e9:7.1f} TF/s (algorithmic)  {byts/t/1e6:7.1f} GB/s")
lib.ldit_set_attention_impl(0)
