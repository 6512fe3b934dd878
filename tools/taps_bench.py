import os, sys, math
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load(); st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, G, D in [(64, 14, 768), (32, 32, 768)]:
    x = torch.randn(B, G * G + 1, D, device="cuda")
    for scale in (4.0, 2.0, 1.0, 0.5):
        oh = int(math.floor(G * scale))
        out = torch.empty(B, oh, oh, D, device="cuda", dtype=torch.bfloat16)
        ts = []
        for _ in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): lib.ldit_resample_taps(x.data_ptr(), out.data_ptr(), B, G, G, D, scale, st)
            b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / 5)
        t = sorted(ts)[3]
        byts = out.numel() * 2 + x.numel() * 4
        print(f"B={B} G={G} scale={scale}: {t*1e3:7.1f} us  {byts/t/1e6:7.1f} GB/s (algorithmic: read tokens once + write taps)")
