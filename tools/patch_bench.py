"""ldit_patch_embed: fp32 pixels (cast/gather pass + CLS kernel + GEMM) vs fp16 pixels (one TMA-fed kernel)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
for (B, H, W, D) in [(64, 224, 224, 768), (32, 512, 512, 768), (64, 224, 224, 1024)]:
    G = (H // 16) * (W // 16)
    w = (torch.randn(D, 768, device="cuda") * 0.04).to(torch.bfloat16)
    pb, cp = torch.randn(G, D, device="cuda"), torch.randn(D, device="cuda")
    scratch = torch.empty(B * G * 768, device="cuda", dtype=torch.bfloat16)
    x = torch.empty(B, G + 1, D, device="cuda")
    for dt, code in ((torch.float32, 0), (torch.bfloat16, 2)):
        px = [(torch.rand(B, 3, H, W, device="cuda") * 2 - 1).to(dt) for _ in range(4)]
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for i in range(4):
                _lib.check(lib.ldit_patch_embed(px[i].data_ptr(), code, w.data_ptr(), pb.data_ptr(), cp.data_ptr(), scratch.data_ptr(), x.data_ptr(), B, H, W, D, s.cuda_stream), "pe")
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            st = torch.cuda.current_stream().cuda_stream
            for i in range(16):
                lib.ldit_patch_embed(px[i % 4].data_ptr(), code, w.data_ptr(), pb.data_ptr(), cp.data_ptr(), scratch.data_ptr(), x.data_ptr(), B, H, W, D, st)
        ts = []
        for _ in range(7):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gr.replay(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e) / 16)
        t = sorted(ts)[3]
        fl = 2.0 * B * G * 768 * D
        print(f"B={B} {H}x{W} D={D} pixels {str(dt)[6:]:8s}: {t*1e3:7.1f} us  {fl/t/1e9:7.1f} TFLOP/s")
