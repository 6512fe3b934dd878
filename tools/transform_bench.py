"""Fused input transform (raw pages -> normalized, resized patch operand) next to the plain patch gather:
device time of ldit_patch_embed_pages vs ldit_patch_embed at batch 64.  usage: transform_bench.py [page_side]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
side = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B, H, W, D = 64, 224, 224, 768
MAXW = 0 if os.environ.get('DIRECT') else side
P = 196
pages = [torch.rand(B, 3, side, side, device="cuda") for _ in range(2)]      # 2 x 805 MB at 1024: rotates past L2
ptrs = [torch.tensor([p[i].data_ptr() for i in range(B)], dtype=torch.int64).cuda() for p in pages]
hw = torch.tensor([[side, side]] * B, dtype=torch.int32).cuda()
x224 = [torch.rand(B, 3, H, W, device="cuda") for _ in range(4)]
scratch = torch.empty(B * P, 768, device="cuda", dtype=torch.bfloat16)
w = torch.randn(D, 768, device="cuda").to(torch.bfloat16); posb = torch.zeros(P, D, device="cuda"); clsp = torch.zeros(D, device="cuda")
x = torch.empty(B * (P + 1), D, device="cuda")
def fused(i, st): return lib.ldit_patch_embed_pages(ptrs[i % 2].data_ptr(), hw.data_ptr(), MAXW, 0, .5, .5, .5, .5, .5, .5, w.data_ptr(), posb.data_ptr(), clsp.data_ptr(), scratch.data_ptr(), x.data_ptr(), B, H, W, D, st)
def plain(i, st): return lib.ldit_patch_embed(x224[i % 4].data_ptr(), 0, w.data_ptr(), posb.data_ptr(), clsp.data_ptr(), scratch.data_ptr(), x.data_ptr(), B, H, W, D, st)
def timeit(fn, reps=8):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(4): _lib.check(fn(i, s.cuda_stream), "warm")
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps): fn(i, torch.cuda.current_stream().cuda_stream)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / reps * 1e3)
    return sorted(ts)[2]
tp, tf = timeit(plain), timeit(fused)
rows_used = min(side, 2 * 224)          # each output row blends 2 source rows; sectors of the rows it touches are read whole
touched = B * 3 * rows_used * side * 4 if side > 224 else B * 3 * side * side * 4
print(f"patch embed from a resident 224x224 fp32 batch: {tp:.1f} us (3 launches)")
print(f"patch embed from raw {side}x{side} fp32 pages (normalize + bilinear resize fused): {tf:.1f} us "
      f"(+{tf - tp:.1f} us); source rows touched {touched / 1e6:.0f} MB -> {touched / ((tf - tp + 11) * 1e-6) / 1e9:.0f} GB/s "
      f"if the extra time plus the plain gather's ~11 us is charged to them; full pages are {B * 3 * side * side * 4 / 1e6:.0f} MB")
