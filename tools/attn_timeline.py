import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
B, heads, G = (64, 12, 14) if len(sys.argv) < 2 else (32, 12, 32)
N, D = G * G + 1, heads * 64
st = torch.cuda.current_stream().cuda_stream
qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
ctx = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3): lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), None, B, N, heads, G, G, st)
buf = torch.zeros(148 * 2 * 16 * 8, dtype=torch.int64, device="cuda")
lib.ldit_debug_attention_timeline(buf.data_ptr())
lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), None, B, N, heads, G, G, st)
torch.cuda.synchronize()
lib.ldit_debug_attention_timeline(None)
t = buf.cpu().reshape(148, 2, 16, 8)
for cta in (0, 77):
    base = int(t[cta, 0, 0, 0])
    for g in (0, 1):
        for i in range(5):
            r = [int(v) - base for v in t[cta, g, i, :5]]
            print(f"cta{cta} wg{g} item{i}: start {r[0]:6d} | S0 ready {r[1]:6d} | last P {r[2]:6d} | O ready {r[3]:6d} | stored {r[4]:6d}   (S wait {r[1]-r[0]}, softmax {r[2]-r[1]}, O wait {r[3]-r[2]}, store {r[4]-r[3]})")
