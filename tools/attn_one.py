import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
B, heads, G = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
impl = int(sys.argv[4]) if len(sys.argv) > 4 else 0
lib.ldit_set_attention_impl(impl)
N, D = G * G + 1, heads * 64
st = torch.cuda.current_stream().cuda_stream
qkv = (torch.randn(B * N, 3 * D, device="cuda")).to(torch.bfloat16)
ctx = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); rc = lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), None, B, N, heads, G, G, st); b.record()
    torch.cuda.synchronize(); assert rc == 0
print(f"B={B} heads={heads} G={G} impl={impl}: {a.elapsed_time(b)*1e3:.1f} us")
