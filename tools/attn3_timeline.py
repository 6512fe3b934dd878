"""clock64 timeline of attention_v3 (needs a -DLDIT_A3_TIMELINE build: LDIT_LIB_PATH=build_variants/lib_tl.so).
Softmax warps stamp per step: [before s_full wait, after, after pass 1, after pass 2, after pv catch-up];
the issuer per step: [before v/k wait, after, p_full(g0) seen, p_full(g1) seen]."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
B, heads, G = (64, 12, 14) if len(sys.argv) < 2 or sys.argv[1] == "224" else (32, 12, 32)
N, D = G * G + 1, heads * 64
st = torch.cuda.current_stream().cuda_stream
lib.ldit_set_attention_impl(0)
qkv = torch.randn(B * N, 3 * D, device="cuda").to(torch.bfloat16)
ctx = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3): lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), None, B, N, heads, G, G, st)
buf = torch.zeros(8 * 10 * 512, dtype=torch.int64, device="cuda")
lib.ldit_debug_attention_timeline(buf.data_ptr())
lib.ldit_attention(qkv.data_ptr(), ctx.data_ptr(), None, B, N, heads, G, G, st)
torch.cuda.synchronize()
lib.ldit_debug_attention_timeline(None)
t = buf.cpu().reshape(8, 10, 512)
T = (N + 31) // 32
for cta in (0, 1):
    base = int(t[cta, 0, 0])
    for w in (0, 4):
        v = [int(x) - base for x in t[cta, w] if int(x) != 0]
        print(f"cta{cta} warp{w}: {len(v)} stamps; steps (s-wait, pass1, pass2, pv-wait | step total):")
        for s in range(0, min(len(v) // 5, 2 * T + 2)):
            a = v[5 * s: 5 * s + 5]
            nxt = v[5 * s + 5] if 5 * s + 5 < len(v) else a[4]
            print(f"   step {s:3d} @{a[0]:7d}: {a[1]-a[0]:5d} {a[2]-a[1]:5d} {a[3]-a[2]:5d} {a[4]-a[3]:5d} | {nxt-a[0]:5d}")
    v = [int(x) - base for x in t[cta, 9] if int(x) != 0]
    print(f"cta{cta} issuer: {len(v)} stamps; steps (kv wait, p_full g0 wait, p_full g1 wait | step total)")
    for s in range(0, min(len(v) // 4, 2 * T + 2)):
        a = v[4 * s: 4 * s + 4]
        nxt = v[4 * s + 4] if 4 * s + 4 < len(v) else a[3]
        print(f"   step {s:3d} @{a[0]:7d}: {a[1]-a[0]:5d} {a[2]-a[1]:5d} {a[3]-a[2]:5d} | {nxt-a[0]:5d}")
