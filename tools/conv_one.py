"""Run the FPN's 3x3 convolution (implicit tcgen05 GEMM) a few times at the p2 size of base224 (for ncu captures):
conv_one.py [B H W]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layoutdit_b200 import _lib
lib = _lib.load()
B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 56, 56)
C = 256
x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
w = (torch.randn(C, 9 * C, device="cuda") * 0.03).to(torch.bfloat16)
bias = torch.randn(C, device="cuda")
out = torch.empty(B, H, W, C, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    _lib.check(lib.ldit_conv3x3_bias(x.data_ptr(), w.data_ptr(), bias.data_ptr(), out.data_ptr(), B, H, W, C, C,
                                     torch.cuda.current_stream().cuda_stream), "conv")
torch.cuda.synchronize()
print("ok")
