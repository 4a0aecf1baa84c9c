"""Dev (GPU): phase timeline of the long-KV attention kernel (library built with -DFD_ATTN_TRACE): clock64 stamps of one
CTA's first softmax warp and of its MMA thread over key blocks 16..31."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faithdiff_dev_b200 import ops
from faithdiff_dev_b200._lib import lib
B, H, L = 14, 10, 4096
qkv = torch.randn(B * L, 3 * H * 64, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:]
for _ in range(3):
    ops.attention(q, k, v, B, H, L, L, 0.125)
torch.cuda.synchronize()
buf = (C.c_longlong * (16 * 12))()
print("fd_debug_read_attn_trace ->", lib.fd_debug_read_attn_trace(buf))
t = [[buf[j * 12 + kk] for kk in range(12)] for j in range(16)]
t0 = t[0][0]
for j in range(1, 15):
    r = t[j]
    print("block %2d softmax: top %6d | wait S %5d | load S %5d | max %5d | exps %5d | wait PV %5d | store+hand-off %5d || MMA: S(j+1) issued at %6d (%4d) | P ready at %6d | PV issued by %6d"
          % (j + 16, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], r[5] - r[4], r[6] - r[5], r[8] - t0, r[9] - r[8], r[10] - t0, r[11] - t0))
