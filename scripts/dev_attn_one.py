"""Dev (GPU): run one fd_attention_d64 shape a few times (for ncu). usage: dev_attn_one.py batch heads lq lk [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from faithdiff_dev_b200 import ops
B, H, LQ, LK = [int(v) for v in sys.argv[1:5]]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
qkv = torch.randn(B * max(LQ, LK), 3 * H * 64, device="cuda").to(torch.bfloat16)
q, k, v = qkv[:B * LQ, :H * 64], qkv[:B * LK, H * 64:2 * H * 64], qkv[:B * LK, 2 * H * 64:]
out = torch.empty(B * LQ, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(reps): ops.attention(q, k, v, B, H, LQ, LK, 0.125, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(reps): ops.attention(q, k, v, B, H, LQ, LK, 0.125, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("B=%d H=%d LQ=%d LK=%d: %.3f ms %.1f TFLOP/s" % (B, H, LQ, LK, ms, 4.0 * B * H * LQ * LK * 64 / ms / 1e9))
