"""Dev (GPU): fd_attention_d64 correctness (vs torch SDPA in fp32) and TFLOP/s at the tile-step's two self-attention shapes,
for the variant selected by FD_ATTN_POLY (0 = all exponentials on MUFU; 2/3/4 of every 8 pairs on the FMA pipe)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from faithdiff_dev_b200 import ops
tag = "FD_ATTN_POLY=%s" % os.environ.get("FD_ATTN_POLY", "0")
torch.manual_seed(0)
for B, H, L, spread in [(2, 3, 640, 1.0), (1, 2, 1000, 6.0), (1, 1, 130, 1.0), (2, 2, 200, 30.0), (14, 10, 4096, 1.0), (14, 20, 1024, 1.0)]:
    qkv = (torch.randn(B * L, 3 * H * 64, device="cuda") * spread).to(torch.bfloat16)
    q, k, v = qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:]
    out = ops.attention(q, k, v, B, H, L, L, 0.125)
    if B * H * L * L <= 2 ** 27:
        r = lambda t: t.float().view(B, L, H, 64).transpose(1, 2)
        ref = F.scaled_dot_product_attention(r(q), r(k), r(v)).transpose(1, 2).reshape(B * L, H * 64)
        print("%s B=%d H=%d L=%d spread %.0f: rel-L2 vs SDPA fp32 %.3e" % (tag, B, H, L, spread, ((out.float() - ref).norm() / ref.norm()).item()), flush=True)
    else:
        for _ in range(3):
            ops.attention(q, k, v, B, H, L, L, 0.125, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            ops.attention(q, k, v, B, H, L, L, 0.125, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("%s B=%d H=%d L=%d: %.3f ms %.1f TFLOP/s" % (tag, B, H, L, ms, 4.0 * B * H * L * L * 64 / ms / 1e9), flush=True)
