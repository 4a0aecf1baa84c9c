"""Dev (GPU): fd_attention_d512 correctness vs torch SDPA fp32 and TFLOP/s at the VAE / LQ-encoder shapes, next to the round-1
GEMM -> softmax -> GEMM path (FD_VAE_ATTN_GEMM=1) measured through blocks.VaeAttention.  Run under `timeout`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from faithdiff_dev_b200 import ops
from faithdiff_dev_b200.blocks import VaeAttention, Weights
BF = torch.bfloat16
torch.manual_seed(0)
for B, L in [(1, 256), (2, 1000), (1, 4096)]:
    qkv = torch.randn(B * L, 1536, device="cuda").to(BF)
    q, k, v = qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:]
    out = ops.attention_d512(q, k, v, B, L, 512 ** -0.5)
    torch.cuda.synchronize()
    r = lambda t: t.float().reshape(B, 1, L, 512)
    ref = F.scaled_dot_product_attention(r(q), r(k), r(v)).reshape(B * L, 512)
    print("B=%d L=%d rel-L2 vs SDPA fp32 %.3e" % (B, L, ((out.float() - ref).norm() / ref.norm()).item()), flush=True)
for B, L in [(8, 4096), (1, 16384), (4, 16384)]:
    qkv = torch.randn(B * L, 1536, device="cuda").to(BF)
    q, k, v = qkv[:, :512], qkv[:, 512:1024], qkv[:, 1024:]
    out = ops.attention_d512(q, k, v, B, L, 512 ** -0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        ops.attention_d512(q, k, v, B, L, 512 ** -0.5, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("fd_attention_d512 B=%d L=%d: %.3f ms, %.1f TFLOP/s (useful 4*L^2*512 per sample; executed x1.5)" % (B, L, ms, 4.0 * B * L * L * 512 / ms / 1e9), flush=True)
g = torch.Generator().manual_seed(1)
sd = {"group_norm.weight": torch.ones(512), "group_norm.bias": torch.zeros(512)}
for n in ("to_q", "to_k", "to_v", "to_out.0"):
    sd[n + ".weight"] = torch.randn(512, 512, generator=g) / 512 ** 0.5
    sd[n + ".bias"] = torch.zeros(512)
att = VaeAttention(Weights(sd, torch.device("cuda")), 32, 1e-6)
for B, hw in [(8, 64), (1, 128)]:
    x = torch.randn(B, hw, hw, 512, device="cuda").to(BF)
    res = {}
    for mode in ("0", "1"):
        os.environ["FD_VAE_ATTN_GEMM"] = mode
        y = att(x); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(3):
            y = att(x)
        e1.record(); torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 3, y)
    print("VaeAttention block B=%d %dx%d: flash d512 %.3f ms | GEMM-softmax-GEMM %.3f ms | rel-L2 between them %.3e" % (
        B, hw, hw, res["0"][0], res["1"][0], ((res["0"][1].float() - res["1"][1].float()).norm() / res["1"][1].float().norm()).item()), flush=True)
