"""Dev (GPU): full-size SDXL-arch UNet tile-step timing and per-op breakdown."""
import sys, os, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import faithdiff_dev_b200 as fd
from faithdiff_dev_b200 import ops, arch
BF = torch.bfloat16
t0 = time.time()
cfg = arch.SDXL_UNET_CONFIG
sd = arch.random_state_dict(arch.unet_param_shapes(cfg))
unet = fd.UNet2DConditionModel(**cfg); unet.load_state_dict(sd); del sd
torch.cuda.synchronize(); print("weights ready %.1fs, mem %.1f GB" % (time.time() - t0, torch.cuda.memory_allocated() / 1e9), flush=True)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(2, 4, 128, 128, generator=g, device="cuda").to(BF)
lq = torch.randn(2, 512, 128, 128, generator=g, device="cuda").to(BF)
ehs = torch.randn(2, 77, 2048, generator=g, device="cuda").to(BF)
te = torch.randn(2, 1280, generator=g, device="cuda").to(BF)
ids = torch.tensor([[2160., 3840., 64., 128., 1024., 1024.]] * 2, device="cuda")
kw = dict(encoder_hidden_states=ehs, input_embedding=lq, added_cond_kwargs={"text_embeds": te, "time_ids": ids}, return_dict=False)
for _ in range(2): out = unet(x, torch.tensor(951.), **kw)[0]
torch.cuda.synchronize()
print("out finite:", bool(torch.isfinite(out.float()).all()), "std %.3f" % out.float().std().item())
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
n0 = ops.launch_count()
e0.record()
for _ in range(5): out = unet(x, torch.tensor(951.), **kw)[0]
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("UNet forward (B=2, 128x128, incl. stem+cond) %.2f ms  -> %.1f TFLOP/s (14.216 TF)  launches/forward %d" % (ms, 14.216 / ms * 1e3, (ops.launch_count() - n0) // 5), flush=True)
th = time.time()
for _ in range(3): out = unet(x, torch.tensor(951.), **kw)[0]
print("host time per forward (async) %.2f ms" % ((time.time() - th) / 3 * 1e3)); torch.cuda.synchronize()
# per-op breakdown
rec = collections.defaultdict(list)
def wrap(name):
    f = getattr(ops, name)
    def g(*a, **k):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); r = f(*a, **k); e.record()
        shp = []
        for v in a[:2]:
            if torch.is_tensor(v): shp.append(tuple(v.shape))
        rec[(name, tuple(shp), k.get("act", 0), k.get("stride", 1))].append((s, e))
        return r
    setattr(ops, name, g)
for n in ("gemm", "conv3x3", "attention", "groupnorm", "layernorm", "upsample_nearest", "conv3x3_small_in", "conv3x3_small_out",
          "linear_smallm", "batch_axis_attention", "nchw_to_nhwc", "timestep_embedding"):
    wrap(n)
import faithdiff_dev_b200.blocks as B, faithdiff_dev_b200.unet as U
out = unet(x, torch.tensor(951.), **kw)[0]
torch.cuda.synchronize()
rows = []
for k, evs in rec.items():
    t = sum(s.elapsed_time(e) for s, e in evs)
    rows.append((t, len(evs), k))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("sum of per-op event times %.2f ms" % tot)
bycls = collections.defaultdict(float)
for t, n, k in rows: bycls[k[0]] += t
for k, t in sorted(bycls.items(), key=lambda kv: -kv[1]): print("  %-22s %7.2f ms %5.1f%%" % (k, t, 100 * t / tot))
def flops(k):
    name, shp = k[0], k[1]
    if name == "gemm": return 2.0 * shp[0][0] * shp[0][1] * shp[1][0] * (shp[1][1] / shp[0][1])
    if name == "conv3x3":
        b, h, w, c = shp[0]; n, kk = shp[1]
        s = k[3]
        return 2.0 * b * (h // s) * (w // s) * n * kk
    return 0
for t, n, k in rows[:40]:
    f = flops(k)
    print("  %7.3f ms x%3d %-10s %-44s act=%d %s" % (t, n, k[0], str(k[1]), k[2], ("%.0f TF/s" % (f * n / t / 1e9)) if f else ""))
