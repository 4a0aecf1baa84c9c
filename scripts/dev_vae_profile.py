"""Dev (GPU): per-op timing of the tiled LQ encoder / VAE encode / VAE decode stages at the 4K workload's sizes."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import faithdiff_dev_b200 as fd
from faithdiff_dev_b200 import ops, arch
import faithdiff_dev_b200.ops as O
BF = torch.bfloat16
H, W = int(os.environ.get("FD_H", 2160)), int(os.environ.get("FD_W", 3840))
unet = fd.UNet2DConditionModel(**arch.SDXL_UNET_CONFIG); unet.load_state_dict(arch.random_state_dict(arch.unet_param_shapes(arch.SDXL_UNET_CONFIG)))
vae = fd.AutoencoderKL(**arch.SDXL_VAE_CONFIG); vae.load_state_dict(arch.random_state_dict(arch.vae_param_shapes(), seed=4321))
enc = unet.denoise_encoder
enc.tile_sample_min_size, enc.tile_overlap_factor = 1024, 0.25
enc.enable_tiling(); vae.enable_tiling()
x = (torch.rand(1, 3, H, W, device="cuda") * 2 - 1).to(BF)
z = torch.randn(1, 4, H // 8, W // 8, device="cuda").to(BF)
stages = {"lq_encoder": lambda: enc(x), "vae_encode": lambda: vae.encode(x).latent_dist.mode(), "vae_decode": lambda: vae.decode(z, return_dict=False)[0]}
for _ in range(2):
    for name, fn in stages.items():
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        print("%-12s %.1f ms" % (name, e0.elapsed_time(e1)), flush=True)
rec = collections.defaultdict(list)
names = [n for n in dir(O) if callable(getattr(O, n)) and not n.startswith("_") and n not in ("check", "launch_count")]
cur = [None]
def wrap(name):
    f = getattr(O, name)
    if not callable(f) or isinstance(f, type): return
    def g(*a, **k):
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record(); r = f(*a, **k); e.record()
        shp = tuple(tuple(v.shape) for v in a[:2] if torch.is_tensor(v))
        rec[(cur[0], name, shp)].append((s, e))
        return r
    setattr(O, name, g)
for n in ("gemm", "conv3x3", "attention", "layernorm", "groupnorm", "softmax_rows", "upsample_nearest", "nchw_to_nhwc", "nhwc_to_nchw",
          "seam_blend", "conv3x3_small", "conv1x1_small", "scale_div"):
    if hasattr(O, n): wrap(n)
for name, fn in stages.items():
    cur[0] = name
    torch.cuda._sleep(int(2e8)); fn(); torch.cuda.synchronize()
for st in stages:
    rows = []
    for (s_, name, shp), evs in rec.items():
        if s_ != st: continue
        rows.append((sum(a.elapsed_time(b) for a, b in evs), len(evs), name, shp))
    rows.sort(reverse=True)
    print("== %s: instrumented op total %.1f ms" % (st, sum(r[0] for r in rows)))
    for t, n, name, shp in rows[:14]:
        print("  %8.2f ms x%4d (%8.1f us) %-16s %s" % (t, n, t / n * 1e3, name, shp))
