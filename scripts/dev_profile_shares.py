"""Dev (GPU): UNFILTERED share of every kernel — ours (fd::) and torch's / NCCL's glue — in one whole 4K restoration, from CUPTI
kernel activity records (torch.profiler; near-native speed, covers kernels inside CUDA-graph replays), since an ncu launch list of
a whole 20-step restoration (~90 k launches) is hours.  usage: dev_profile_shares.py [workload=4k] [denoise_steps=20]"""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from PIL import Image
import bench
import faithdiff_dev_b200 as fd
from faithdiff_dev_b200 import arch
from faithdiff_dev_b200.color_fix import color_fix_u8
wl = bench.workload(sys.argv[1] if len(sys.argv) > 1 else "4k")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else wl["steps"]
unet = fd.UNet2DConditionModel(**arch.SDXL_UNET_CONFIG)
unet.load_state_dict(arch.random_state_dict(arch.unet_param_shapes(arch.SDXL_UNET_CONFIG), seed=1234))
vae = fd.AutoencoderKL(**arch.SDXL_VAE_CONFIG)
vae.load_state_dict(arch.random_state_dict(arch.vae_param_shapes(), seed=4321))
pipe = fd.FaithDiffStableDiffusionXLPipeline(vae, unet.denoise_encoder, None, None, None, None, unet, fd.EulerDiscreteScheduler(), fd.DDPMScheduler())
pipe.set_encoder_tile_settings(); pipe.enable_vae_tiling()
W, H = wl["lq"][0] * wl["scale"], wl["lq"][1] * wl["scale"]
dev_u8 = torch.from_numpy(np.array(bench.synthetic_lq(*wl["lq"]).resize((W, H), Image.LANCZOS), dtype=np.uint8)).cuda()
g = torch.Generator().manual_seed(7)
pe, npe = torch.randn(1, 77, 2048, generator=g).cuda(), torch.randn(1, 77, 2048, generator=g).cuda()
pp, npp = torch.randn(1, 1280, generator=g).cuda(), torch.randn(1, 1280, generator=g).cuda()


def run():
    torch.manual_seed(42)
    out = pipe(lr_img=dev_u8, prompt_embeds=pe, negative_prompt_embeds=npe, pooled_prompt_embeds=pp, negative_pooled_prompt_embeds=npp,
               num_inference_steps=steps, start_point="lr", height=H, width=W, overlap=0.5, guidance_scale=5.0, target_size=(1024, 1024),
               generator=torch.Generator(device="cuda").manual_seed(42), output_type="u8").images[0]
    return color_fix_u8(out, dev_u8, "adain")


run(); run()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        t = getattr(ev, "device_time_total", None) or getattr(ev, "cuda_time_total", 0.0)
        tot[ev.name][0] += 1
        tot[ev.name][1] += t
S = sum(v[1] for v in tot.values())
ours = sum(v[1] for k, v in tot.items() if "fd::" in k)
print("workload %s, %d scheduler steps: %d kernel records, %.1f ms of kernel time; fd:: kernels %.2f %%, everything else (torch copies / "
      "cat / index_select / fill / memcpy, NCCL) %.2f %%" % (wl["name"][:40], steps, sum(v[0] for v in tot.values()), S / 1e3, 100 * ours / S,
                                                               100 * (S - ours) / S))
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
    print("  %6.2f %%  x%7d  %10.1f us  %s" % (100 * t / S, c, t, k[:150]))
