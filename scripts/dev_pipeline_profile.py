"""Dev (GPU): full-size pipeline timing (SDXL-arch random weights). usage: dev_pipeline_profile.py W H steps nb [graph]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import faithdiff_dev_b200 as fd
from faithdiff_dev_b200 import ops, arch
W, H, steps, nb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
use_graph = (len(sys.argv) < 6) or sys.argv[5] != "nograph"
BF = torch.bfloat16
unet = fd.UNet2DConditionModel(**arch.SDXL_UNET_CONFIG); unet.load_state_dict(arch.random_state_dict(arch.unet_param_shapes(arch.SDXL_UNET_CONFIG)))
vae = fd.AutoencoderKL(**arch.SDXL_VAE_CONFIG); vae.load_state_dict(arch.random_state_dict(arch.vae_param_shapes(), seed=4321))
pipe = fd.FaithDiffStableDiffusionXLPipeline(vae, unet.denoise_encoder, None, None, None, None, unet, fd.EulerDiscreteScheduler(), fd.DDPMScheduler())
pipe.set_encoder_tile_settings(); pipe.enable_vae_tiling()
pipe.tiles_per_launch, pipe.use_cuda_graph = nb, use_graph
g = torch.Generator().manual_seed(7)
pe, npe = torch.randn(1, 77, 2048, generator=g), torch.randn(1, 77, 2048, generator=g)
pp, npp = torch.randn(1, 1280, generator=g), torch.randn(1, 1280, generator=g)
rng = np.random.default_rng(0)
img = torch.from_numpy(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).pin_memory()
for it in range(int(os.environ.get("FD_CALLS", "3"))):
    torch.cuda.synchronize(); t0 = time.time(); n0 = ops.launch_count()
    out = pipe(lr_img=img, prompt_embeds=pe, negative_prompt_embeds=npe, pooled_prompt_embeds=pp, negative_pooled_prompt_embeds=npp,
               num_inference_steps=steps, start_point="lr", height=H, width=W, overlap=0.5, guidance_scale=5.0, target_size=(1024, 1024),
               generator=torch.Generator(device="cuda").manual_seed(42), output_type="pil")
    torch.cuda.synchronize(); dt = time.time() - t0
    print("call %d: %.2f s  -> %.3f MP/s  stages(ms) %s  host launches %d  mem %.1f GB" % (
        it, dt, W * H / 1e6 / dt, {k: round(v) for k, v in pipe.stage_times_ms().items()}, ops.launch_count() - n0,
        torch.cuda.max_memory_allocated() / 1e9), flush=True)
a = np.asarray(out.images[0]); print("image", a.shape, a.dtype, a.mean(), a.std())
T = len(pipe.tlc_vae_latents.idxes)
print("tiles", T, "=> denoise ms/step/tile %.2f" % (pipe.stage_times_ms()["denoise"] / steps / T))
