"""Dev (torchrun, N GPUs): N-rank pipeline result == 1-rank result (tiny config) + timing of the full-size 4K path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
import faithdiff_dev_b200 as fd
from faithdiff_dev_b200 import arch
from oracle import faithdiff_oracle as O
BF = torch.bfloat16
cfg = O.TINY_UNET_CONFIG
sd = arch.random_state_dict(arch.unet_param_shapes(cfg, lq_channels=64, encoder_channels=(32, 64, 64, 64)), seed=1, device="cuda")
unet = fd.UNet2DConditionModel(**cfg); unet.load_state_dict(sd)
vae = fd.AutoencoderKL(**O.TINY_VAE_CONFIG); vae.load_state_dict(arch.random_state_dict(arch.vae_param_shapes(dict(O.TINY_VAE_CONFIG, layers_per_block=2, latent_channels=4)), seed=2))
pipes = []
def make_pipe(dist):
    p = fd.FaithDiffStableDiffusionXLPipeline(vae, unet.denoise_encoder, None, None, None, None, unet, fd.EulerDiscreteScheduler(), fd.DDPMScheduler())
    if dist: p.set_distributed(rank, world)
    else: p.set_distributed(0, 1)
    pipes.append(p)
    return p
for e in (unet.denoise_encoder,):
    e.tile_sample_min_size, e.tile_latent_min_size, e.use_tiling = 128, 16, True
vae.enable_tiling()
g = torch.Generator().manual_seed(7)
pe, npe = torch.randn(1, 77, 256, generator=g), torch.randn(1, 77, 256, generator=g)
pp, npp = torch.randn(1, 64, generator=g), torch.randn(1, 64, generator=g)
lr = O.synthetic_lq_image(328, 232, seed=0)
lat = torch.randn(1, 4, 29, 41, generator=g).to(BF).cuda()
def run(p, seed, given_latents):
    # RNG contract (pipeline.set_distributed): ranks may be seeded differently — rank 0's latents / posterior sample / DDPM noise
    # are broadcast, so the N-rank run below (seed 3 + rank, latents drawn inside) must equal rank 0's 1-rank run with seed 3
    torch.manual_seed(seed)
    return p(lr_img=lr, prompt_embeds=pe, negative_prompt_embeds=npe, pooled_prompt_embeds=pp, negative_pooled_prompt_embeds=npp,
             num_inference_steps=3, start_point="lr", height=232, width=328, overlap=0.5, guidance_scale=5.0, target_size=(128, 128),
             latents=lat.clone() if given_latents else None, generator=torch.Generator(device="cuda").manual_seed(seed),
             output_type="u8").images[0]
allp = []
for given in (True, False):
    a = run(make_pipe(False), 3, given).float()                    # 1-rank reference (every rank computes rank 0's: seed 3)
    b = run(make_pipe(True), 3 + (0 if given else rank), given).float()
    mse = ((a - b) ** 2).mean().item()
    psnr = 99.0 if mse == 0 else 10 * np.log10(255 ** 2 / mse)
    got = [None] * world
    torch.distributed.all_gather_object(got, (psnr, float((a - b).abs().max()), float(b.sum())))
    allp.append(got)
if rank == 0:
    print("world=%d: PSNR(N-rank vs 1-rank) per rank [latents given | drawn with per-rank seeds]:" % world, allp, flush=True)
    for got in allp:
        assert min(p for p, _, _ in got) > 45, allp
        assert len({c for _, _, c in got}) == 1, ("ranks disagree on the N-rank image", got)
torch.distributed.barrier()
for p in pipes:
    p.shutdown()
torch.distributed.destroy_process_group()
