"""Factory mirroring FaithDiff/create_FaithDiff_model.py:11-37 (`FaithDiff_pipeline`) for local checkpoint files: the SDXL
UNet (`unet/diffusion_pytorch_model*.safetensors`), the fp16-fix VAE and `FaithDiff.bin` load unchanged because the key layout
is the reference's (arch.py).  `use_fp8` reproduces utils/system.py:13-32 `quantize_8bit`: UNet parameters (not the LQ
Encoder) are rounded through float8_e4m3fn — storage in the reference, compute stays 16-bit there as here."""
import glob
import os

import torch

from . import arch
from .pipeline import FaithDiffStableDiffusionXLPipeline
from .scheduler import DDPMScheduler, EulerDiscreteScheduler
from .text_encoder import CLIPTextEncoder
from .unet import UNet2DConditionModel
from .bsrnet import create_bsrnet  # noqa: F401  (create_FaithDiff_model.py:39-42)
from .vae import AutoencoderKL


def _load_file(path):
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=False)


def _find(dirpath, patterns):
    for p in patterns:
        hits = sorted(glob.glob(os.path.join(dirpath, p)))
        if hits:
            return hits[0]
    raise FileNotFoundError("none of %s under %s" % (patterns, dirpath))


def load_text_encoder(dirpath):
    """`<sdxl>/text_encoder{,_2}` (transformers layout: config.json + model*.safetensors) → CLIPTextEncoder on B200 kernels;
    None when the directory is absent (the pipeline then needs prompt_embeds)."""
    import json
    cfg_file = os.path.join(dirpath, "config.json")
    if not os.path.isfile(cfg_file):
        return None
    cfg = json.load(open(cfg_file))
    with_proj = "WithProjection" in "".join(cfg.get("architectures") or [])
    enc = CLIPTextEncoder(vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"], intermediate_size=cfg["intermediate_size"],
                          num_hidden_layers=cfg["num_hidden_layers"], num_attention_heads=cfg["num_attention_heads"],
                          max_position_embeddings=cfg.get("max_position_embeddings", 77), hidden_act=cfg.get("hidden_act", "quick_gelu"),
                          layer_norm_eps=cfg.get("layer_norm_eps", 1e-5), eos_token_id=cfg.get("eos_token_id", 2),
                          projection_dim=cfg.get("projection_dim") if with_proj else None)
    return enc.load_state_dict(_load_file(_find(dirpath, ["model.fp16.safetensors", "model*.safetensors", "*.bin"])), strict=False)


def load_tokenizer(dirpath):
    """CLIP BPE tokenizer files are plain text/json; the tokenizer itself is host-side string processing (transformers)."""
    if not os.path.isdir(dirpath):
        return None
    try:
        from transformers import CLIPTokenizer
        return CLIPTokenizer.from_pretrained(dirpath)
    except Exception:
        return None


def quantize_8bit_state_dict(sd):
    """utils/system.py:13-32: every UNet tensor → float8_e4m3fn (unscaled) except the `denoise_encoder.` module."""
    out = {}
    for k, v in sd.items():
        if k.startswith("denoise_encoder.") or not v.is_floating_point():
            out[k] = v
        else:
            out[k] = v.to(torch.float8_e4m3fn).to(v.dtype)
    return out


def FaithDiff_pipeline(sdxl_path, VAE_FP16_path, FaithDiff_path, use_fp8=False, unet_config=None, vae_config=None):
    vae = AutoencoderKL(**(vae_config or arch.SDXL_VAE_CONFIG))
    vae.load_state_dict(_load_file(_find(VAE_FP16_path, ["diffusion_pytorch_model*.safetensors", "*.safetensors", "*.bin"])))
    unet = UNet2DConditionModel(**(unet_config or arch.SDXL_UNET_CONFIG))
    sd = dict(_load_file(_find(os.path.join(sdxl_path, "unet"), ["diffusion_pytorch_model.fp16.safetensors",
                                                                  "diffusion_pytorch_model*.safetensors", "*.bin"])))
    fd_file = FaithDiff_path if os.path.isfile(FaithDiff_path) else _find(FaithDiff_path, ["FaithDiff.bin", "*.bin"])
    sd.update(_load_file(fd_file))           # FaithDiff.bin holds the fine-tuned UNet + the FaithDiff modules (README.md:85-93)
    if use_fp8:
        sd = quantize_8bit_state_dict(sd)
    unet.load_state_dict(sd, strict=True)
    pipe = FaithDiffStableDiffusionXLPipeline(vae=vae, denoise_encoder=unet.denoise_encoder,
                                              text_encoder=load_text_encoder(os.path.join(sdxl_path, "text_encoder")),
                                              text_encoder_2=load_text_encoder(os.path.join(sdxl_path, "text_encoder_2")),
                                              tokenizer=load_tokenizer(os.path.join(sdxl_path, "tokenizer")),
                                              tokenizer_2=load_tokenizer(os.path.join(sdxl_path, "tokenizer_2")), unet=unet,
                                              scheduler=EulerDiscreteScheduler(), DDPM_scheduler=DDPMScheduler())
    return pipe
