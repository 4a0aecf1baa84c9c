"""B200-native drop-in for `FaithDiffStableDiffusionXLPipeline` (FaithDiff/pipelines/pipeline_FaithDiff_tlc.py:374-1459):
same constructor components and `__call__` signature, same tile layout / blend / scheduler semantics; the hot loop is
re-organised for the hardware (step-invariant work hoisted, CFG-identical stem evaluated once, per-step host logic
reduced to three scalars, the whole per-step tail fused into one kernel, tiles sharded over ranks with one NCCL
all-reduce of the fp32 overlap accumulation per scheduler step)."""
import os
from dataclasses import dataclass
from typing import List, Optional, Union

import numpy as np
import PIL.Image
import torch

from . import ops
from .blocks import BF16
from .dist import exchange_canvas, plan_launch_groups, plan_unit_groups, shard_tiles, shard_units
from .tiling import LocalAttention, gaussian_weights, tile_origins


@dataclass
class FaithDiffStableDiffusionXLPipelineOutput:
    """FaithDiff/pipelines/pipeline_output.py:10-21."""
    images: Union[List[PIL.Image.Image], np.ndarray]


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class TileStepGraph:
    """One CUDA graph = the UNet evaluation of a launch group of (tile, CFG-branch) samples.  The FaithDiff stem runs once per
    distinct tile of the group (the CFG halves are identical there, pipeline_FaithDiff_tlc.py:1334-1335); its output rows
    are gathered into the trunk batch (one row per sample).  `slot_of_sample[s]` = index of sample s's tile among the group's
    tiles, `branch_of_sample[s]` = 0 (uncond) / 1 (text): both are baked into the graph, all tensors are static buffers that
    are refilled before each replay.  Removes ~1000 host launches per tile-step."""

    def __init__(self, unet, kv, lk, slot_of_sample, branch_of_sample, k1, k2, lowp, add_sample=True, use_graph=True):
        dev = unet.device
        c0 = unet.conv_in[0].shape[0]
        self.unet, self.add_sample, self.lk = unet, add_sample, lk
        self.slot_of_sample, self.branch_of_sample = tuple(slot_of_sample), tuple(branch_of_sample)
        self.ns, self.nt = len(self.slot_of_sample), max(self.slot_of_sample) + 1
        self.paired = self.slot_of_sample == tuple(i // 2 for i in range(self.ns)) and self.ns % 2 == 0
        self.gather_idx = torch.tensor(self.slot_of_sample, device=dev, dtype=torch.long)
        self.scaled = torch.zeros((self.nt, unet.config.in_channels, k1, k2), device=dev, dtype=lowp)
        # stem input [conv_in(sample) | condition embedding]: the cached, step-invariant embedding of the group's tiles is
        # written straight into its channel half (one strided copy per replay; launch groups share a graph, so it cannot be
        # baked in), conv_in writes the other half inside the graph
        self.cat = torch.zeros((self.nt, k1, k2, 2 * c0), device=dev, dtype=BF16)
        self.c0 = c0
        self.temb = torch.zeros((self.ns, unet.temb_w.shape[0]), device=dev, dtype=torch.float32)
        # static text K/V per sample (the attention kernel indexes K/V rows by sample); a new prompt is copied in with
        # set_text_kv() — no re-capture
        self.kv = [[torch.zeros((self.ns * lk, b.shape[1]), device=dev, dtype=b.dtype) for b in tr] for tr in kv]
        self.set_text_kv(kv)
        self.graph = None
        self.replays = 0
        n0 = ops.launch_count()
        self._run()                               # warm-up (sets kernel attributes, fills allocator pools)
        self.kernels_per_replay = ops.launch_count() - n0
        torch.cuda.current_stream().synchronize()
        if use_graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run()
            self.graph = g

    def set_text_kv(self, kv):
        """kv: per transformer block [2*lk, 2C] (rows [0,lk) = uncond text, [lk,2lk) = cond text)."""
        lk = self.lk
        for dst_tr, src_tr in zip(self.kv, kv):
            for dst, src in zip(dst_tr, src_tr):
                for s_, b_ in enumerate(self.branch_of_sample):
                    dst[s_ * lk:(s_ + 1) * lk].copy_(src[b_ * lk:(b_ + 1) * lk])

    def _run(self):
        u = self.unet
        x1 = u.stem(self.scaled, None, add_sample=self.add_sample, groups=self.nt, cat=self.cat)  # [nt,h,w,c0]
        if self.paired:
            x2 = x1.unsqueeze(1).expand(-1, 2, -1, -1, -1).reshape((self.ns,) + tuple(x1.shape[1:])).contiguous()
        else:
            x2 = x1.index_select(0, self.gather_idx)
        st = getattr(x1, "_fd_colstats", None)          # GroupNorm statistics of the stem output follow the sample gather
        if st is not None:
            ops.tag_stats(x2, st.select_samples(self.gather_idx, self.nt))
        self.eps = u.trunk(x2, self.temb, self.kv, self.lk, self.scaled.dtype)                  # [ns,4,k1,k2]

    def __call__(self, scaled_tiles, cond_tiles, temb_rows, eps_out):
        """scaled_tiles [nt,4,k1,k2] and cond_tiles [nt,k1,k2,c0] (the group's tile slots, in slot order); temb_rows [ns, ΣC];
        eps_out [ns,4,k1,k2] destination view."""
        self.scaled.copy_(scaled_tiles)
        self.cat[..., self.c0:].copy_(cond_tiles)
        self.temb.copy_(temb_rows)
        if self.graph is not None:
            self.graph.replay()
            self.replays += 1
        else:
            self._run()
        eps_out.copy_(self.eps)


class FaithDiffStableDiffusionXLPipeline:
    model_cpu_offload_seq = "text_encoder->text_encoder_2->denoise_encoder->unet->vae"

    def __init__(self, vae, denoise_encoder, text_encoder, text_encoder_2, tokenizer, tokenizer_2, unet, scheduler,
                 DDPM_scheduler, force_zeros_for_empty_prompt=True, add_watermarker=None):
        self.vae, self.denoise_encoder = vae, denoise_encoder
        self.text_encoder, self.text_encoder_2 = text_encoder, text_encoder_2
        self.tokenizer, self.tokenizer_2 = tokenizer, tokenizer_2
        self.unet, self.scheduler, self.DDPM_scheduler = unet, scheduler, DDPM_scheduler
        self.config = _Cfg(force_zeros_for_empty_prompt=force_zeros_for_empty_prompt)
        self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1)
        self.default_sample_size = self.unet.config.sample_size
        self.watermark = None
        self.device = torch.device("cuda")
        self.rank, self.world_size, self.process_group = 0, 1, None
        self._dist = None
        self.tiles_per_launch = 8          # upper bound; the launch plan mixes group sizes 1..N by measured cost
        # dtype of the latent state x between scheduler steps.  The reference casts it to the model dtype every step
        # (pipeline…:1373 `.to(model_output.dtype)`, :1409) — fp16 there.  In bf16 (3 fewer mantissa bits) those casts alone let
        # the iterated state drift to rel-L2 1.2e-2 from the fp32 result over 20 steps (profiles/r02_parity_unnormalised_eps.json),
        # so the 2 MB state stays in fp32 by default; None reproduces the reference's per-step casts bit for bit.
        self.latent_state_dtype = torch.float32
        self.use_cuda_graph = True
        self._graphs = {}

    def shutdown(self):
        """Releases the opt-in peer-memory exchange region (collective over the ranks; a no-op otherwise)."""
        if getattr(self, "_peer", None) is not None:
            self._peer.close()
            self._peer, self._peer_key = None, None

    # ------------------------------------------------------------------ reference-facing knobs
    def to(self, device=None, *a, **k):
        if device is not None and torch.device(device).type != "cuda":
            raise RuntimeError("faithdiff_dev_b200 has no CPU path")
        return self

    def set_progress_bar_config(self, **kw):
        pass

    def enable_model_cpu_offload(self, *a, **k):
        raise NotImplementedError("cpu offload is not supported: the B200 path keeps all weights resident in HBM")

    def set_encoder_tile_settings(self, denoise_encoder_tile_sample_min_size=1024, denoise_encoder_sample_overlap_factor=0.25,
                                  vae_sample_size=1024, vae_tile_overlap_factor=0.25):
        """pipeline…:888-896 (including the quirk that vae.config.sample_size does not retune the VAE tile size)."""
        self.denoise_encoder.tile_sample_min_size = denoise_encoder_tile_sample_min_size
        self.denoise_encoder.tile_overlap_factor = denoise_encoder_sample_overlap_factor
        self.vae.config.sample_size = vae_sample_size
        self.vae.tile_overlap_factor = vae_tile_overlap_factor

    def enable_vae_tiling(self):
        self.vae.enable_tiling()
        self.denoise_encoder.enable_tiling()

    def disable_vae_tiling(self):
        self.vae.disable_tiling()
        self.denoise_encoder.disable_tiling()

    def set_distributed(self, rank, world_size, process_group=None):
        """Multi-GPU: (tile, CFG-branch) units in contiguous balanced ranges per rank, one exchange of the fp32 canvas per
        scheduler step; the tiled LQ-encoder / VAE stages shard their tiles by area.  The context lives on THIS pipeline (it is
        handed to the encoder / VAE per call), so pipelines of one process do not share it.
        RNG contract: the initial latents (and, with start_point='lr', the posterior sample and the DDPM noise) are drawn on
        every rank and then REPLACED by rank 0's values (one broadcast per image), so ranks may be seeded differently."""
        self.rank, self.world_size, self.process_group = rank, world_size, process_group
        from .vae import DistCtx
        self._dist = DistCtx(rank, world_size, process_group) if world_size > 1 else None

    def check_inputs(self, lr_img, prompt, height, width, prompt_embeds, negative_prompt, negative_prompt_embeds,
                     pooled_prompt_embeds, negative_pooled_prompt_embeds):
        """pipeline…:737-817 (the rules that apply to this call path)."""
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError("Cannot forward both `prompt` and `prompt_embeds`. Please make sure to only forward one of the two.")
        if prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`. Cannot leave both `prompt` and `prompt_embeds` undefined.")
        if prompt is not None and not isinstance(prompt, (str, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if negative_prompt is not None and negative_prompt_embeds is not None:
            raise ValueError("Cannot forward both `negative_prompt` and `negative_prompt_embeds`.")
        if prompt_embeds is not None and pooled_prompt_embeds is None:
            raise ValueError("If `prompt_embeds` are provided, `pooled_prompt_embeds` also have to be passed.")
        if negative_prompt_embeds is not None and negative_pooled_prompt_embeds is None:
            raise ValueError("If `negative_prompt_embeds` are provided, `negative_pooled_prompt_embeds` also have to be passed.")

    def encode_prompt(self, prompt=None, prompt_2=None, device=None, num_images_per_prompt=1, do_classifier_free_guidance=True,
                      negative_prompt=None, negative_prompt_2=None, prompt_embeds=None, negative_prompt_embeds=None,
                      pooled_prompt_embeds=None, negative_pooled_prompt_embeds=None, lora_scale=None, clip_skip=None,
                      input_ids=None, input_ids_2=None, negative_input_ids=None, negative_input_ids_2=None):
        """SDXL prompt encoding, pipeline…:485-717: both text encoders over the (padded, truncated) 77 token ids,
        `hidden_states[-2]` (or `-(clip_skip+2)`) of each concatenated along channels, pooled output of the LAST encoder;
        negative branch = zeros when no negative prompt is given and `force_zeros_for_empty_prompt` (SDXL default), else
        the encoded negative prompt ("" by default).  Superset of the reference: pre-computed `prompt_embeds` are honoured
        (the reference ignores them, :1237-1243) and token ids can be passed directly (`input_ids*`) so no tokenizer is
        needed.  The encoders are `faithdiff_dev_b200.CLIPTextEncoder` (B200 kernels) or anything with the same call."""
        if lora_scale is not None:
            raise NotImplementedError("encode_prompt: LoRA scaling of the text encoders")
        encoders = [e for e in (self.text_encoder, self.text_encoder_2) if e is not None]
        tokenizers = [t for t in (self.tokenizer, self.tokenizer_2)] if self.text_encoder is not None else [self.tokenizer_2]
        if prompt_embeds is None and not encoders:
            raise RuntimeError("encode_prompt: the pipeline was built without text encoders; pass prompt_embeds / "
                               "pooled_prompt_embeds, or construct it with faithdiff_dev_b200.CLIPTextEncoder instances")

        def ids_for(which, text, given, max_length=None):
            if given is not None:
                return torch.as_tensor(given).reshape(-1, torch.as_tensor(given).shape[-1]).to(torch.int64)
            tok = tokenizers[which]
            if tok is None:
                raise RuntimeError("encode_prompt: no tokenizer %d; pass input_ids instead of text" % (which + 1))
            return tok(text, padding="max_length", max_length=max_length or tok.model_max_length, truncation=True,
                       return_tensors="pt").input_ids

        def run(texts, given_ids, skip, max_length=None):
            embeds, pooled = [], None
            for which, enc in enumerate(encoders):
                out = enc(ids_for(which, texts[which], given_ids[which], max_length), output_hidden_states=True)
                pooled = out[0]                                  # only the final encoder's survives (pipeline…:608)
                embeds.append(out.hidden_states[-2] if skip is None else out.hidden_states[-(skip + 2)])
            return torch.cat(embeds, dim=-1), pooled

        if isinstance(prompt, str):
            prompt = [prompt]
        if prompt is not None:
            batch_size = len(prompt)
        elif prompt_embeds is not None:
            batch_size = prompt_embeds.shape[0]
        else:
            batch_size = torch.as_tensor(input_ids).reshape(-1, torch.as_tensor(input_ids).shape[-1]).shape[0]
        if prompt_embeds is None:
            p2 = prompt_2 or prompt
            p2 = [p2] if isinstance(p2, str) else p2
            ids2 = input_ids_2 if input_ids_2 is not None else input_ids
            prompt_embeds, pooled_prompt_embeds = run([prompt, p2], [input_ids, ids2], clip_skip)
        zero_neg = negative_prompt is None and negative_input_ids is None and self.config.get("force_zeros_for_empty_prompt", True)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            if zero_neg:
                negative_prompt_embeds = torch.zeros_like(prompt_embeds)
                negative_pooled_prompt_embeds = torch.zeros_like(pooled_prompt_embeds)
            else:
                negative_prompt = negative_prompt or ""
                negative_prompt_2 = negative_prompt_2 or negative_prompt
                neg = batch_size * [negative_prompt] if isinstance(negative_prompt, str) else negative_prompt
                neg2 = batch_size * [negative_prompt_2] if isinstance(negative_prompt_2, str) else negative_prompt_2
                if prompt is not None and negative_input_ids is None and batch_size != len(neg):
                    raise ValueError(f"`negative_prompt`: {negative_prompt} has batch size {len(neg)}, but `prompt`: {prompt} has "
                                     f"batch size {batch_size}.")
                nids2 = negative_input_ids_2 if negative_input_ids_2 is not None else negative_input_ids
                negative_prompt_embeds, negative_pooled_prompt_embeds = run([neg, neg2], [negative_input_ids, nids2], None,
                                                                            max_length=prompt_embeds.shape[1])
        dt = BF16
        bs, seq, _ = prompt_embeds.shape
        prompt_embeds = prompt_embeds.to(self.device, dt).repeat(1, num_images_per_prompt, 1).view(bs * num_images_per_prompt, seq, -1)
        pooled_prompt_embeds = pooled_prompt_embeds.to(self.device, dt).repeat(1, num_images_per_prompt).view(bs * num_images_per_prompt, -1)
        if do_classifier_free_guidance:
            seq = negative_prompt_embeds.shape[1]
            negative_prompt_embeds = negative_prompt_embeds.to(self.device, dt).repeat(1, num_images_per_prompt, 1).view(
                batch_size * num_images_per_prompt, seq, -1)
            negative_pooled_prompt_embeds = negative_pooled_prompt_embeds.to(self.device, dt).repeat(1, num_images_per_prompt).view(
                bs * num_images_per_prompt, -1)
        return prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds

    # ------------------------------------------------------------------ stages
    def _preprocess(self, lr_img, height, width):
        """VaeImageProcessor.preprocess (pipeline…:1246): host u8 → pinned → H2D → kernel."""
        if isinstance(lr_img, PIL.Image.Image):
            w, h = width - width % 8, height - height % 8
            arr = np.asarray(lr_img.convert("RGB"), dtype=np.uint8)
            host = torch.from_numpy(np.ascontiguousarray(arr))
            if lr_img.size != (w, h):
                # VaeImageProcessor.resize (PIL LANCZOS) on the GPU, bit-identical to PIL (fd_resample_u8_*): the smaller
                # source image is what crosses PCIe
                from .image_process import resize_lanczos
                self._h2d_bytes = host.numel()
                dev = resize_lanczos(host.pin_memory(), (w, h), device=self.device)
                self._last_input_u8 = dev
                return ops.image_preprocess(dev)[0], dev
        else:
            host = lr_img            # HWC uint8 tensor (host, ideally pinned)
        if not host.is_cuda:
            if not host.is_pinned():
                host = host.pin_memory()
            self._h2d_bytes = host.numel()
            dev = host.to(self.device, non_blocking=True)
        else:
            self._h2d_bytes = 0
            dev = host
        self._last_input_u8 = dev
        nhwc, _ = ops.image_preprocess(dev)
        return nhwc, dev

    def _encode_lq(self, lr_nhwc):
        """prepare_image_latents (pipeline…:950-1004) → denoise_encoder (LQ features), NHWC bf16 [1,h,w,512]."""
        enc = self.denoise_encoder
        x_nchw = ops.nhwc_to_nchw(lr_nhwc, BF16)
        feats = enc(x_nchw, dist=self._dist) if self._dist is not None else enc(x_nchw)
        return ops.nchw_to_nhwc(feats.contiguous()), x_nchw

    @torch.no_grad()
    def __call__(self, lr_img=None, prompt=None, prompt_2=None, height=None, width=None, num_inference_steps=50,
                 start_point="noise", timesteps=None, denoising_end=None, overlap=0.5, guidance_scale=5.0,
                 negative_prompt=None, negative_prompt_2=None, num_images_per_prompt=1, eta=0.0, generator=None,
                 latents=None, prompt_embeds=None, negative_prompt_embeds=None, pooled_prompt_embeds=None,
                 negative_pooled_prompt_embeds=None, output_type="pil", return_dict=True, cross_attention_kwargs=None,
                 guidance_rescale=0.0, original_size=None, target_size=None, clip_skip=None, callback_on_step_end=None,
                 callback_on_step_end_tensor_inputs=("latents",), add_sample=True, **kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError("faithdiff_dev_b200 has no CPU path: a CUDA device is required")
        if timesteps is not None or cross_attention_kwargs or guidance_rescale:
            raise NotImplementedError("custom timesteps / cross_attention_kwargs / guidance_rescale")
        if guidance_scale <= 1:
            raise NotImplementedError("the hot path implements the classifier-free-guidance branch (guidance_scale > 1)")
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
        original_size = original_size or (height, width)
        target_size = target_size or (height, width)
        self.check_inputs(lr_img, prompt, height, width, prompt_embeds, negative_prompt, negative_prompt_embeds,
                          pooled_prompt_embeds, negative_pooled_prompt_embeds)
        kv_src = (prompt_embeds, negative_prompt_embeds) if prompt_embeds is not None and negative_prompt_embeds is not None else None
        # always through encode_prompt: it returns given embeddings unchanged and builds the negative branch (zeros under
        # force_zeros_for_empty_prompt, else the encoded negative prompt) when only the positive embeddings were passed
        prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds = self.encode_prompt(
            prompt, prompt_2, negative_prompt=negative_prompt, negative_prompt_2=negative_prompt_2, clip_skip=clip_skip,
            num_images_per_prompt=num_images_per_prompt, prompt_embeds=prompt_embeds, negative_prompt_embeds=negative_prompt_embeds,
            pooled_prompt_embeds=pooled_prompt_embeds, negative_pooled_prompt_embeds=negative_pooled_prompt_embeds)
        dev, dt = self.device, BF16
        unet, vae, sch = self.unet, self.vae, self.scheduler
        ev = {k: torch.cuda.Event(enable_timing=True) for k in ("t0", "enc", "loop", "dec")}
        ev["t0"].record()

        # -- image → LQ features → tiles (once per image)
        lr_nhwc, _ = self._preprocess(lr_img, height, width)
        sch.set_timesteps(num_inference_steps, device="cpu")
        ts = sch.timesteps
        if denoising_end is not None and isinstance(denoising_end, float) and 0 < denoising_end < 1:
            # pipeline…:1297-1311: stop after the timesteps >= the cutoff (the returned sample keeps the remaining noise)
            n_train = sch.config.num_train_timesteps
            cutoff = int(round(n_train - denoising_end * n_train))
            ts = ts[:int((ts >= cutoff).sum())]
        lq_nhwc, lr_nchw = self._encode_lq(lr_nhwc)
        h, w = height // 8, width // 8
        k = (target_size[0] // 8, target_size[1] // 8)
        self.tlc_vae_latents = LocalAttention(k, overlap)
        self.tlc_vae_img = LocalAttention(k, overlap)
        idx, k1, k2 = tile_origins(h, w, k[0], k[1], overlap)
        T = len(idx)
        la = self.tlc_vae_latents

        # -- initial latents (torch RNG on purpose: seeds must reproduce the reference's Philox stream)
        shape = (1, vae.config.latent_channels, h, w)
        sdt = self.latent_state_dtype or dt
        if latents is None:
            gdev = generator.device if generator is not None else dev
            latents = torch.randn(shape, generator=generator, device=gdev, dtype=dt).to(dev)
        latents = latents.to(dev, sdt) * sch.init_noise_sigma
        if start_point == "lr":
            enc_out = vae.encode(lr_nchw, dist=self._dist) if self._dist is not None else vae.encode(lr_nchw)
            z0 = enc_out.latent_dist.sample().to(sdt) * vae.config.scaling_factor
            t0 = torch.randint(999, 1000, (1,), device=dev).long()
            latents = self.DDPM_scheduler.add_noise(z0[0:1], latents, t0)
        if self.world_size > 1:
            # every rank must blend the SAME x: rank 0's draw wins (see set_distributed)
            latents = latents.contiguous()
            src = torch.distributed.get_global_rank(self.process_group, 0) if self.process_group is not None else 0
            torch.distributed.broadcast(latents, src=src, group=self.process_group)
        tiles = la.grids(latents.contiguous())                      # [T,4,k1,k2]
        weights, origins = la.weights64, la.origins

        # -- step/tile-invariant conditioning
        pe = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0).to(dev, dt)
        te = torch.cat([negative_pooled_prompt_embeds, pooled_prompt_embeds], dim=0).to(dev, dt)
        # text K/V of the 70 cross-attention blocks: recomputed (and re-copied into the tile-step graphs) only when the prompt changed
        # (keyed on the IDENTITY + version counter of the caller's embedding tensors; the cache keeps them referenced, so their
        # storage cannot be recycled for different data)
        old = getattr(self, "_kv_src", None)
        wver = (unet, getattr(unet, "_weights_version", 0))       # compared by identity (the class defines no __eq__) + version
        kv_fresh = (kv_src is None or old is None or getattr(self, "_kv_unet", None) != wver
                    or any(a is not b for a, b in zip(kv_src, old[0])) or tuple(t._version for t in kv_src) != old[1])
        if kv_fresh:
            self._kv_val, self._kv_unet = unet.text_kv(pe), wver
            self._kv_src = None if kv_src is None else (kv_src, tuple(t._version for t in kv_src))
        kv, lk = self._kv_val
        # work units = (tile, CFG branch); contiguous balanced ranges per rank (a whole CFG pair per launch slot on 1 GPU)
        my_units = shard_units(T, self.rank, self.world_size)
        my_tiles = sorted({u // 2 for u in my_units})
        # ControlNetConditioningEmbedding: once per tile of this rank, kept as one tensor in tile order (launch groups are
        # contiguous tile ranges, so a group's embeddings are a slice)
        cond_all = torch.empty((len(my_tiles), k1, k2, unet.conv_in[0].shape[0]), device=dev, dtype=BF16)
        for n0_ in range(0, len(my_tiles), 8):                      # batches of 8 tiles: 3 launches per batch instead of per tile
            part = my_tiles[n0_:n0_ + 8]
            lq_tiles = torch.cat([lq_nhwc[:, idx[t][0]:idx[t][0] + k1, idx[t][1]:idx[t][1] + k2, :] for t in part], dim=0)
            cond_all[n0_:n0_ + len(part)].copy_(unet.condition_embed(lq_tiles))
        ids_all = torch.tensor([[original_size[0], original_size[1], i, j, target_size[0], target_size[1]] for i, j in idx],
                               dtype=torch.float32).to(dt).to(torch.float32)   # rounded through the working dtype
        # [S, T_mine, 2, ΣC]: the per-resnet time biases of every (step, tile of THIS rank) — 140 small-M launches over the 52 MB
        # stacked time_emb_proj at 4K, so a rank computes only its own tiles' rows (at N = 8 it was 21 of the stage's 54 ms)
        temb_table = unet.time_embedding_table(ts, te, ids_all[my_tiles[0]:my_tiles[-1] + 1].to(dev))
        temb_units = temb_table.reshape(temb_table.shape[0], 2 * len(my_tiles), -1)   # [S, U_mine, ΣC], local u = 2*(tile - first) + branch
        u_first = 2 * my_tiles[0]
        if self.world_size == 1:
            groups = [[u for t in g for u in (2 * t, 2 * t + 1)]
                      for g in plan_launch_groups(list(range(T)), range(1, max(1, self.tiles_per_launch) + 1))]
        else:
            groups = plan_unit_groups(my_units, 2 * max(1, self.tiles_per_launch))
        gbase = (k1, k2, bool(add_sample), lk, self.use_cuda_graph, wver)     # captured graphs hold the weights' addresses
        if self._graphs.get("key") != gbase:
            self._graphs = {"key": gbase}
        plans = []
        for g in groups:
            gt = sorted({u // 2 for u in g})
            slot = tuple(gt.index(u // 2) for u in g)
            branch = tuple(u % 2 for u in g)
            key = (slot, branch)
            if key not in self._graphs:
                self._graphs[key] = TileStepGraph(unet, kv, lk, slot, branch, k1, k2, dt, add_sample, self.use_cuda_graph)
            assert gt == list(range(gt[0], gt[-1] + 1)), "launch groups are contiguous tile ranges"
            c_lo = my_tiles.index(gt[0])
            plans.append((self._graphs[key], (gt[0], gt[-1] + 1), (c_lo, c_lo + len(gt)),
                          torch.tensor([u - u_first for u in g], device=dev, dtype=torch.long)))
        for key, gobj in self._graphs.items():
            if key != "key" and (kv_fresh or getattr(gobj, "_kv_id", None) is not kv):
                gobj.set_text_kv(kv)
                gobj._kv_id = kv
        count = None
        if self.world_size > 1:
            count = torch.zeros(shape, device=dev, dtype=torch.float32)
            ops.tile_accumulate(torch.ones((T, shape[1], k1, k2), device=dev, dtype=torch.float32), None, weights, origins, count)
            unit_tile = torch.tensor([u // 2 for u in my_units], device=dev, dtype=torch.int32)
            unit_addx = torch.tensor([1 - u % 2 for u in my_units], device=dev, dtype=torch.int32)
            # per-step coefficients of the linearised CFG + Euler update: (1-g)*dt for uncond units, g*dt for text units
            dts = torch.tensor([sch.step_scalars(i)[1] - sch.step_scalars(i)[0] for i in range(len(ts))], dtype=torch.float32)
            gsel = torch.tensor([guidance_scale if u % 2 else 1.0 - guidance_scale for u in my_units], dtype=torch.float32)
            unit_coef = (dts[:, None] * gsel[None, :]).to(dev).contiguous()       # [S, U_local]
        ev["enc"].record()

        # -- denoising loop (pipeline…:1324-1409)
        s0, _, div0 = sch.step_scalars(0)
        scaled = ops.scale_div(tiles, div0, out_dtype=dt)
        n_local = len(my_units)
        eps = torch.empty((n_local,) + tuple(tiles.shape[1:]), device=dev, dtype=dt)     # unit order (pairs when 1 GPU)
        canvas_lowp = None
        peer = None
        num = torch.empty(shape, device=dev, dtype=torch.float32) if self.world_size > 1 else None
        if self.world_size > 1 and os.environ.get("FD_P2P_EXCHANGE", "0") == "1":
            key = (tuple(shape), self.rank, self.world_size)
            if getattr(self, "_peer_key", None) != key:
                if getattr(self, "_peer", None) is not None:
                    self._peer.close()
                from .dist import PeerCanvas
                self._peer, self._peer_key, self._peer_step = PeerCanvas(shape, self.rank, self.world_size, self.process_group, dev), key, 0
            peer = self._peer
        for i in range(len(ts)):
            sigma, sigma_next, _ = sch.step_scalars(i)
            next_div = sch.step_scalars(i + 1)[2] if i + 1 < len(ts) else 1.0
            tval = ts[i]
            pos = 0
            for gobj, (t_lo, t_hi), (c_lo, c_hi), gu in plans:
                gobj(scaled[t_lo:t_hi], cond_all[c_lo:c_hi], temb_units[i].index_select(0, gu), eps[pos:pos + gobj.ns])
                pos += gobj.ns
            if self.world_size == 1:
                canvas, canvas_lowp, tiles, scaled = ops.cfg_euler_blend(eps.view((T, 2) + tuple(tiles.shape[1:])), tiles, weights,
                                                                         origins, h, w, guidance_scale, sigma, sigma_next, next_div)
            elif peer is not None:
                # opt-in: exchange + divide as one kernel over NVLink peer memory (dist.PeerCanvas, csrc/peer_exchange.cu)
                self._peer_step += 1
                num = peer.numerator(self._peer_step)
                num.zero_()
                ops.unit_accumulate(tiles, eps, weights, origins, unit_tile, unit_coef[i], unit_addx, num)
                canvas, canvas_lowp = peer.exchange_finalize(self._peer_step, count, dt)
                tiles = ops.tile_gather(canvas if sdt == torch.float32 else canvas_lowp, origins, k1, k2)
                scaled = ops.scale_div(tiles, next_div, out_dtype=dt)
            else:
                # three launches: this rank's numerator (overwrite, no zero-fill), the exchange, divide + cast + re-tile + scale
                ops.unit_numerator(tiles, eps, weights, origins, unit_tile, unit_coef[i], unit_addx, num)
                exchange_canvas(num, self.process_group)
                canvas, canvas_lowp, tiles, scaled = ops.canvas_finalize_retile(num, count, origins, k1, k2, next_div, dt, sdt)
            if callback_on_step_end is not None:
                callback_on_step_end(self, i, tval, {"latents": canvas if sdt == torch.float32 else canvas_lowp})
        latents = canvas_lowp
        ev["loop"].record()
        if output_type == "latent":
            image = latents
        else:
            # -- decode (pipeline…:1411-1451): latents / scaling_factor → VAE → u8
            z = latents / vae.config.scaling_factor
            image = (vae.decode(z, return_dict=False, dist=self._dist) if self._dist is not None else vae.decode(z, return_dict=False))[0]
            u8 = ops.image_postprocess(ops.nchw_to_nhwc(image.contiguous()), True)
            self._last_u8 = u8
            if output_type == "pil":
                host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
                host.copy_(u8, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                self._d2h_bytes = host.numel()
                image = [PIL.Image.fromarray(host.numpy())]
            elif output_type == "u8":
                image = [u8]
            else:
                image = ops.nchw_to_nhwc(image.contiguous())
        ev["dec"].record()
        self._events = ev
        if not return_dict:
            return (image,)
        return FaithDiffStableDiffusionXLPipelineOutput(images=image)

    def graph_kernel_launches(self):
        """Kernels of this library executed through CUDA-graph replays so far (fd_launch_count() only sees eager launches)."""
        return sum(g.replays * g.kernels_per_replay for k, g in self._graphs.items() if k != "key")

    def stage_times_ms(self):
        e = self._events
        torch.cuda.synchronize()
        return {"encode": e["t0"].elapsed_time(e["enc"]), "denoise": e["enc"].elapsed_time(e["loop"]),
                "decode": e["loop"].elapsed_time(e["dec"])}
