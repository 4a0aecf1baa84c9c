"""B200-native drop-ins for the FaithDiff LQ `Encoder` (FaithDiff/models/unet_2d_condition_vae_extension.py:39-196)
and for the `AutoencoderKL` surface the pipeline uses (encode().latent_dist.sample(), decode(), tiling;
pipeline_FaithDiff_tlc.py:1272,1438).  NCHW tensors at the boundary, NHWC bf16 + sm_100a kernels inside."""
import torch

from . import ops
from .blocks import BF16, Downsample, ResnetBlock, Upsample, VaeAttention, Weights


class _Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _load_down_path(W, n_blocks, groups):
    blocks = []
    for i in range(n_blocks):
        rs = []
        j = 0
        while W.has("down_blocks.%d.resnets.%d.norm1.weight" % (i, j)):
            rs.append(ResnetBlock(W.sub("down_blocks.%d.resnets.%d." % (i, j)), groups, 1e-6))
            j += 1
        down = Downsample(W.sub("down_blocks.%d.downsamplers.0." % i), 0) if W.has(
            "down_blocks.%d.downsamplers.0.conv.weight" % i) else None
        blocks.append((rs, down))
    mid = (ResnetBlock(W.sub("mid_block.resnets.0."), groups, 1e-6),
           VaeAttention(W.sub("mid_block.attentions.0."), groups, 1e-6),
           ResnetBlock(W.sub("mid_block.resnets.1."), groups, 1e-6))
    return blocks, mid


def _run_down_path(x, conv_in, blocks, mid):
    x = ops.conv3x3_small_in(x, *conv_in, nhwc_in=True)
    for rs, down in blocks:
        for r in rs:
            x = r(x)
        if down is not None:
            x = down(x)
    x = mid[0](x)
    x = mid[1](x)
    return mid[2](x)


def _count_blocks(sd, prefix):
    n = 0
    while (prefix + "%d.resnets.0.norm1.weight" % n) in sd:
        n += 1
    return n


class DistCtx:
    """Multi-GPU context of the one-time tiled stages.  It belongs to a pipeline (`pipeline.set_distributed`) and is handed to
    every encode / decode call (`dist=`), so two pipelines of one process (e.g. a 1-rank and an N-rank one over the same VAE)
    do not share it.  Tiles are independent up to the seam blend: every rank computes a FLOP-balanced subset of the tiles, ONE
    all-gather moves them, the (cheap) seam blend is replicated."""
    __slots__ = ("rank", "world", "group")

    def __init__(self, rank=0, world=1, group=None):
        self.rank, self.world, self.group = rank, world, group


def assign_tiles(areas, world):
    """Owner rank of each tile, balanced by tile area (edge tiles are smaller, so round-robin is not balanced): largest tile
    first onto the least-loaded rank; ties broken by tile index / rank, hence identical on every rank."""
    load = [0] * world
    owner = [0] * len(areas)
    for idx in sorted(range(len(areas)), key=lambda i: (-areas[i], i)):
        r = min(range(world), key=lambda q: (load[q], q))
        owner[idx] = r
        load[r] += areas[idx]
    return owner


def _down8(n):
    for _ in range(3):
        n = (n - 2) // 2 + 1          # F.pad(0,1,0,1) + conv3x3 stride 2 pad 0
    return n


class StageGraphs:
    """CUDA graphs for the one-time tiled stages: one graph per (network, batch shape), static input / output buffers, all graphs
    in ONE shared memory pool (they never replay concurrently, so their intermediates alias).  At N = 8 a rank runs only two LQ-
    encoder tiles but still ~1000 launches for them: eager, the stage was bound by the host's launch rate (encode 60 ms against
    27 ms of GPU work); replayed, it is one launch.  FD_STAGE_GRAPHS=0 keeps the eager path."""

    def __init__(self):
        import os
        self.enabled = os.environ.get("FD_STAGE_GRAPHS", "1") == "1"
        self.pool = None

    def __call__(self, fn, x):
        owner = getattr(fn, "__self__", None)        # graphs live ON the network object: they die with its weights
        if not self.enabled or owner is None or torch.cuda.is_current_stream_capturing():
            return fn(x)
        cache = owner.__dict__.setdefault("_stage_graphs", {})
        key = (fn.__name__, tuple(x.shape), x.dtype)
        ent = cache.get(key)
        if ent is None:
            xin = x.clone()
            fn(xin)                                   # eager warm-up: kernel attributes, persistent sync words, allocator pools
            torch.cuda.current_stream().synchronize()
            if self.pool is None:
                self.pool = torch.cuda.graph_pool_handle()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.pool):
                out = fn(xin)
            ent = cache[key] = (g, xin, out)
        g, xin, out = ent
        xin.copy_(x)
        g.replay()
        return out.clone()                            # the static output is overwritten by the next replay


STAGE_GRAPHS = StageGraphs()


# tiles of one shape are pushed through the network together (batch axis): the tiled stages are hundreds of small
# launches per tile — at 45 VAE tiles per 4K image the per-launch latency, not the arithmetic, was their cost
TILE_BATCH = 8


def _tiled(x, fn, tile_in, stride, blend_extent, row_limit, out_shape=None, dist=None):
    """Shared structure of Encoder.tiled_encode (…extension.py:164-190) and AutoencoderKL.tiled_encode/decode:
    row-major tiles, in-place seam cross-fade with the tile above then the tile to the left, crop, concat.
    Tiles are independent up to the seam blend, so equal-shaped tiles run as one batch."""
    world, rank = (dist.world, dist.rank) if dist is not None else (1, 0)
    shard = world > 1 and out_shape is not None
    bx = x.shape[0]
    layout, srcs = [], []
    for i in range(0, x.shape[2], stride):
        row = []
        for j in range(0, x.shape[3], stride):
            row.append(len(srcs))
            srcs.append(x[:, :, i:i + tile_in, j:j + tile_in])
        layout.append(row)
    n = len(srcs)
    owner = assign_tiles([s.shape[2] * s.shape[3] for s in srcs], world) if shard else [rank] * n
    mine, tiles = {}, {}
    for idx, src in enumerate(srcs):
        if owner[idx] == rank:
            mine.setdefault(tuple(src.shape), []).append((idx, src))
    for items in mine.values():
        for k in range(0, len(items), TILE_BATCH):
            chunk = items[k:k + TILE_BATCH]
            out = STAGE_GRAPHS(fn, torch.cat([src for _, src in chunk], dim=0) if len(chunk) > 1 else chunk[0][1].contiguous())
            for t, (idx, _) in enumerate(chunk):
                tiles[idx] = out[t * bx:(t + 1) * bx]
    if shard:
        # ONE all-gather after all local tiles are enqueued (the ranks compute concurrently): every rank packs its tiles, in
        # tile order, into a flat buffer padded to the largest rank's size; the other ranks' tiles are views into the result
        shapes = [tuple(out_shape(s.shape)) for s in srcs]
        numel = [sh[0] * sh[1] * sh[2] * sh[3] for sh in shapes]
        per_rank = [sum(numel[i] for i in range(n) if owner[i] == r) for r in range(world)]
        cap = max(per_rank)
        send = torch.empty(cap, device=x.device, dtype=x.dtype)
        pos = 0
        for idx in range(n):
            if owner[idx] == rank:
                send[pos:pos + numel[idx]].copy_(tiles[idx].reshape(-1))
                pos += numel[idx]
        recv = torch.empty(world * cap, device=x.device, dtype=x.dtype)
        torch.distributed.all_gather_into_tensor(recv, send, group=dist.group)
        pos = [r * cap for r in range(world)]
        for idx in range(n):
            r = owner[idx]
            tiles[idx] = recv[pos[r]:pos[r] + numel[idx]].view(shapes[idx])
            pos[r] += numel[idx]
    out_rows = []
    for i, row in enumerate(layout):
        out = []
        for j, idx in enumerate(row):
            tile = tiles[idx]
            if i > 0:
                ops.seam_blend(tiles[layout[i - 1][j]], tile, blend_extent, True)
            if j > 0:
                ops.seam_blend(tiles[row[j - 1]], tile, blend_extent, False)
            out.append(tile[:, :, :row_limit, :row_limit])
        out_rows.append(torch.cat(out, dim=3))
    return torch.cat(out_rows, dim=2)


class Encoder:
    """FaithDiff LQ feature extractor = SDXL-VAE encoder body without conv_norm_out / conv_out."""

    def __init__(self, **unused):
        self.tile_sample_min_size = 1024
        self.tile_latent_min_size = int(self.tile_sample_min_size / 8)
        self.tile_overlap_factor = 0.25
        self.use_tiling = False
        self.dtype = BF16

    def load_state_dict(self, sd, strict=True, device="cuda"):
        self.__dict__.pop("_stage_graphs", None)      # captured graphs hold the old weights' addresses
        W = Weights(sd, torch.device(device))
        self.device = torch.device(device)
        self.conv_in = (W.conv3("conv_in.weight"), W.f32("conv_in.bias"))
        groups = 32
        self.blocks, self.mid = _load_down_path(W, _count_blocks(sd, "down_blocks."), groups)
        return self

    def enable_tiling(self):
        self.use_tiling = True

    def disable_tiling(self):          # missing in the reference (latent bug, pipeline…:913); provided here
        self.use_tiling = False

    def to(self, *a, **k):
        return self

    def encode_nhwc(self, x_nhwc):
        return _run_down_path(x_nhwc, self.conv_in, self.blocks, self.mid)

    def encode(self, sample):
        """NCHW in → NCHW out (1,512,H/8,W/8)."""
        return ops.nhwc_to_nchw(self.encode_nhwc(ops.nchw_to_nhwc(sample.contiguous())), sample.dtype)

    def tiled_encode(self, x, dist=None):
        stride = int(self.tile_sample_min_size * (1 - self.tile_overlap_factor))
        blend_extent = int(self.tile_latent_min_size * self.tile_overlap_factor)
        row_limit = self.tile_latent_min_size - blend_extent
        c_out = self.mid[2].cout
        return _tiled(x, self.encode, self.tile_sample_min_size, stride, blend_extent, row_limit,
                      out_shape=lambda s: (s[0], c_out, _down8(s[2]), _down8(s[3])), dist=dist)

    def forward(self, sample, dist=None):
        """`dist` (a DistCtx, multi-GPU pipelines only) is an addition to the reference signature."""
        if not sample.is_cuda:
            raise RuntimeError("faithdiff_dev_b200 has no CPU path: input must be a CUDA tensor")
        if self.use_tiling and (sample.shape[-1] > self.tile_latent_min_size or sample.shape[-2] > self.tile_latent_min_size):
            return self.tiled_encode(sample, dist)
        return self.encode(sample)

    __call__ = forward


class DiagonalGaussianDistribution:
    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator=None):
        # torch RNG on purpose: the reference draws this from torch's (global) generator (pipeline…:1272)
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class _EncOut:
    def __init__(self, d):
        self.latent_dist = d


class AutoencoderKL:
    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, norm_num_groups=32, sample_size=512, scaling_factor=0.13025, force_upcast=False,
                 latents_mean=None, latents_std=None, **unused):
        self.config = _Cfg(in_channels=in_channels, out_channels=out_channels, block_out_channels=tuple(block_out_channels),
                           layers_per_block=layers_per_block, latent_channels=latent_channels,
                           norm_num_groups=norm_num_groups, sample_size=sample_size, scaling_factor=scaling_factor,
                           force_upcast=force_upcast, latents_mean=latents_mean, latents_std=latents_std)
        self.use_tiling = False
        self.tile_sample_min_size = sample_size
        self.tile_latent_min_size = int(sample_size / (2 ** (len(block_out_channels) - 1)))
        self.tile_overlap_factor = 0.25
        self.dtype = BF16

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def enable_tiling(self, use_tiling=True):
        self.use_tiling = use_tiling

    def disable_tiling(self):
        self.use_tiling = False

    def load_state_dict(self, sd, strict=True, device="cuda"):
        self.__dict__.pop("_stage_graphs", None)      # captured graphs hold the old weights' addresses
        dev = torch.device(device)
        self.device = dev
        g = self.config.norm_num_groups
        W = Weights(sd, dev, "encoder.")
        self.e_conv_in = (W.conv3("conv_in.weight"), W.f32("conv_in.bias"))
        self.e_blocks, self.e_mid = _load_down_path(W, _count_blocks(sd, "encoder.down_blocks."), g)
        self.e_norm = (W.f32("conv_norm_out.weight"), W.f32("conv_norm_out.bias"))
        self.e_conv_out = (W.conv3("conv_out.weight"), W.f32("conv_out.bias"))
        D = Weights(sd, dev, "decoder.")
        self.d_conv_in = (D.conv3("conv_in.weight"), D.f32("conv_in.bias"))
        self.d_mid = (ResnetBlock(D.sub("mid_block.resnets.0."), g, 1e-6),
                      VaeAttention(D.sub("mid_block.attentions.0."), g, 1e-6),
                      ResnetBlock(D.sub("mid_block.resnets.1."), g, 1e-6))
        self.d_blocks = []
        i = 0
        while D.has("up_blocks.%d.resnets.0.norm1.weight" % i):
            rs, j = [], 0
            while D.has("up_blocks.%d.resnets.%d.norm1.weight" % (i, j)):
                rs.append(ResnetBlock(D.sub("up_blocks.%d.resnets.%d." % (i, j)), g, 1e-6))
                j += 1
            up = Upsample(D.sub("up_blocks.%d.upsamplers.0." % i)) if D.has("up_blocks.%d.upsamplers.0.conv.weight" % i) else None
            self.d_blocks.append((rs, up))
            i += 1
        self.d_norm = (D.f32("conv_norm_out.weight"), D.f32("conv_norm_out.bias"))
        self.d_conv_out = (D.conv3("conv_out.weight"), D.f32("conv_out.bias"))
        Q = Weights(sd, dev)
        self.quant = (Q.f32("quant_conv.weight").reshape(Q.raw("quant_conv.weight").shape[0], -1).contiguous(), Q.f32("quant_conv.bias"))
        self.post_quant = (Q.f32("post_quant_conv.weight").reshape(Q.raw("post_quant_conv.weight").shape[0], -1).contiguous(),
                           Q.f32("post_quant_conv.bias"))
        return self

    # ------------------------------------------------------------------ encode
    def _encode_moments(self, x):
        """NCHW image → NCHW moments (quant_conv(encoder(x)))."""
        h = _run_down_path(ops.nchw_to_nhwc(x.contiguous()), self.e_conv_in, self.e_blocks, self.e_mid)
        n = ops.groupnorm(h, *self.e_norm, self.config.norm_num_groups, 1e-6, True)
        m = ops.conv3x3_small_out(n, *self.e_conv_out, out_dtype=x.dtype, nchw_out=True)
        return ops.conv1x1_small(m, *self.quant)

    def encode(self, x, return_dict=True, dist=None):
        if not x.is_cuda:
            raise RuntimeError("faithdiff_dev_b200 has no CPU path: input must be a CUDA tensor")
        if self.use_tiling and (x.shape[-1] > self.tile_sample_min_size or x.shape[-2] > self.tile_sample_min_size):
            stride = int(self.tile_sample_min_size * (1 - self.tile_overlap_factor))
            blend_extent = int(self.tile_latent_min_size * self.tile_overlap_factor)
            row_limit = self.tile_latent_min_size - blend_extent
            moments = _tiled(x, self._encode_moments, self.tile_sample_min_size, stride, blend_extent, row_limit,
                             out_shape=lambda s: (s[0], 2 * self.config.latent_channels, _down8(s[2]), _down8(s[3])), dist=dist)
        else:
            moments = self._encode_moments(x)
        return _EncOut(DiagonalGaussianDistribution(moments))

    # ------------------------------------------------------------------ decode
    def _decode_tile(self, z, nchw_out=True, in_scale=1.0):
        z = ops.conv1x1_small(z.contiguous(), *self.post_quant, in_scale=in_scale)
        x = ops.conv3x3_small_in(z, *self.d_conv_in)
        x = self.d_mid[0](x)
        x = self.d_mid[1](x)
        x = self.d_mid[2](x)
        for rs, up in self.d_blocks:
            for r in rs:
                x = r(x)
            if up is not None:
                x = up(x)
        n = ops.groupnorm(x, *self.d_norm, self.config.norm_num_groups, 1e-6, True)
        return ops.conv3x3_small_out(n, *self.d_conv_out, out_dtype=z.dtype, nchw_out=nchw_out)

    def decode(self, z, return_dict=True, generator=None, dist=None):
        if not z.is_cuda:
            raise RuntimeError("faithdiff_dev_b200 has no CPU path: input must be a CUDA tensor")
        if self.use_tiling and (z.shape[-1] > self.tile_latent_min_size or z.shape[-2] > self.tile_latent_min_size):
            stride = int(self.tile_latent_min_size * (1 - self.tile_overlap_factor))
            blend_extent = int(self.tile_sample_min_size * self.tile_overlap_factor)
            row_limit = self.tile_sample_min_size - blend_extent
            up = 2 ** (len(self.config.block_out_channels) - 1)
            dec = _tiled(z, self._decode_tile, self.tile_latent_min_size, stride, blend_extent, row_limit,
                         out_shape=lambda s: (s[0], self.config.out_channels, s[2] * up, s[3] * up), dist=dist)
        else:
            dec = self._decode_tile(z)
        if not return_dict:
            return (dec,)
        return _Cfg(sample=dec)
