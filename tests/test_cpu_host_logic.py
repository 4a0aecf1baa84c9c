"""-m "not gpu": host-side logic of the product package (no kernel launches) and the C-ABI surface."""
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_loads_and_exports_every_declared_symbol():
    import ctypes
    from faithdiff_dev_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "faithdiff_b200.h")).read()
    declared = set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 28
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.lib.fd_version() == 100 and _lib.lib.fd_launch_count() == 0
    assert _lib.lib.fd_groupnorm_work_floats(2, 32, 320) > 0
    # the ctypes GemmDesc mirrors the C struct size (catches field drift)
    m = re.search(r"typedef struct fd_gemm_desc \{(.*?)\} fd_gemm_desc;", hdr, re.S)
    assert m and ctypes.sizeof(_lib.GemmDesc) == lib.fd_sizeof_gemm_desc() == 176


def test_cabi_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++ types, no torch types) and as C++."""
    import shutil, subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "faithdiff_b200.h"\nint main(void) { return fd_version == 0; }\n')
    inc = os.path.join(ROOT, "include")
    for cmd in (["gcc", "-std=c99", "-pedantic", "-Werror", "-fsyntax-only"], ["g++", "-std=c++17", "-Werror", "-fsyntax-only", "-x", "c++"]):
        r = subprocess.run(cmd + ["-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_no_cuda_no_fallback():
    """The product path must fail loudly without a GPU — never compute on the CPU."""
    import faithdiff_dev_b200 as fd
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    la = fd.LocalAttention((16, 16), 0.5)
    with pytest.raises(RuntimeError):
        la.grids(torch.zeros(1, 4, 32, 32))
    enc = fd.Encoder()
    with pytest.raises(RuntimeError):
        enc(torch.zeros(1, 3, 64, 64))
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import faithdiff_dev_b200._lib as L; import os; "
            "L.LIB_PATH = '/nonexistent.so'; \ntry:\n  L._load()\n  print('LOADED')\nexcept ImportError as e:\n  print('IMPORTERROR')") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "IMPORTERROR" in out


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "faithdiff_dev_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), fn


def test_tile_layout_matches_golden(golden_dir):
    from faithdiff_dev_b200.tiling import gaussian_weights, tile_origins
    man = json.load(open(os.path.join(golden_dir, "manifest.json")))
    for name, c in man["layout"].items():
        idx, k1, k2 = tile_origins(c["h"], c["w"], c["k"], c["k"], c["overlap"])
        assert [list(x) for x in idx] == c["idxes"] and (k1, k2) == (c["k1"], c["k2"]), name
    assert np.array_equal(gaussian_weights(128, 128), np.load(os.path.join(golden_dir, "gaussian_weights_128.npy")))
    assert np.array_equal(gaussian_weights(64, 64), np.load(os.path.join(golden_dir, "gaussian_weights_64.npy")))
    # empty / degenerate: one tile when the canvas is smaller than the kernel
    assert tile_origins(10, 12, 128, 128, 0.5) == ([(0, 0)], 10, 12)


def test_scheduler_scalars_match_golden(golden_dir):
    from faithdiff_dev_b200.scheduler import DDPMScheduler, EulerDiscreteScheduler
    man = json.load(open(os.path.join(golden_dir, "manifest.json")))
    s = EulerDiscreteScheduler()
    s.set_timesteps(20)
    assert s.timesteps.tolist() == man["euler20"]["timesteps"]
    assert [float(v) for v in s.sigmas] == man["euler20"]["sigmas"]
    assert float(s.init_noise_sigma) == man["euler20"]["init_noise_sigma"]
    sig, nxt, div = s.step_scalars(0)
    assert sig == man["euler20"]["sigmas"][0] and nxt == man["euler20"]["sigmas"][1]
    assert div == float((torch.tensor(sig) ** 2 + 1) ** 0.5)
    x0, n = torch.randn(1, 4, 8, 8), torch.randn(1, 4, 8, 8)
    d = DDPMScheduler()
    a = d.alphas_cumprod[999]
    assert torch.equal(d.add_noise(x0, n, torch.tensor([999])), a ** 0.5 * x0 + (1 - a) ** 0.5 * n)


def test_weight_packing():
    from faithdiff_dev_b200.blocks import Weights, interleave_geglu
    w = torch.arange(2 * 3 * 9, dtype=torch.float32).reshape(2, 3, 3, 3)
    W = Weights({"c.weight": w}, "cpu")
    p = W.conv3("c.weight").float()
    for o in range(2):
        for r in range(3):
            for s in range(3):
                for c in range(3):
                    assert p[o, (r * 3 + s) * 3 + c] == w[o, c, r, s]
    ff = torch.arange(8 * 2, dtype=torch.float32).reshape(8, 2)
    b = torch.arange(8, dtype=torch.float32)
    wi, bi = interleave_geglu(ff, b)
    assert torch.equal(wi[0::2], ff[:4]) and torch.equal(wi[1::2], ff[4:]) and torch.equal(bi, torch.tensor([0., 4, 1, 5, 2, 6, 3, 7]))


def test_pipeline_input_validation():
    import faithdiff_dev_b200 as fd

    class V:
        config = type("C", (), {"block_out_channels": (1, 2, 3, 4), "latent_channels": 4, "scaling_factor": 0.13025})()
    class U:
        config = type("C", (), {"sample_size": 128})()
    pipe = fd.FaithDiffStableDiffusionXLPipeline(V(), None, None, None, None, None, U(), fd.EulerDiscreteScheduler(), fd.DDPMScheduler())
    with pytest.raises(ValueError):
        pipe.check_inputs(None, "x", 100, 128, None, None, None, None, None)       # not divisible by 8
    with pytest.raises(ValueError):
        pipe.check_inputs(None, None, 128, 128, None, None, None, None, None)      # neither prompt nor embeds
    with pytest.raises(ValueError):
        pipe.check_inputs(None, "x", 128, 128, torch.zeros(1), None, None, None, None)
    with pytest.raises(ValueError):
        pipe.check_inputs(None, None, 128, 128, torch.zeros(1), None, None, None, None)   # embeds without pooled
    with pytest.raises(NotImplementedError):
        pipe.enable_model_cpu_offload()
    with pytest.raises(RuntimeError):
        pipe.to("cpu")
    assert pipe.vae_scale_factor == 8


def test_arch_key_layout_matches_oracle_and_known_counts():
    """The product's own checkpoint key layout equals the oracle's state dicts (which equal the reference's, see
    oracle/make_golden.py) and the public SDXL parameter totals."""
    import math
    from faithdiff_dev_b200 import arch
    from oracle import faithdiff_oracle as O
    with torch.device("meta"):
        u = O.FaithDiffUNet(**O.SDXL_UNET_CONFIG)
        u.load_additional_layers(channel=512)
        v = O.AutoencoderKL()
        ut = O.build_unet.__wrapped__ if hasattr(O.build_unet, "__wrapped__") else None
    shapes = arch.unet_param_shapes(arch.SDXL_UNET_CONFIG)
    ref = {k: tuple(p.shape) for k, p in u.state_dict().items()}
    assert shapes == ref
    assert sum(math.prod(s) for s in shapes.values()) == 2612197892
    vs = arch.vae_param_shapes()
    assert vs == {k: tuple(p.shape) for k, p in v.state_dict().items()}
    assert sum(math.prod(s) for s in vs.values()) == 83653863
    with torch.device("meta"):
        t = O.FaithDiffUNet(**O.TINY_UNET_CONFIG)
        t.load_additional_layers(channel=64, encoder_kwargs=O.TINY_ENCODER_CONFIG)
    assert arch.unet_param_shapes(O.TINY_UNET_CONFIG, lq_channels=64, encoder_channels=(32, 64, 64, 64)) == \
        {k: tuple(p.shape) for k, p in t.state_dict().items()}


def test_check_image_size_matches_oracle():
    from PIL import Image
    from faithdiff_dev_b200 import check_image_size
    from oracle import faithdiff_oracle as O
    for (w, h) in [(50, 37), (64, 64), (1, 1), (3840, 2160), (123, 8)]:
        img = Image.fromarray(np.random.default_rng(w).integers(0, 256, (h, w, 3), dtype=np.uint8))
        a, b = check_image_size(img), O.check_image_size(img)
        assert a[1:] == b[1:] and np.array_equal(np.asarray(a[0]), np.asarray(b[0]))
        assert a[3] % 8 == 0 and a[4] % 8 == 0
        # the tensor form (what `resize_lanczos` returns, padded without leaving the device) is the same operation
        t = check_image_size(torch.from_numpy(np.asarray(img).copy()))
        assert t[1:] == b[1:] and np.array_equal(t[0].numpy(), np.asarray(b[0]))


def test_quantize_8bit_state_dict():
    """utils/system.py:13-32 semantics: UNet tensors rounded through e4m3 (unscaled), the LQ Encoder untouched."""
    from faithdiff_dev_b200 import quantize_8bit_state_dict
    sd = {"conv_in.weight": torch.randn(8, 4, 3, 3) * 0.1, "denoise_encoder.conv_in.weight": torch.randn(8, 3, 3, 3)}
    q = quantize_8bit_state_dict(sd)
    assert torch.equal(q["denoise_encoder.conv_in.weight"], sd["denoise_encoder.conv_in.weight"])
    ref = sd["conv_in.weight"].to(torch.float8_e4m3fn).float()
    assert torch.equal(q["conv_in.weight"], ref) and not torch.equal(ref, sd["conv_in.weight"])


def test_text_encoder_host_logic():
    """encode_prompt host pieces that need no GPU: pooled-token row (transformers CLIPTextTransformer semantics) and the
    ModelOutput-style indexing the reference relies on (`prompt_embeds[0]`, `.hidden_states[-2]`, pipeline…:605-612)."""
    import torch
    from faithdiff_dev_b200.text_encoder import CLIP_L_CONFIG, OPENCLIP_BIGG_CONFIG, CLIPTextEncoder, TextEncoderOutput, eos_positions
    ids = torch.tensor([[49406, 320, 1125, 49407, 49407, 49407], [49406, 5, 6, 7, 8, 49407]])
    assert eos_positions(ids, 2).tolist() == [3, 5]                 # legacy config: argmax (first max)
    assert eos_positions(ids, 49407).tolist() == [3, 5]             # first occurrence of the eos id
    assert eos_positions(torch.tensor([[1, 9, 4, 9]]), 9).tolist() == [1]
    last, pooled, hs, te = object(), object(), (1, 2, 3), object()
    assert TextEncoderOutput(last, pooled, hs, None)[0] is last and TextEncoderOutput(last, pooled, hs, None)[1] is pooled
    assert TextEncoderOutput(last, pooled, hs, te)[0] is te and TextEncoderOutput(last, pooled, hs, te)[1] is last
    assert TextEncoderOutput(last, pooled, hs, te)[2] is hs
    assert CLIP_L_CONFIG["hidden_size"] == 64 * CLIP_L_CONFIG["num_attention_heads"]
    assert OPENCLIP_BIGG_CONFIG["hidden_size"] == 64 * OPENCLIP_BIGG_CONFIG["num_attention_heads"]
    import pytest
    with pytest.raises(ValueError):
        CLIPTextEncoder(hidden_size=640, num_attention_heads=8)     # 80-wide heads: not this kernel
    with pytest.raises(RuntimeError):
        CLIPTextEncoder()(ids)                                      # weights not loaded


def _np_resample(a, ow, oh):
    """The two integer passes of fd_resample_u8_* in numpy (checker only), fed by the product's tap tables."""
    from faithdiff_dev_b200.image_process import resample_taps
    H, W, _ = a.shape

    def clip8(v):
        return np.clip(v >> 22, 0, 255).astype(np.uint8)
    row0 = 0
    if ow != W:
        hb, hk = resample_taps(W, ow)
        rows = a
        if oh != H:
            vb, _ = resample_taps(H, oh)
            row0 = int(vb[0, 0])
            rows = a[row0:int(vb[-1, 0] + vb[-1, 1])]
        tmp = np.empty((rows.shape[0], ow, a.shape[2]), np.uint8)
        for xx in range(ow):
            x0, n = hb[xx]
            tmp[:, xx] = clip8((rows[:, x0:x0 + n].astype(np.int64) * hk[xx, :n].astype(np.int64)[None, :, None]).sum(1) + (1 << 21))
        a = tmp
    if oh != H:
        vb, vk = resample_taps(H, oh)
        out = np.empty((oh,) + a.shape[1:], np.uint8)
        for yy in range(oh):
            y0, n = vb[yy]
            y0 -= row0
            out[yy] = clip8((a[y0:y0 + n].astype(np.int64) * vk[yy, :n].astype(np.int64)[:, None, None]).sum(0) + (1 << 21))
        a = out
    return a


@pytest.mark.parametrize("w,h,ow,oh", [(37, 23, 74, 46), (64, 48, 128, 96), (50, 40, 200, 160), (100, 80, 50, 40), (33, 17, 40, 17),
                                       (31, 29, 31, 60), (120, 90, 37, 29), (16, 16, 129, 67), (240, 135, 480, 270)])
def test_lanczos_tap_tables_reproduce_pil(w, h, ow, oh):
    """The caller's `image.resize(..., Image.LANCZOS)` (test_wo_llava.py:81): the product's host-side tap tables pushed
    through the kernels' integer arithmetic are bit-identical to Pillow — up-, down- and one-axis scaling."""
    from PIL import Image
    a = np.random.default_rng(w * 1000 + oh).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(a).resize((ow, oh), Image.LANCZOS))
    assert np.array_equal(_np_resample(a, ow, oh), ref)


def test_lanczos_tap_table_properties():
    from faithdiff_dev_b200.image_process import resample_taps, resize_lanczos
    for n_in, n_out in [(1920, 3840), (1080, 2160), (512, 2048), (2048, 512), (7, 7)]:
        b, k = resample_taps(n_in, n_out)
        assert b.shape == (n_out, 2) and k.shape[0] == n_out and k.dtype == np.int32
        assert (b[:, 0] >= 0).all() and (b[:, 0] + b[:, 1] <= n_in).all() and (b[:, 1] <= k.shape[1]).all()
        assert (np.abs(k.sum(1) - (1 << 22)) <= k.shape[1]).all()              # taps are normalised before rounding
        assert (np.diff(b[:, 0]) >= 0).all()
    with pytest.raises(RuntimeError):
        resize_lanczos(np.zeros((8, 8, 3), np.uint8), (16, 16), device="cpu")   # no CPU path
    with pytest.raises(ValueError):
        resize_lanczos(np.zeros((8, 8, 3), np.float32), (16, 16))


def test_bench_tile_count_matches_the_tile_walk():
    """bench.count_tiles (used to scale the CPU sample) against the product's LocalAttention walk for every workload."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from faithdiff_dev_b200.tiling import tile_origins
    for name, want in (("4k", 28), ("8k", 112), ("1k", 1), ("batch32", 9), ("tiny", 1)):
        wl = bench.workload(name)
        h, w = wl["lq"][1] * wl["scale"] // 8, wl["lq"][0] * wl["scale"] // 8
        assert bench.count_tiles(wl) == len(tile_origins(h, w, 128, 128, 0.5)[0]) == want


def test_bsrnet_host_logic():
    """RRDBNet keeps the reference's state-dict key set (bsrnet_arch.py:38-43, :82-92) and argument checks; no GPU needed."""
    import faithdiff_dev_b200 as fd
    from oracle import bsrnet_oracle as B
    net = fd.RRDBNet(in_nc=3, out_nc=3, nf=64, nb=23, gc=32, sf=4)
    assert net.expected_keys() == sorted(B.param_shapes())
    assert fd.RRDBNet(nb=2, sf=2).expected_keys() == sorted(B.param_shapes(nb=2, sf=2))
    with pytest.raises(ValueError):
        fd.RRDBNet(sf=3)
    with pytest.raises(NotImplementedError):
        fd.RRDBNet(in_nc=1)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            net.load_state_dict({})                      # no CUDA device: fails loudly, no CPU path
        with pytest.raises(RuntimeError):
            net.deg_remove(torch.zeros(1, 3, 8, 8))
    # the tile plan of deg_remove is the reference's (checked bit-exactly through the golden outputs on the GPU); its geometry:
    plan = B.tile_plan(40, 530)
    assert [(p[2], p[3], p[6], p[7]) for p in plan] == [(0, 512, 0, 528), (512, 530, 496, 530)]


def test_fold_layernorm_host_preparation():
    """blocks.fold_layernorm (the host half of FD_LN_FOLD): with W' = bf16(W*gamma), colsum(W') and bias' = bias + W beta, the
    consumer epilogue's formula rstd*(x W'^T) - rstd*mean*colsum + bias' equals layer_norm(x) W^T + bias (fp32 emulation of the
    kernel's arithmetic on bf16 inputs, one-pass row sums as the producer epilogue accumulates them)."""
    from faithdiff_dev_b200.blocks import fold_layernorm, interleave_geglu
    torch.manual_seed(0)
    M, Cc, N = 64, 640, 256
    x = (torch.randn(M, Cc) * 1.7 + 0.8).to(torch.bfloat16)
    w = (torch.randn(N, Cc) / Cc ** 0.5).to(torch.bfloat16)
    b = torch.randn(N) * 0.1
    gamma, beta = 1 + 0.3 * torch.randn(Cc), 0.2 * torch.randn(Cc)
    for bias, ww in ((b, w), (None, w), interleave_geglu(w, b)[::-1]):
        wp, cs, bp = fold_layernorm(ww, bias, gamma, beta)
        assert wp.dtype == torch.bfloat16 and cs.dtype == bp.dtype == torch.float32 and cs.shape == bp.shape == (N,)
        xf = x.float()
        s1, s2 = xf.sum(1, keepdim=True), (xf * xf).sum(1, keepdim=True)
        mean = s1 / Cc
        rstd = torch.rsqrt((s2 / Cc - mean * mean).clamp_min(0) + 1e-5)
        out = rstd * (xf @ wp.float().t()) - rstd * mean * cs + bp
        ref = torch.nn.functional.layer_norm(xf, (Cc,), gamma, beta, 1e-5) @ ww.float().t() + (bias if bias is not None else 0)
        assert ((out - ref).norm() / ref.norm()).item() < 3e-3


def test_upsample_phase_weights_reproduce_conv_over_nearest_upsample():
    """blocks.upsample_phase_weights (host half of fd_gemm mode 3): the four 2x2-tap phase convs with pre-summed weights equal
    conv3x3(nearest_upsample_2x(x)) including the zero padding of the UPSAMPLED image (Upsample2D, …extension.py:661-684)."""
    import torch
    import torch.nn.functional as F
    from faithdiff_dev_b200.blocks import upsample_phase_weights
    torch.manual_seed(0)
    w, x = torch.randn(6, 5, 3, 3), torch.randn(2, 5, 7, 9)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w.to(torch.bfloat16).float(), padding=1)
    wp = upsample_phase_weights(w)
    assert len(wp) == 4 and all(p.shape == (6, 20) and p.dtype == torch.bfloat16 for p in wp)
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for a in range(2):
        for b in range(2):
            k = wp[a * 2 + b].float().reshape(6, 2, 2, 5).permute(0, 3, 1, 2)          # [O, I, 2, 2], taps (i-1+a+r', j-1+b+s')
            out[:, :, a::2, b::2] = F.conv2d(xp[:, :, a:a + 8, b:b + 10], k)
    assert ((out - ref).norm() / ref.norm()).item() < 5e-3                                 # bf16 rounding of the summed taps only


def test_tiled_stage_assignment_is_balanced_and_deterministic():
    """vae.assign_tiles: owner rank of each tile of a tiled one-time stage by area (edge tiles are smaller).  4K LQ encoder: 3 x 5
    tiles of 1024 px at stride 768 over 2160 x 3840 — twelve full tiles and three 624-row ones."""
    from faithdiff_dev_b200.vae import assign_tiles
    areas = []
    for i in range(0, 2160, 768):
        for j in range(0, 3840, 768):
            areas.append(min(1024, 2160 - i) * min(1024, 3840 - j))
    assert len(areas) == 15
    for world in (2, 4, 8):
        owner = assign_tiles(areas, world)
        assert owner == assign_tiles(list(areas), world) and set(owner) == set(range(world))
        load = [sum(a for a, o in zip(areas, owner) if o == r) for r in range(world)]
        ideal = sum(areas) / world
        assert max(load) <= ideal + max(areas)                  # greedy largest-first: within one tile of the mean
        rr = [sum(a for n, a in enumerate(areas) if n % world == r) for r in range(world)]
        assert max(load) <= max(rr)                             # never worse than round-robin


def test_conv_patch_rule_mirrors_the_kernel():
    """ops._conv_tiles mirrors gemm_tcgen05.cu pick_patch (the 128-pixel patch w x 128/w that wastes least, widest first): it sizes
    the GroupNorm-statistics buffers, and the C side re-checks the count (colstats_boxes) on every call."""
    from faithdiff_dev_b200.ops import _conv_tiles
    assert _conv_tiles(128, 128) == 128 and _conv_tiles(64, 64) == 32 and _conv_tiles(32, 32) == 8
    assert _conv_tiles(8, 8) == 1                               # 16 x 8 patch covers it
    assert _conv_tiles(24, 40) == 9                             # 16 x 8 patches: 3 columns x 3 rows waste 1.2x (8 x 16: 1.33x, 64 x 2: 1.6x)
    assert _conv_tiles(18, 14) == 3 and _conv_tiles(135, 240) == 255
