"""Thin tensor-level wrappers over the C-ABI (include/faithdiff_b200.h).  torch supplies device memory and the
stream; every call lands in a hand-written sm_100a kernel of libfaithdiff_b200.so.  No fallbacks."""
import ctypes as C

import torch

from ._lib import GemmDesc, GemmLnDesc, check, lib

BF16 = torch.bfloat16
_DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _row_view(t):
    """2-D view requirements: unit inner stride; returns (ptr, cols, pitch)."""
    assert t.dim() == 2 and t.stride(1) == 1 and t.dtype == BF16, (t.shape, t.stride(), t.dtype)
    return t.data_ptr(), t.shape[1], t.stride(0)


class RowStats:
    """Per-row (sum, sum of squares) slots a producer GEMM leaves for the LayerNorm-folded consumer (fd_gemm_ln)."""
    __slots__ = ("slots", "n")

    def __init__(self, slots, n):
        self.slots, self.n = slots, n          # fp32 [M, ld, 2] tensor; number of valid slots per row


class ColStats:
    """Per-(32-row box, channel pair) (sum, sum of squares) a producer conv / GEMM leaves for the GroupNorm that consumes its
    output (fd_gemm_ln colstats_out -> fd_groupnorm_nhwc_from_stats).  Travels as the `_fd_colstats` attribute of the output
    tensor object; any op that makes a new tensor drops it and the consumer falls back to the two-pass GroupNorm."""
    __slots__ = ("buf", "ld", "boxes_per_sample", "segments", "seg_stride")

    def __init__(self, buf, ld, boxes_per_sample, segments=1, seg_stride=0):
        self.buf, self.ld, self.boxes_per_sample, self.segments, self.seg_stride = buf, ld, boxes_per_sample, segments, seg_stride

    def select_samples(self, index, n_src):
        """Statistics of x.index_select(0, index) for x of n_src samples (single segment)."""
        assert self.segments == 1
        b = self.buf[:n_src * self.boxes_per_sample].view(n_src, self.boxes_per_sample, self.ld, 2).index_select(0, index)
        return ColStats(b.reshape(-1, self.ld, 2).contiguous(), self.ld, self.boxes_per_sample)


GN_STATS = __import__("os").environ.get("FD_GN_STATS", "1") == "1"


def tag_stats(t, st):
    if st is not None:
        t._fd_colstats = st
    return t


def _conv_tiles(H, W):
    """Tiles per image of the implicit-GEMM conv (gemm_tcgen05.cu pick_patch: the 128-pixel patch w x 128/w that wastes least)."""
    best, tiles = None, 0
    w = 128
    while w >= 8:
        h = 128 // w
        tx, ty = (W + w - 1) // w, (H + h - 1) // h
        waste = (tx * w) * (ty * h) / float(W * H)
        if best is None or waste < best - 1e-9:
            best, tiles = waste, tx * ty
        w >>= 1
    return tiles


def gemm(a, w, a2=None, bias=None, rowbias=None, rows_per_batch=0, residual=None, act=0, out=None, out_fp32=False,
         alpha=1.0, bias_along_m=False, block_n=0, ln_in=None, ln_colsum=None, ln_eps=1e-5, stats_out=False,
         colstats_rows_per_sample=0):
    """out[M,N] = epilogue(alpha * [a | a2] @ w^T).  a/a2: bf16 [M,K*] (unit inner stride); w: bf16 [N,K] contiguous.
    LayerNorm folding (opt-in): `stats_out=True` returns (out, RowStats of out); `ln_in=RowStats of a` with
    `ln_colsum` = row sums of w (= W*gamma) computes LayerNorm(a) W^T + bias' without the LayerNorm pass."""
    d = GemmDesc()
    d.a[0], d.a_cols[0], d.a_ld[0] = _row_view(a)
    if a2 is not None:
        d.a[1], d.a_cols[1], d.a_ld[1] = _row_view(a2)
    M, N = a.shape[0], w.shape[0]
    assert w.is_contiguous() and w.dtype == BF16 and w.shape[1] == a.shape[1] + (a2.shape[1] if a2 is not None else 0)
    n_out = N // 2 if act == 2 else N
    if out is None:
        out = torch.empty((M, n_out), device=a.device, dtype=torch.float32 if out_fp32 else BF16)
    assert out.stride(1) == 1 and out.shape[0] == M and out.shape[1] == n_out
    d.b, d.M, d.N, d.mode = w.data_ptr(), M, N, 0
    d.act, d.bias_along_m, d.out_fp32 = act, int(bias_along_m), int(out.dtype == torch.float32)
    if bias is not None:
        assert bias.dtype == torch.float32
        d.bias = bias.data_ptr()
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.stride(1) == 1
        d.rowbias, d.rows_per_batch, d.ld_rowbias = rowbias.data_ptr(), rows_per_batch, rowbias.stride(0)
    if residual is not None:
        assert residual.dtype == BF16 and residual.stride(1) == 1
        d.residual, d.ld_res = residual.data_ptr(), residual.stride(0)
    d.out, d.ld_out, d.alpha, d.block_n = out.data_ptr(), out.stride(0), alpha, block_n
    if (GN_STATS and colstats_rows_per_sample and colstats_rows_per_sample % 128 == 0 and M % colstats_rows_per_sample == 0 and act == 0
            and out.dtype == BF16 and not bias_along_m and ln_in is None and not stats_out and N % 16 == 0):
        # GroupNorm statistics for the consumer of `out`: per (32-row box, column pair), tiles never straddle a sample
        lnd = GemmLnDesc()
        buf = torch.empty(((M + 127) // 128 * 4, N // 2, 2), device=a.device, dtype=torch.float32)
        lnd.colstats_out, lnd.colstats_ld = buf.data_ptr(), N // 2
        check(lib.fd_gemm_ln(C.byref(d), C.byref(lnd), _stream()), "fd_gemm_ln(colstats)")
        assert lnd.colstats_boxes == buf.shape[0]
        return tag_stats(out, ColStats(buf, N // 2, colstats_rows_per_sample // 32))
    if ln_in is None and not stats_out:
        check(lib.fd_gemm(C.byref(d), _stream()), "fd_gemm")
        return out
    ln = GemmLnDesc()
    if ln_in is not None:
        assert ln_colsum is not None and ln_colsum.dtype == torch.float32 and ln_colsum.numel() == N and a2 is None
        assert ln_in.slots.shape[0] == M and ln_in.slots.dtype == torch.float32 and ln_in.slots.is_contiguous()
        ln.stats_in, ln.stats_in_ld, ln.stats_in_n = ln_in.slots.data_ptr(), ln_in.slots.shape[1], ln_in.n
        ln.colsum, ln.inv_c, ln.eps = ln_colsum.data_ptr(), 1.0 / a.shape[1], ln_eps
    slots = None
    if stats_out:
        # slots per row = 2 per N tile; the tile choice is a pure function of the shape, so the count seen on the first call
        # of a shape sizes the buffer of the later ones (2 * ceil(N / 32) always suffices)
        key = (M, N, a.shape[1])
        slots = torch.empty((M, _LN_SLOTS.get(key, 2 * ((N + 31) // 32)), 2), device=a.device, dtype=torch.float32)
        ln.stats_out, ln.stats_out_ld = slots.data_ptr(), slots.shape[1]
    check(lib.fd_gemm_ln(C.byref(d), C.byref(ln), _stream()), "fd_gemm_ln")
    if stats_out:
        _LN_SLOTS[key] = ln.stats_out_n
        return out, RowStats(slots, ln.stats_out_n)
    return out


_LN_SLOTS = {}


def conv3x3(x, w, x2=None, bias=None, rowbias=None, residual=None, act=0, stride=1, pad_lo=1, out=None, block_n=0, alpha=1.0,
            colstats=False):
    """3x3 conv on NHWC bf16 x [B,H,W,C] (optionally channel-concatenated with x2); w packed [N, 9*(C+C2)].
    colstats: also leave GroupNorm statistics of the output for its consumer (ColStats, attribute of the returned tensor)."""
    assert x.dim() == 4 and x.dtype == BF16 and x.stride(3) == 1
    B, H, W, C0 = x.shape
    assert x.stride(1) == W * x.stride(2) and x.stride(0) == H * x.stride(1), "NHWC images must be dense in B,H,W"
    d = GemmDesc()
    d.a[0], d.a_cols[0], d.a_ld[0] = x.data_ptr(), C0, x.stride(2)
    ct = C0
    if x2 is not None:
        assert x2.shape[:3] == x.shape[:3] and x2.stride(3) == 1
        d.a[1], d.a_cols[1], d.a_ld[1] = x2.data_ptr(), x2.shape[3], x2.stride(2)
        ct += x2.shape[3]
    N = w.shape[0]
    assert w.is_contiguous() and w.shape[1] == 9 * ct, (w.shape, ct)
    if stride == 1:
        Ho, Wo = H, W
    else:
        Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if pad_lo else ((H - 2) // 2 + 1, (W - 2) // 2 + 1)
    if out is None:
        out = torch.empty((B, Ho, Wo, N), device=x.device, dtype=BF16)
    assert out.stride(3) == 1 and out.stride(1) == Wo * out.stride(2) and out.stride(0) == Ho * out.stride(1)
    d.b, d.N, d.mode, d.batch, d.in_h, d.in_w, d.pad_lo = w.data_ptr(), N, 1 if stride == 1 else 2, B, H, W, pad_lo
    d.act = act
    if bias is not None:
        d.bias = bias.data_ptr()
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.stride(1) == 1 and rowbias.shape[0] == B
        d.rowbias, d.ld_rowbias = rowbias.data_ptr(), rowbias.stride(0)
    if residual is not None:
        assert residual.dtype == BF16 and residual.stride(3) == 1
        d.residual, d.ld_res = residual.data_ptr(), residual.stride(2)
    d.out, d.ld_out, d.alpha, d.block_n = out.data_ptr(), out.stride(2), alpha, block_n
    if colstats and GN_STATS and act == 0 and N % 16 == 0:
        lnd = GemmLnDesc()
        bps = 4 * _conv_tiles(Ho, Wo)
        buf = torch.empty((B * bps, N // 2, 2), device=x.device, dtype=torch.float32)
        lnd.colstats_out, lnd.colstats_ld = buf.data_ptr(), N // 2
        check(lib.fd_gemm_ln(C.byref(d), C.byref(lnd), _stream()), "fd_gemm_ln(conv, colstats)")
        assert lnd.colstats_boxes == buf.shape[0] and lnd.colstats_boxes_per_image == bps, (lnd.colstats_boxes, buf.shape, bps)
        return tag_stats(out, ColStats(buf, N // 2, bps))
    check(lib.fd_gemm(C.byref(d), _stream()), "fd_gemm(conv)")
    return out


def conv3x3_upsampled2x(x, w_phases, bias=None, out=None, colstats=False):
    """conv3x3(nearest_upsample_2x(x)) without the upsampled tensor: four launches, one per output phase (a, b), each a 2x2-tap
    conv over x [B,H,W,C] with the phase's pre-summed weights w_phases[a*2+b] [N, 4*C] (blocks.upsample_phase_weights) writing
    every second pixel of every second row of out [B,2H,2W,N]."""
    assert x.dim() == 4 and x.dtype == BF16 and x.stride(3) == 1
    B, H, W, C0 = x.shape
    assert x.stride(1) == W * x.stride(2) and x.stride(0) == H * x.stride(1), "NHWC images must be dense in B,H,W"
    N = w_phases[0].shape[0]
    if out is None:
        out = torch.empty((B, 2 * H, 2 * W, N), device=x.device, dtype=BF16)
    assert out.is_contiguous() and out.shape == (B, 2 * H, 2 * W, N)
    st = None
    if colstats and GN_STATS and N % 16 == 0:
        bps = 4 * _conv_tiles(H, W)                    # per phase: the phase's output grid is the input grid
        st = ColStats(torch.empty((4 * B * bps, N // 2, 2), device=x.device, dtype=torch.float32), N // 2, bps, 4, B * bps)
    for ph in range(4):
        w = w_phases[ph]
        assert w.is_contiguous() and w.shape == (N, 4 * C0)
        d = GemmDesc()
        d.a[0], d.a_cols[0], d.a_ld[0] = x.data_ptr(), C0, x.stride(2)
        d.b, d.N, d.mode, d.batch, d.in_h, d.in_w, d.pad_lo = w.data_ptr(), N, 3, B, H, W, ph
        if bias is not None:
            d.bias = bias.data_ptr()
        d.out, d.ld_out, d.alpha = out.data_ptr(), out.stride(2), 1.0
        if st is not None:
            lnd = GemmLnDesc()
            lnd.colstats_out, lnd.colstats_ld = st.buf[ph * st.seg_stride:].data_ptr(), N // 2
            check(lib.fd_gemm_ln(C.byref(d), C.byref(lnd), _stream()), "fd_gemm_ln(upsampled conv, colstats)")
            assert lnd.colstats_boxes == st.seg_stride
        else:
            check(lib.fd_gemm(C.byref(d), _stream()), "fd_gemm(upsampled conv)")
    return tag_stats(out, st)


def attention(q, k, v, batch, heads, lq, lk, scale, out=None, causal=False):
    """q [batch*lq, >=heads*64] / k, v [batch*lk, ...] bf16 row views (unit inner stride) -> out [batch*lq, heads*64]."""
    if out is None:
        out = torch.empty((batch * lq, heads * 64), device=q.device, dtype=BF16)
    fn = lib.fd_attention_d64_causal if causal else lib.fd_attention_d64
    check(fn(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0), batch,
                               heads, lq, lk, scale, _stream()), "fd_attention_d64")
    return out


def attention_d512(q, k, v, batch, l, scale, out=None):
    """Single head, head_dim 512 (VAE / LQ-encoder mid block): q, k, v bf16 row views [batch*l, 512] -> out [batch*l, 512]."""
    assert q.shape[1] == 512 and k.shape[1] == 512 and v.shape[1] == 512 and q.shape[0] == batch * l
    if out is None:
        out = torch.empty((batch * l, 512), device=q.device, dtype=BF16)
    check(lib.fd_attention_d512(_p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0), batch, l,
                                scale, _stream()), "fd_attention_d512")
    return out


def softmax_rows(s, alpha=1.0, cols=None):
    """Row softmax over the first `cols` columns (default all); remaining columns of p are zero."""
    cols = cols or s.shape[1]
    p = torch.empty(s.shape, device=s.device, dtype=BF16) if cols == s.shape[1] else torch.zeros(s.shape, device=s.device, dtype=BF16)
    check(lib.fd_softmax_rows(_p(s), s.stride(0), _p(p), p.stride(0), s.shape[0], cols, alpha, _stream()), "fd_softmax_rows")
    return p


def groupnorm(x, gamma, beta, groups, eps, silu, x2=None, out=None):
    """GroupNorm(+SiLU) on NHWC bf16 [B,H,W,C] (contiguous); x2 concatenates along C."""
    assert x.is_contiguous() and x.dtype == BF16 and (x2 is None or x2.is_contiguous())
    B, C0 = x.shape[0], x.shape[-1]
    c1 = x2.shape[-1] if x2 is not None else 0
    hw = x.numel() // (B * C0)
    if out is None:
        out = torch.empty(x.shape[:-1] + (C0 + c1,), device=x.device, dtype=BF16)
    st0, st1 = getattr(x, "_fd_colstats", None), getattr(x2, "_fd_colstats", None) if x2 is not None else None
    if GN_STATS and st0 is not None and (x2 is None or st1 is not None) and ((C0 + c1) // groups) % 2 == 0 and (C0 + c1) // 8 <= 320:
        # one streaming pass: the producers of x (and x2) left the statistics
        work = torch.empty(2 * B * (C0 + c1), device=x.device, dtype=torch.float32)
        a1 = (_p(x2), c1, _p(st1.buf), st1.ld, st1.boxes_per_sample, st1.segments, st1.seg_stride) if x2 is not None else \
            (None, 0, None, 0, 0, 1, 0)
        check(lib.fd_groupnorm_nhwc_from_stats(_p(x), C0, _p(st0.buf), st0.ld, st0.boxes_per_sample, st0.segments, st0.seg_stride, *a1,
                                               _p(gamma), _p(beta), _p(out), B, hw, groups, eps, int(silu), _p(work), _stream()),
              "fd_groupnorm_nhwc_from_stats")
        return out
    work = torch.empty(lib.fd_groupnorm_work_floats(B, groups, C0 + c1), device=x.device, dtype=torch.float32)
    check(lib.fd_groupnorm_nhwc_fused(_p(x), C0, _p(x2), c1, _p(gamma), _p(beta), _p(out), B, hw, groups, eps, int(silu),
                                      _p(work), _p(_gn_sync(x.device)), _stream()), "fd_groupnorm_nhwc_fused")
    return out


_GN_SYNC = {}


def _gn_sync(device):
    """Per-device flag words of the one-launch GroupNorm: zeroed once, recycled by the kernel.  One buffer per device is
    enough because the path issues GroupNorm on one stream at a time (graph capture, replay and eager calls are ordered);
    two such kernels running CONCURRENTLY on one device are not supported.  Never allocated from a CUDA graph's private
    pool: the first call must be eager (the tile-step warm-up before capture is), else the three-launch path is used."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    s = _GN_SYNC.get(key)
    if s is None and not torch.cuda.is_current_stream_capturing():
        s = _GN_SYNC[key] = torch.zeros(lib.fd_groupnorm_sync_bytes() // 4, device=device, dtype=torch.int32)
    return s


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    assert x.is_contiguous() and x.dtype == BF16
    if out is None:
        out = torch.empty_like(x)
    cols = x.shape[-1]
    check(lib.fd_layernorm(_p(x), _p(gamma), _p(beta), _p(out), x.numel() // cols, cols, eps, _stream()), "fd_layernorm")
    return out


def axpby(a, b, alpha, beta, out):
    """out = alpha*a + beta*b on bf16 tensors whose last dim has unit stride and whose leading dims collapse to one row pitch
    (NHWC tensors or channel-slice views of them)."""
    for t in (a, b, out):
        assert t.dtype == BF16 and t.stride(-1) == 1 and t.shape == a.shape
        assert all(t.stride(i) == t.stride(i + 1) * t.shape[i + 1] for i in range(t.dim() - 2)), "rows must share one pitch"
    cols = a.shape[-1]
    rows = a.numel() // cols
    check(lib.fd_axpby_bf16(_p(a), a.stride(-2), _p(b), b.stride(-2), _p(out), out.stride(-2), rows, cols, alpha, beta, _stream()),
          "fd_axpby_bf16")
    return out


def upsample_nearest(x, oh=None, ow=None):
    B, H, W, Cc = x.shape
    oh, ow = oh or 2 * H, ow or 2 * W
    out = torch.empty((B, oh, ow, Cc), device=x.device, dtype=BF16)
    check(lib.fd_upsample_nearest_nhwc(_p(x), _p(out), B, H, W, oh, ow, Cc, _stream()), "fd_upsample_nearest_nhwc")
    return out


def nchw_to_nhwc(x):
    assert x.is_contiguous()
    B, Cc, H, W = x.shape
    out = torch.empty((B, H, W, Cc), device=x.device, dtype=BF16)
    check(lib.fd_nchw_to_nhwc(_p(x), _DT[x.dtype], _p(out), B, Cc, H * W, _stream()), "fd_nchw_to_nhwc")
    return out


def nhwc_to_nchw(x, dtype=BF16):
    assert x.is_contiguous() and x.dtype == BF16
    B, H, W, Cc = x.shape
    out = torch.empty((B, Cc, H, W), device=x.device, dtype=dtype)
    check(lib.fd_nhwc_to_nchw(_p(x), _p(out), _DT[dtype], B, Cc, H * W, _stream()), "fd_nhwc_to_nchw")
    return out


_IN_MODE = {torch.bfloat16: 1, torch.float32: 2, torch.float16: 3}


def conv3x3_small_in(x_nchw_or_nhwc, w, bias, nhwc_in=False, out=None):
    """cin <= 8. x: NCHW (bf16/fp16/fp32) or NHWC bf16 -> NHWC bf16 [B,H,W,cout] (out may be a channel-slice view)."""
    x = x_nchw_or_nhwc
    assert x.is_contiguous()
    if nhwc_in:
        B, H, W, cin = x.shape
        mode = 0
    else:
        B, cin, H, W = x.shape
        mode = _IN_MODE[x.dtype]
    cout = w.shape[0]
    if out is None:
        out = torch.empty((B, H, W, cout), device=x.device, dtype=BF16)
    if (not nhwc_in and cin <= 8 and cout % 16 == 0 and B * H * W >= 1 << 16 and x.dtype in _DT
            and __import__("os").environ.get("FD_SMALLCIN_TC", "1") == "1"):
        # tensor-core path (the UNet's conv_in over a launch group: the direct kernel took 0.49 ms for 14 x 128^2 x 4 -> 320):
        # NCHW -> NHWC with the channels zero-padded to 8, weights re-packed to 8 channels per tap (cached), implicit-GEMM conv
        key = (w.data_ptr(), w._version)
        ent = _PADDED_IN_W.get(key)
        if ent is None or ent[0] is not w:
            wp = torch.zeros((cout, 9, 8), device=w.device, dtype=BF16)
            wp[:, :, :cin] = w.view(cout, 9, cin)
            ent = _PADDED_IN_W[key] = (w, wp.view(cout, 72).contiguous())
        x8 = torch.empty((B, H, W, 8), device=x.device, dtype=BF16)
        check(lib.fd_nchw_to_nhwc_padded(_p(x), _DT[x.dtype], _p(x8), B, cin, 8, H * W, _stream()), "fd_nchw_to_nhwc_padded")
        return conv3x3(x8, ent[1], bias=bias, out=out)
    check(lib.fd_conv3x3_small(_p(x), mode, _p(w), _p(bias), _p(out), 0, out.stride(2), B, H, W, cin, cout, _stream()),
          "fd_conv3x3_small(in)")
    return out


_PADDED_IN_W = {}


_PADDED_OUT_W = {}


def conv3x3_small_out(x, w, bias, out_dtype=BF16, nchw_out=True):
    """cout <= 8. x NHWC bf16 -> NCHW out_dtype (or NHWC bf16).
    Large images with a bf16 / fp16 result run on the tensor cores: the weights are zero-padded to 32 output channels (cached), the
    implicit-GEMM conv writes [B,H,W,32] bf16 and a strided NHWC -> NCHW pass emits the first cout channels — the direct kernel
    (a warp per pixel, 9 x cin x cout FMAs through shuffles) took 0.85 ms for the UNet's 14 x 128^2 x 320 -> 4 conv_out and 2.5 ms
    for the VAE decoder's 8 x 512^2 x 128 -> 3, i.e. 0.2 TB/s on what is a 147 / 537 MB read."""
    assert x.is_contiguous() and x.dtype == BF16
    B, H, W, cin = x.shape
    cout = w.shape[0]
    if nchw_out and out_dtype in (BF16, torch.float16) and cin % 8 == 0 and B * H * W >= 1 << 16 and __import__("os").environ.get("FD_SMALLCOUT_TC", "1") == "1":
        key = (w.data_ptr(), w._version, None if bias is None else bias.data_ptr())
        ent = _PADDED_OUT_W.get(key)
        if ent is None or ent[0] is not w:
            wp = torch.zeros((32, w.shape[1]), device=w.device, dtype=BF16)
            wp[:cout] = w
            bp = torch.zeros(32, device=w.device, dtype=torch.float32)
            if bias is not None:
                bp[:cout] = bias
            ent = _PADDED_OUT_W[key] = (w, wp, bp)              # holds w: its address cannot be recycled while the entry lives
        y = conv3x3(x, ent[1], bias=ent[2])
        out = torch.empty((B, cout, H, W), device=x.device, dtype=out_dtype)
        check(lib.fd_nhwc_to_nchw_strided(_p(y), 32, _p(out), _DT[out_dtype], B, cout, H * W, _stream()), "fd_nhwc_to_nchw_strided")
        return out
    if nchw_out:
        out = torch.empty((B, cout, H, W), device=x.device, dtype=out_dtype)
        mode = _IN_MODE[out_dtype]
    else:
        out = torch.empty((B, H, W, cout), device=x.device, dtype=BF16)
        mode = 0
    check(lib.fd_conv3x3_small(_p(x), 0, _p(w), _p(bias), _p(out), mode, 0, B, H, W, cin, cout, _stream()),
          "fd_conv3x3_small(out)")
    return out


def conv1x1_small(x, w, bias, out_dtype=None, in_scale=1.0):
    assert x.is_contiguous()
    B, cin, H, W = x.shape
    cout = w.shape[0]
    out = torch.empty((B, cout, H, W), device=x.device, dtype=out_dtype or x.dtype)
    check(lib.fd_conv1x1_small(_p(x), _DT[x.dtype], _p(w), _p(bias), _p(out), _DT[out.dtype], B, cin, cout, H * W,
                               in_scale, _stream()), "fd_conv1x1_small")
    return out


def linear_smallm(x, w, bias=None, addend=None, silu_in=False, silu_out=False):
    assert x.dtype == torch.float32 and x.is_contiguous() and w.dtype == BF16 and w.is_contiguous()
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), device=x.device, dtype=torch.float32)
    for m0 in range(0, M, 8):
        m = min(8, M - m0)
        check(lib.fd_linear_smallm(_p(x[m0:]), _p(w), _p(bias), _p(addend[m0:]) if addend is not None else None,
                                   _p(out[m0:]), m, N, K, int(silu_in), int(silu_out), _stream()), "fd_linear_smallm")
    return out


def timestep_embedding(values, dim, out=None, inner=1, outer_ld=None):
    assert values.dtype == torch.float32 and values.is_contiguous()
    n = values.numel()
    if out is None:
        out = torch.empty((n, dim), device=values.device, dtype=torch.float32)
        outer_ld = dim
    check(lib.fd_timestep_embedding(_p(values), _p(out), n, dim, inner, outer_ld, _stream()), "fd_timestep_embedding")
    return out


def batch_axis_attention(qkv, groups, seq, hw, embed, heads):
    out = torch.empty((groups * seq * hw, embed), device=qkv.device, dtype=BF16)
    check(lib.fd_batch_axis_attention(_p(qkv), _p(out), groups, seq, hw, embed, heads, _stream()), "fd_batch_axis_attention")
    return out


# ------------------------------------------------------------------ tiled latents
def tile_gather(canvas, origins, k1, k2):
    _, c, h, w = canvas.shape
    T = origins.shape[0]
    tiles = torch.empty((T, c, k1, k2), device=canvas.device, dtype=canvas.dtype)
    check(lib.fd_tile_gather(_p(canvas), _p(tiles), _p(origins), T, c, h, w, k1, k2, canvas.element_size(), _stream()),
          "fd_tile_gather")
    return tiles


def tile_blend(tiles, weights, origins, h, w, lowp_dtype=None):
    T, c, k1, k2 = tiles.shape
    canvas = torch.empty((1, c, h, w), device=tiles.device, dtype=torch.float32)
    lowp = torch.empty((1, c, h, w), device=tiles.device, dtype=lowp_dtype) if lowp_dtype is not None else None
    check(lib.fd_tile_blend(_p(tiles), _DT[tiles.dtype], _p(weights), _p(origins), T, c, h, w, k1, k2, _p(canvas), _p(lowp),
                            _DT[lowp_dtype] if lowp_dtype is not None else 0, _stream()), "fd_tile_blend")
    return canvas, lowp


def cfg_euler_blend(eps, x_tiles, weights, origins, h, w, guidance, sigma, sigma_next, next_div, want_next=True):
    """x_tiles in eps.dtype (the reference's per-step cast) or fp32 (latent state kept in fp32); -> (canvas fp32, canvas in the
    model dtype, next tiles in x_tiles.dtype, next scaled tiles in the model dtype)."""
    T, c, k1, k2 = x_tiles.shape
    dev, dt = x_tiles.device, eps.dtype
    canvas = torch.empty((1, c, h, w), device=dev, dtype=torch.float32)
    lowp = torch.empty((1, c, h, w), device=dev, dtype=dt)
    nxt = torch.empty_like(x_tiles) if want_next else None
    nxs = torch.empty(x_tiles.shape, device=dev, dtype=dt) if want_next else None
    check(lib.fd_cfg_euler_blend(_p(eps), _p(x_tiles), _DT[dt], _DT[x_tiles.dtype], _p(weights), _p(origins), T, c, h, w, k1, k2,
                                 guidance, sigma, sigma_next, next_div, _p(canvas), _p(lowp), _p(nxt), _p(nxs), _stream()),
          "fd_cfg_euler_blend")
    return canvas, lowp, nxt, nxs


def tile_accumulate(tiles, eps, weights, origins, canvas_num, guidance=0.0, sigma=1.0, sigma_next=1.0):
    T, c, k1, k2 = tiles.shape
    _, _, h, w = canvas_num.shape
    check(lib.fd_tile_accumulate(_p(tiles), _p(eps), _DT[tiles.dtype], _p(weights), _p(origins), T, c, h, w, k1, k2, guidance,
                                 sigma, sigma_next, _p(canvas_num), _stream()), "fd_tile_accumulate")


def unit_accumulate(x_tiles, eps_units, weights, origins, unit_tile, unit_coef, unit_addx, canvas_num):
    U, c, k1, k2 = eps_units.shape
    _, _, h, w = canvas_num.shape
    check(lib.fd_unit_accumulate(_p(x_tiles), _DT[x_tiles.dtype], _p(eps_units), _DT[eps_units.dtype], _p(weights), _p(origins),
                                 _p(unit_tile), _p(unit_coef), _p(unit_addx), U, c, h, w, k1, k2, _p(canvas_num), _stream()),
          "fd_unit_accumulate")


def unit_numerator(x_tiles, eps_units, weights, origins, unit_tile, unit_coef, unit_addx, canvas_num):
    """canvas_num := this rank's blend numerator (overwrites; no zero-fill needed)."""
    U, c, k1, k2 = eps_units.shape
    _, _, h, w = canvas_num.shape
    check(lib.fd_unit_numerator(_p(x_tiles), _DT[x_tiles.dtype], _p(eps_units), _DT[eps_units.dtype], _p(weights), _p(origins),
                                _p(unit_tile), _p(unit_coef), _p(unit_addx), U, c, h, w, k1, k2, _p(canvas_num), _stream()),
          "fd_unit_numerator")


def canvas_finalize_retile(num, count, origins, k1, k2, next_div, lowp_dtype, state_dtype=None):
    """(canvas fp32, canvas lowp, next tiles in state_dtype, next scaled tiles in lowp_dtype) from the exchanged numerator."""
    _, c, h, w = num.shape
    T = origins.shape[0]
    state_dtype = state_dtype or lowp_dtype
    out = torch.empty_like(num)
    lowp = torch.empty(num.shape, device=num.device, dtype=lowp_dtype)
    nxt = torch.empty((T, c, k1, k2), device=num.device, dtype=state_dtype)
    nxs = torch.empty((T, c, k1, k2), device=num.device, dtype=lowp_dtype)
    check(lib.fd_canvas_finalize_retile(_p(num), _p(count), _p(out), _p(lowp), _DT[lowp_dtype], _DT[state_dtype], _p(origins), T, c, h,
                                        w, k1, k2, next_div, _p(nxt), _p(nxs), _stream()), "fd_canvas_finalize_retile")
    return out, lowp, nxt, nxs


def canvas_finalize(num, count, lowp_dtype):
    out = torch.empty_like(num)
    lowp = torch.empty(num.shape, device=num.device, dtype=lowp_dtype)
    check(lib.fd_canvas_finalize(_p(num), _p(count), _p(out), _p(lowp), _DT[lowp_dtype], num.numel(), _stream()),
          "fd_canvas_finalize")
    return out, lowp


def scale_div(x, div, out_dtype=None):
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    check(lib.fd_scale_div(_p(x), _DT[x.dtype], _p(out), _DT[out.dtype], div, x.numel(), _stream()), "fd_scale_div")
    return out


def seam_blend(a, b, extent, vertical):
    """In-place on b. a, b: [1,C,H,W] contiguous."""
    assert a.is_contiguous() and b.is_contiguous() and a.dtype == b.dtype
    check(lib.fd_seam_blend(_p(a), _p(b), _DT[a.dtype], a.shape[1], a.shape[2], a.shape[3], b.shape[2], b.shape[3], extent,
                            int(vertical), _stream()), "fd_seam_blend")
    return b


# ------------------------------------------------------------------ image / colour fix
def image_preprocess(rgb_u8, want_nchw_f32=False):
    h, w, _ = rgb_u8.shape
    nhwc = torch.empty((1, h, w, 3), device=rgb_u8.device, dtype=BF16)
    nchw = torch.empty((1, 3, h, w), device=rgb_u8.device, dtype=torch.float32) if want_nchw_f32 else None
    check(lib.fd_image_preprocess(_p(rgb_u8), _p(nhwc), _p(nchw), h, w, _stream()), "fd_image_preprocess")
    return nhwc, nchw


def image_postprocess(x, nhwc_bf16=True):
    if nhwc_bf16:
        _, h, w, _ = x.shape
    else:
        _, _, h, w = x.shape
    out = torch.empty((h, w, 3), device=x.device, dtype=torch.uint8)
    check(lib.fd_image_postprocess(_p(x), int(nhwc_bf16), _p(out), h, w, _stream()), "fd_image_postprocess")
    return out


def resample_u8_horizontal(img, bounds, taps, row0=0, rows=None):
    """Horizontal pass of the PIL-compatible resample: u8 HWC device image -> [rows, out_w, C] over source rows
    [row0, row0+rows).  bounds int32 [out_w,2], taps int32 [out_w,ksize] (image_process.resample_taps) on the device."""
    assert img.dtype == torch.uint8 and img.dim() == 3 and img.is_contiguous()
    H, W, Cc = img.shape
    rows = H - row0 if rows is None else rows
    assert 0 <= row0 and row0 + rows <= H
    out = torch.empty((rows, bounds.shape[0], Cc), device=img.device, dtype=torch.uint8)
    check(lib.fd_resample_u8_horizontal(_p(img[row0:]) if rows else _p(img), W, rows, _p(out), bounds.shape[0], Cc, _p(bounds),
                                        _p(taps), taps.shape[1], _stream()), "fd_resample_u8_horizontal")
    return out


def resample_u8_vertical(img, bounds, taps):
    """Vertical pass: u8 HWC device image -> [out_h, W, C]; bounds index the rows of `img`."""
    assert img.dtype == torch.uint8 and img.dim() == 3 and img.is_contiguous()
    out = torch.empty((bounds.shape[0], img.shape[1], img.shape[2]), device=img.device, dtype=torch.uint8)
    check(lib.fd_resample_u8_vertical(_p(img), img.shape[1] * img.shape[2], _p(out), bounds.shape[0], _p(bounds), _p(taps),
                                      taps.shape[1], _stream()), "fd_resample_u8_vertical")
    return out


def color_fix(target_u8, source_u8, mode):
    h, w, _ = target_u8.shape
    out = torch.empty_like(target_u8)
    work = torch.empty(lib.fd_colorfix_work_floats(h, w), device=target_u8.device, dtype=torch.float32)
    fn = lib.fd_adain_color_fix if mode == "adain" else lib.fd_wavelet_color_fix
    check(fn(_p(target_u8), _p(source_u8), _p(out), h, w, _p(work), _stream()), "fd_%s_color_fix" % mode)
    return out


def png_encode(img):
    """Complete PNG file of a u8 [H, W, C] device image (C = 1..4; rows may be strided, e.g. a crop of a larger image).
    Returns (file buffer u8 on the device, result int64 [2] on the device: bytes used, overflow flag) without synchronising."""
    assert img.dtype == torch.uint8 and img.dim() == 3 and img.is_cuda
    H, W, Cc = img.shape
    if img.stride(2) != 1 or img.stride(1) != Cc or img.stride(0) < W * Cc:
        img = img.contiguous()
    dev = img.device
    filtered = torch.empty(lib.fd_png_filtered_bytes(H, W, Cc), device=dev, dtype=torch.uint8)
    work = torch.empty((lib.fd_png_work_words(H, W, Cc) + 1) // 2, device=dev, dtype=torch.int64)
    cap = lib.fd_png_file_capacity(H, W, Cc)
    out = torch.empty(cap // 4, device=dev, dtype=torch.int32).view(torch.uint8)
    result = torch.empty(2, device=dev, dtype=torch.int64)
    check(lib.fd_png_encode(_p(img), H, W, Cc, img.stride(0), _p(filtered), _p(work), _p(out), cap, _p(result), _stream()),
          "fd_png_encode")
    return out, result


def launch_count():
    return int(lib.fd_launch_count())
