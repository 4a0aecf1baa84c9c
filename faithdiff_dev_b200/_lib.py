"""ctypes binding of libfaithdiff_b200.so (the C-ABI in include/faithdiff_b200.h).

There is NO fallback: if the shared library is missing the import raises, and every op raises RuntimeError
with fd_last_error() on a non-zero status.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FAITHDIFF_B200_LIB: A/B-compare two builds of the library in one process launch each (development only)
LIB_PATH = os.environ.get("FAITHDIFF_B200_LIB") or os.path.join(_HERE, "libfaithdiff_b200.so")


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p * 2), ("a_cols", C.c_int64 * 2), ("a_ld", C.c_int64 * 2), ("b", C.c_void_p),
        ("M", C.c_int64), ("N", C.c_int64), ("mode", C.c_int32), ("batch", C.c_int32), ("in_h", C.c_int32),
        ("in_w", C.c_int32), ("pad_lo", C.c_int32), ("act", C.c_int32), ("bias_along_m", C.c_int32),
        ("out_fp32", C.c_int32), ("bias", C.c_void_p), ("rowbias", C.c_void_p), ("rows_per_batch", C.c_int64), ("ld_rowbias", C.c_int64),
        ("residual", C.c_void_p), ("ld_res", C.c_int64), ("out", C.c_void_p), ("ld_out", C.c_int64),
        ("alpha", C.c_float), ("block_n", C.c_int32),
    ]


class GemmLnDesc(C.Structure):
    """fd_gemm_ln_desc (LayerNorm folded into the GEMMs around it; opt-in)."""
    _fields_ = [
        ("stats_in", C.c_void_p), ("stats_in_ld", C.c_int64), ("stats_in_n", C.c_int32), ("pad0_", C.c_int32),
        ("colsum", C.c_void_p), ("inv_c", C.c_float), ("eps", C.c_float),
        ("stats_out", C.c_void_p), ("stats_out_ld", C.c_int64), ("stats_out_n", C.c_int32), ("pad1_", C.c_int32),
        ("colstats_out", C.c_void_p), ("colstats_ld", C.c_int64), ("colstats_boxes", C.c_int32), ("colstats_boxes_per_image", C.c_int32),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "faithdiff_dev_b200: %s not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the CUDA path)" % LIB_PATH)
    return C.CDLL(LIB_PATH)


lib = _load()
lib.fd_last_error.restype = C.c_char_p
lib.fd_launch_count.restype = C.c_int64
lib.fd_groupnorm_work_floats.restype = C.c_int64
lib.fd_groupnorm_sync_bytes.restype = C.c_int64
lib.fd_peer_region_bytes.restype = C.c_int64
lib.fd_colorfix_work_floats.restype = C.c_int64
lib.fd_png_filtered_bytes.restype = C.c_int64
lib.fd_png_work_words.restype = C.c_int64
lib.fd_png_file_capacity.restype = C.c_int64

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p
_SIGS = {
    "fd_gemm": [C.POINTER(GemmDesc), vp],
    "fd_gemm_ln": [C.POINTER(GemmDesc), C.POINTER(GemmLnDesc), vp],
    "fd_attention_d64": [vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, f32, vp],
    "fd_attention_d64_causal": [vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, f32, vp],
    "fd_attention_d512": [vp, i64, vp, i64, vp, i64, vp, i64, i32, i32, f32, vp],
    "fd_softmax_rows": [vp, i64, vp, i64, i64, i64, f32, vp],
    "fd_groupnorm_work_floats": [i32, i32, i32],
    "fd_groupnorm_nhwc": [vp, i32, vp, i32, vp, vp, vp, i32, i64, i32, f32, i32, vp, vp],
    "fd_groupnorm_sync_bytes": [],
    "fd_groupnorm_nhwc_fused": [vp, i32, vp, i32, vp, vp, vp, i32, i64, i32, f32, i32, vp, vp, vp],
    "fd_groupnorm_nhwc_from_stats": [vp, i32, vp, i64, i32, i32, i64, vp, i32, vp, i64, i32, i32, i64, vp, vp, vp, i32, i64, i32, f32, i32, vp, vp],
    "fd_layernorm": [vp, vp, vp, vp, i64, i32, f32, vp],
    "fd_axpby_bf16": [vp, i64, vp, i64, vp, i64, i64, i32, f32, f32, vp],
    "fd_upsample_nearest_nhwc": [vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "fd_nchw_to_nhwc": [vp, i32, vp, i32, i32, i64, vp],
    "fd_nchw_to_nhwc_padded": [vp, i32, vp, i32, i32, i32, i64, vp],
    "fd_nhwc_to_nchw": [vp, vp, i32, i32, i32, i64, vp],
    "fd_nhwc_to_nchw_strided": [vp, i64, vp, i32, i32, i32, i64, vp],
    "fd_conv3x3_small": [vp, i32, vp, vp, vp, i32, i64, i32, i32, i32, i32, i32, vp],
    "fd_conv1x1_small": [vp, i32, vp, vp, vp, i32, i32, i32, i32, i64, f32, vp],
    "fd_linear_smallm": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "fd_timestep_embedding": [vp, vp, i32, i32, i32, i64, vp],
    "fd_batch_axis_attention": [vp, vp, i32, i32, i64, i32, i32, vp],
    "fd_tile_gather": [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, vp],
    "fd_tile_blend": [vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, i32, vp],
    "fd_cfg_euler_blend": [vp, vp, i32, i32, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, f32, f32, vp, vp, vp, vp, vp],
    "fd_tile_accumulate": [vp, vp, i32, vp, vp, i32, i32, i32, i32, i32, i32, f32, f32, f32, vp, vp],
    "fd_unit_accumulate": [vp, i32, vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp],
    "fd_canvas_finalize": [vp, vp, vp, vp, i32, i64, vp],
    "fd_unit_numerator": [vp, i32, vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp],
    "fd_canvas_finalize_retile": [vp, vp, vp, vp, i32, i32, vp, i32, i32, i32, i32, i32, i32, f32, vp, vp, vp],
    "fd_scale_div": [vp, i32, vp, i32, f32, i64, vp],
    "fd_peer_region_bytes": [i64],
    "fd_peer_alloc": [C.POINTER(vp), i64],
    "fd_peer_free": [vp],
    "fd_peer_ipc_handle": [vp, vp],
    "fd_peer_ipc_open": [vp, C.POINTER(vp)],
    "fd_peer_ipc_close": [vp],
    "fd_canvas_exchange_finalize": [C.POINTER(vp), i32, i32, i32, C.c_uint32, i64, vp, vp, vp, i32, vp],
    "fd_seam_blend": [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp],
    "fd_image_preprocess": [vp, vp, vp, i32, i32, vp],
    "fd_image_postprocess": [vp, i32, vp, i32, i32, vp],
    "fd_resample_u8_horizontal": [vp, i32, i64, vp, i32, i32, vp, vp, i32, vp],
    "fd_resample_u8_vertical": [vp, i64, vp, i32, vp, vp, i32, vp],
    "fd_colorfix_work_floats": [i32, i32],
    "fd_adain_color_fix": [vp, vp, vp, i32, i32, vp, vp],
    "fd_wavelet_color_fix": [vp, vp, vp, i32, i32, vp, vp],
    "fd_png_filtered_bytes": [i32, i32, i32],
    "fd_png_work_words": [i32, i32, i32],
    "fd_png_file_capacity": [i32, i32, i32],
    "fd_png_encode": [vp, i32, i32, i32, i64, vp, vp, vp, i64, vp, vp],
}
EXPORTS = ["fd_last_error", "fd_version", "fd_launch_count", "fd_sizeof_gemm_desc", "fd_sizeof_gemm_ln_desc"] + list(_SIGS)
for _n, _a in _SIGS.items():
    getattr(lib, _n).argtypes = _a


if lib.fd_sizeof_gemm_desc() != C.sizeof(GemmDesc):
    raise ImportError("faithdiff_dev_b200: GemmDesc mirror (%d B) does not match fd_gemm_desc (%d B)"
                      % (C.sizeof(GemmDesc), lib.fd_sizeof_gemm_desc()))


if lib.fd_sizeof_gemm_ln_desc() != C.sizeof(GemmLnDesc):
    raise ImportError("faithdiff_dev_b200: GemmLnDesc mirror (%d B) does not match fd_gemm_ln_desc (%d B)"
                      % (C.sizeof(GemmLnDesc), lib.fd_sizeof_gemm_ln_desc()))


def check(status, what):
    if status != 0:
        raise RuntimeError("%s failed (%d): %s" % (what, status, lib.fd_last_error().decode()))
