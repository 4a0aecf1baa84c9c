"""Test helper: builds and binds tests/png_host_emul.cpp (the device PNG encoder's phase functions run on the host)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "png_host_emul.cpp")
CORE = os.path.join(HERE, "..", "faithdiff_dev_b200", "csrc", "png_core.cuh")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libpng_host_emul.so")
_lib = None


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    stale = force or not os.path.exists(LIB) or any(os.path.getmtime(f) > os.path.getmtime(LIB) for f in (SRC, CORE))
    if stale:
        tmp = LIB + ".tmp%d" % os.getpid()
        try:
            subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-o", tmp, SRC], check=True)
            os.replace(tmp, LIB)
        except (OSError, subprocess.CalledProcessError):
            # a copied tree can look stale by mtime alone: a library built by __graft_entry__.build() is still the right one
            if not os.path.exists(LIB):
                raise
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        lib.emul_png_file_capacity.restype = C.c_longlong
        lib.emul_png_file_capacity.argtypes = [C.c_int, C.c_int, C.c_int]
        lib.emul_png_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_longlong,
                                        C.POINTER(C.c_longlong), C.c_void_p]
        lib.emul_huffman_lengths.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.emul_set_thread_order.argtypes = [C.c_int, C.c_uint]
        _lib = lib
    return _lib


def encode(a, capacity=None, thread_order=0, seed=1):
    """a: u8 HW / HWC numpy array (any row stride) -> (png file bytes, filtered scanline stream).
    thread_order: order of a block's threads inside each phase (0 ascending, 1 descending, 2 shuffled per phase)."""
    lib = load()
    lib.emul_set_thread_order(thread_order, seed)
    if a.ndim == 2:
        a = a[..., None]
    assert a.dtype == np.uint8 and a.strides[2] == 1 and a.strides[1] == a.shape[2]
    h, w, c = a.shape
    cap = lib.emul_png_file_capacity(h, w, c) if capacity is None else capacity
    out = np.zeros(cap, np.uint8)
    filt = np.zeros(h * (w * c + 1), np.uint8)
    n = C.c_longlong(0)
    rc = lib.emul_png_encode(a.ctypes.data, h, w, c, a.strides[0], out.ctypes.data, cap, C.byref(n), filt.ctypes.data)
    if rc != 0:
        raise OverflowError(n.value)
    return out[:n.value].tobytes(), filt.tobytes()


def huffman_lengths(freq, maxbits):
    lib = load()
    f = np.ascontiguousarray(freq, dtype=np.uint32)
    lens = np.zeros(len(f), np.uint8)
    lib.emul_huffman_lengths(f.ctypes.data, len(f), maxbits, lens.ctypes.data)
    return lens


def test_images(seed=0):
    """Shapes and contents both suites run: ragged sizes, 1-4 channels, flat / noisy / mixed, runs of every match length."""
    rng = np.random.default_rng(seed)

    def smooth(h, w, c, noise=4.0):
        y, x = np.mgrid[0:h, 0:w]
        base = np.stack([128 + 100 * np.sin(x / 37.0 + k) * np.cos(y / 23.0 - k) for k in range(c)], -1)
        return np.clip(base + rng.normal(0, noise, (h, w, c)), 0, 255).astype(np.uint8)

    runs = np.zeros((70, 700, 1), np.uint8)               # row r: runs of length r + 2 of alternating values
    for r in range(70):
        v = (np.arange(700) // (r + 2)) % 2
        runs[r, :, 0] = v * (3 * r + 1)
    mixed = smooth(200, 300, 3)
    mixed[40:120, 50:250] = 255                           # a saturated box inside a noisy image
    mixed[150:, :] = 0
    skew = np.zeros((64, 1024, 1), np.uint8)              # Fibonacci-like byte statistics: forces the 15-bit limit
    vals, cnt, a, b = [], [], 1, 1
    for k in range(24):
        vals.append(k * 7 + 1); cnt.append(a); a, b = b, a + b
    p = np.array(cnt, np.float64) / sum(cnt)
    skew[..., 0] = rng.choice(np.array(vals, np.uint8), size=(64, 1024), p=p)
    big = smooth(520, 640, 3)
    return {
        "smooth_rgb": smooth(64, 64, 3), "ragged_rgb": smooth(37, 53, 3), "one_pixel": smooth(1, 1, 3), "one_row_grey": smooth(1, 7000, 1),
        "one_column": smooth(300, 1, 3), "rgba": smooth(100, 211, 4), "grey_alpha": smooth(128, 130, 2), "zeros": np.zeros((200, 200, 3), np.uint8),
        "noise": rng.integers(0, 256, (257, 129, 3), dtype=np.uint8), "runs": runs, "mixed": mixed, "skew": skew,
        "crop_view": big[3:515, 5:600],                   # strided rows, as `gen_image.crop(...)` of the padded result
        "exact_block": smooth(4, 4095, 1),                # filtered stream = exactly one 16384-byte block
        "block_plus_one": rng.integers(0, 256, (1, 16384, 1), dtype=np.uint8),
    }
