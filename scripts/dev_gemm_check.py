"""Dev check (GPU): fd_gemm (GEMM + implicit conv modes) vs torch on the same bf16 inputs, plus timing."""
import ctypes as C, os, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from faithdiff_dev_b200._lib import GemmDesc
lib = C.CDLL("faithdiff_dev_b200/libfaithdiff_b200.so")
lib.fd_last_error.restype = C.c_char_p
lib.fd_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
dev = "cuda"
torch.manual_seed(0)

def run(d):
    st = lib.fd_gemm(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if st: raise RuntimeError(lib.fd_last_error().decode())

def relerr(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)).item()

def gemm_case(M, N, K, bias=False, act=0, res=False, block_n=0, fp32=False, time_it=False, rowbias=False):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    b = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    out = torch.zeros(M, N // 2 if act == 2 else N, device=dev, dtype=torch.float32 if fp32 else torch.bfloat16)
    d = GemmDesc()
    d.a[0] = a.data_ptr(); d.a_cols[0] = K; d.a_ld[0] = K; d.b = b.data_ptr(); d.M = M; d.N = N; d.mode = 0
    d.out = out.data_ptr(); d.ld_out = out.shape[1]; d.alpha = 1.0; d.act = act; d.out_fp32 = int(fp32); d.block_n = block_n
    ref = a.float() @ b.float().t()
    if bias:
        bv = torch.randn(N, device=dev); d.bias = bv.data_ptr(); ref = ref + bv
    if rowbias:
        rb = torch.randn(2, N, device=dev); d.rowbias = rb.data_ptr(); d.rows_per_batch = M // 2
        ref = ref + rb.repeat_interleave(M // 2, 0)
    if act == 1: ref = F.silu(ref)
    if act == 3: ref = ref * torch.sigmoid(1.702 * ref)
    if act == 2: ref = ref[:, 0::2] * F.gelu(ref[:, 1::2])
    if res:
        r = torch.randn(M, N, device=dev).to(torch.bfloat16); d.residual = r.data_ptr(); d.ld_res = N; ref = ref + r.float()
    run(d); torch.cuda.synchronize()
    e = relerr(out, ref)
    msg = "gemm M=%d N=%d K=%d bias=%d act=%d res=%d bn=%d fp32=%d relerr=%.3e" % (M, N, K, bias, act, res, block_n, fp32, e)
    if time_it:
        for _ in range(3): run(d)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(20): run(d)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        msg += "  %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9)
        c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3): torch.matmul(a, b.t(), out=c)
        e0.record()
        for _ in range(20): torch.matmul(a, b.t(), out=c)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        msg += "  | cublas %.3f ms %.1f TFLOP/s" % (ms, 2.0 * M * N * K / ms / 1e9)
    print(msg, "OK" if e < 1e-2 else "FAIL", flush=True)
    return e < 1e-2

def conv_case(B, H, W, C0, C1, N, stride=1, pad_lo=1, bias=True, time_it=False):
    x0 = torch.randn(B, H, W, C0, device=dev).to(torch.bfloat16)
    x1 = torch.randn(B, H, W, C1, device=dev).to(torch.bfloat16) if C1 else None
    Ct = C0 + C1
    w = (torch.randn(N, Ct, 3, 3, device=dev) / (9 * Ct) ** 0.5).to(torch.bfloat16)
    wp = w.permute(0, 2, 3, 1).reshape(N, 9 * Ct).contiguous()
    bv = torch.randn(N, device=dev)
    xin = torch.cat([x0] + ([x1] if C1 else []), dim=-1).permute(0, 3, 1, 2).float()
    if stride == 1:
        ref = F.conv2d(xin, w.float(), bv if bias else None, padding=1)
    elif pad_lo == 1:
        ref = F.conv2d(xin, w.float(), bv if bias else None, stride=2, padding=1)
    else:
        ref = F.conv2d(F.pad(xin, (0, 1, 0, 1)), w.float(), bv if bias else None, stride=2)
    Ho, Wo = ref.shape[2], ref.shape[3]
    out = torch.zeros(B, Ho, Wo, N, device=dev, dtype=torch.bfloat16)
    d = GemmDesc()
    d.a[0] = x0.data_ptr(); d.a_cols[0] = C0; d.a_ld[0] = C0
    if C1: d.a[1] = x1.data_ptr(); d.a_cols[1] = C1; d.a_ld[1] = C1
    d.b = wp.data_ptr(); d.N = N; d.mode = 1 if stride == 1 else 2; d.batch = B; d.in_h = H; d.in_w = W; d.pad_lo = pad_lo
    if bias: d.bias = bv.data_ptr()
    d.out = out.data_ptr(); d.ld_out = N; d.alpha = 1.0
    run(d); torch.cuda.synchronize()
    e = relerr(out.permute(0, 3, 1, 2), ref)
    msg = "conv B=%d H=%d W=%d C=%d+%d N=%d stride=%d pad_lo=%d relerr=%.3e" % (B, H, W, C0, C1, N, stride, pad_lo, e)
    if time_it:
        for _ in range(3): run(d)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): run(d)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        msg += "  %.3f ms  %.1f TFLOP/s" % (ms, 2.0 * B * Ho * Wo * N * 9 * Ct / ms / 1e9)
    print(msg, "OK" if e < 1e-2 else "FAIL", flush=True)
    return e < 1e-2

ok = True
ok &= gemm_case(128, 64, 64)
ok &= gemm_case(128, 256, 128, block_n=256)
ok &= gemm_case(256, 320, 640, bias=True)
ok &= gemm_case(154, 1280, 2048)
ok &= gemm_case(2048, 1280, 1280, bias=True, res=True, time_it=True)
ok &= gemm_case(2048, 10240, 1280, bias=True, act=2, time_it=True)
ok &= gemm_case(2048, 1280, 5120, bias=True, res=True, time_it=True)
ok &= gemm_case(8192, 1920, 640, time_it=True)
ok &= gemm_case(8192, 8192, 8192, time_it=True, block_n=256)
ok &= gemm_case(4096, 512, 512, fp32=True, bias=True)
ok &= gemm_case(1024, 640, 320, act=1, rowbias=True, bias=True)
ok &= gemm_case(1024, 1280, 640, act=3, bias=True)
ok &= gemm_case(2, 1280, 320, bias=True)
ok &= conv_case(1, 16, 16, 64, 0, 64)
ok &= conv_case(2, 32, 32, 128, 0, 128)
ok &= conv_case(2, 18, 14, 64, 0, 96)
ok &= conv_case(2, 32, 32, 64, 128, 64)
ok &= conv_case(2, 32, 32, 320, 0, 320, stride=2, pad_lo=1)
ok &= conv_case(1, 64, 48, 128, 0, 128, stride=2, pad_lo=0)
ok &= conv_case(2, 128, 128, 320, 0, 320, time_it=True)
ok &= conv_case(2, 32, 32, 1280, 1280, 1280, time_it=True)
ok &= conv_case(2, 64, 64, 640, 0, 640, time_it=True)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
