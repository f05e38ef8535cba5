"""Developer tool: fused layer1 block kernel vs conv3x3_c64 + chained conv3/conv1, batch 256 (cold L2)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
from irp_b200 import _lib
lib = _lib.init(0)
ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
B, H = 256, 56
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ms = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms += e0.elapsed_time(e1)
    return ms / iters * 1e3
for N2 in (64, 128):
    rn = lambda *s: torch.randn(*s, device="cuda")
    t1 = rn(B, H, H, 64).bfloat16(); w2 = (rn(64, 3, 3, 64) / 24).bfloat16(); b2 = rn(64)
    w3 = (rn(256, 64) / 8).bfloat16(); b3 = rn(256); res = rn(B, H, H, 256).bfloat16()
    w1 = (rn(N2, 256) / 16).bfloat16(); b1 = rn(N2)
    y = torch.empty(B, H, H, 256, device="cuda", dtype=torch.bfloat16); t1n = torch.empty(B, H, H, N2, device="cuda", dtype=torch.bfloat16)
    t2 = torch.empty(B, H, H, 64, device="cuda", dtype=torch.bfloat16)
    rows = B * H * H
    fused = lambda: _lib.check(lib.irp_l1_block(ptr(t1), ptr(w2), ptr(b2), ptr(w3), ptr(b3), ptr(res), ptr(y), ptr(w1), ptr(b1), ptr(t1n), B, H, H, N2, stream), "l1")
    c2 = lambda: _lib.check(lib.irp_conv2d_nhwc(ptr(t1), ptr(w2), ptr(b2), None, ptr(t2), B, H, H, 64, 64, 3, 1, 1, stream), "c2")
    ch = lambda: _lib.check(lib.irp_conv1x1_chain(ptr(t2), ptr(w3), ptr(b3), ptr(res), ptr(y), ptr(w1), ptr(b1), ptr(t1n), rows, 64, 256, N2, stream), "chain")
    print(f"N2 {N2}: fused {timeit(fused):7.1f} us | conv2 {timeit(c2):6.1f} + chain {timeit(ch):6.1f} us", flush=True)
