"""Developer tool: time the chained conv3+conv1 kernel against the two separate convolutions per junction shape."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
from irp_b200 import _lib
lib = _lib.init(0)
ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
B = 256
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
only = sys.argv[1] if len(sys.argv) > 1 else None
for tag, hw, K1, N1, N2 in [("L1", 56, 64, 256, 64), ("L1->L2", 56, 64, 256, 128), ("L2", 28, 128, 512, 128),
                            ("L2->L3", 28, 128, 512, 256), ("L3", 14, 256, 1024, 256)]:
    if only and tag != only: continue
    rows = B * hw * hw
    t2 = torch.randn(rows, K1, device="cuda").bfloat16(); w3 = (torch.randn(N1, K1, device="cuda") / 8).bfloat16()
    b3 = torch.randn(N1, device="cuda"); res = torch.randn(rows, N1, device="cuda").bfloat16()
    w1 = (torch.randn(N2, N1, device="cuda") / 16).bfloat16(); b1 = torch.randn(N2, device="cuda")
    y = torch.empty(rows, N1, device="cuda", dtype=torch.bfloat16); t1 = torch.empty(rows, N2, device="cuda", dtype=torch.bfloat16)
    def chain():
        _lib.check(lib.irp_conv1x1_chain(ptr(t2), ptr(w3), ptr(b3), ptr(res), ptr(y), ptr(w1), ptr(b1), ptr(t1), rows, K1, N1, N2, stream), "chain")
    def c3():
        _lib.check(lib.irp_conv2d_nhwc(ptr(t2), ptr(w3), ptr(b3), ptr(res), ptr(y), B, hw, hw, K1, N1, 1, 1, 1, stream), "c3")
    def c1():
        _lib.check(lib.irp_conv2d_nhwc(ptr(y), ptr(w1), ptr(b1), None, ptr(t1), B, hw, hw, N1, N2, 1, 1, 1, stream), "c1")
    tc, t3, t1_ = timeit(chain), timeit(c3), timeit(c1)
    byts = (rows * (K1 + 2 * N1 + N2)) * 2
    print(f"{tag:7s} K1 {K1} N1 {N1} N2 {N2}: chain {tc:7.1f} us | conv3 {t3:7.1f} + conv1 {t1_:7.1f} = {t3 + t1_:7.1f} us | hbm-min(chain) {byts / 6.55e6:6.1f}", flush=True)
