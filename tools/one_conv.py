"""Developer tool: a few launches of ONE convolution shape through irp_conv2d_nhwc (for ncu captures).
usage: one_conv.py B H Cin Cout k stride residual [iters]"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
from irp_b200 import _lib
B, h, ci, co, k, st, res = (int(a) for a in sys.argv[1:8])
iters = int(sys.argv[8]) if len(sys.argv) > 8 else 4
lib = _lib.init(0)
ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
ho = h // st
x = torch.randn(B, h, h, ci, device="cuda").bfloat16()
w = (torch.randn(co, k, k, ci, device="cuda") / (k * k * ci) ** 0.5).bfloat16()
bias = torch.randn(co, device="cuda")
r = torch.randn(B, ho, ho, co, device="cuda").bfloat16() if res else None
out = torch.empty(B, ho, ho, co, device="cuda", dtype=torch.bfloat16)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(iters):
    if i == iters - 1:
        e0.record()
    _lib.check(lib.irp_conv2d_nhwc(ptr(x), ptr(w), ptr(bias), ptr(r), ptr(out), B, h, h, ci, co, k, st, 1, stream), "conv")
e1.record(); torch.cuda.synchronize()
print(f"{sys.argv[1:8]}: last launch {e0.elapsed_time(e1)*1e3:.1f} us")
