"""Developer tool: CUDA-event timing of every distinct convolution of the ResNet-50 trunk at one batch size
through irp_conv2d_nhwc (the single-conv parity hook), with TFLOP/s and the per-layer lower bounds."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
from irp_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lib = _lib.init(0)
ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

shapes = []  # (tag, count, H, Cin, Cout, k, stride, residual)
planes = [64, 128, 256, 512]; blocks = [3, 4, 6, 3]; inpl = 64; hw = 56
seen = {}
for l in range(4):
    for b in range(blocks[l]):
        s = 2 if (b == 0 and l > 0) else 1
        w = planes[l]
        layer = [("c1", hw, inpl, w, 1, 1, False), ("c2", hw, w, w, 3, s, False),
                 ("c3", hw // s, w, w * 4, 1, 1, True)]
        if b == 0:
            layer.append(("ds", hw, inpl, w * 4, 1, s, False))
        for role, h, ci, co, k, st, res in layer:
            key = (h, ci, co, k, st, res)
            if key not in seen:
                seen[key] = [f"L{l+1}.{role}", 0]
            seen[key][1] += 1
        inpl = w * 4; hw //= s
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
total = 0.0
for (h, ci, co, k, st, res), (tag, cnt) in seen.items():
    ho = h // st
    x = torch.randn(B, h, h, ci, device="cuda").bfloat16()
    wgt = (torch.randn(co, k, k, ci, device="cuda") / (k * k * ci) ** 0.5).bfloat16()
    bias = torch.randn(co, device="cuda")
    r = torch.randn(B, ho, ho, co, device="cuda").bfloat16() if res else None
    out = torch.empty(B, ho, ho, co, device="cuda", dtype=torch.bfloat16)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def run():
        _lib.check(lib.irp_conv2d_nhwc(ptr(x), ptr(wgt), ptr(bias), ptr(r), ptr(out), B, h, h, ci, co, k, st, 1,
                                       stream), "conv")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(iters):
        flush.zero_()  # evict the operands from L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    us = ms / iters * 1e3
    fl = 2.0 * B * ho * ho * co * ci * k * k
    byts = (B * h * h * ci + B * ho * ho * co * (2 if res else 1)) * 2
    print(f"{tag:6s} x{cnt} {ci:4d}->{co:4d} k{k}s{st} @{h:2d} res{int(res)}: {us:8.1f} us {fl/us/1e6:7.1f} TF/s "
          f"(tensor-min {fl/1354e6:6.1f}, hbm-min {byts/6.55e6:6.1f})", flush=True)
    total += us * cnt
print(f"sum over the 52 convs (cold L2): {total:.0f} us")
