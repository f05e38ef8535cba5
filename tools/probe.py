"""GPU probe: exercises the CUDA path piece by piece against torch references and prints compact diagnostics.

Developer tool (run under gpurun), not part of the product or of the test-suite:
    python tools/probe.py conv_flat | conv_3x3 | conv_s2 | stem | resnet | preprocess | all
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200"))
sys.path.insert(0, ROOT)

from irp_b200 import _lib  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def conv_ref(x, w, bias, res, stride, relu):
    k = w.shape[1]
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride=stride, padding=k // 2)
    y = y.permute(0, 2, 3, 1)
    if res is not None:
        y = y + res.float()
    if relu:
        y = y.relu()
    return y


def run_conv(lib, B, H, W, Cin, Cout, k, stride, relu=True, residual=False, seed=0, tag=""):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, k, k, Cin, device="cuda", generator=g) / (k * k * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g)
    Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
    res = torch.randn(B, Ho, Wo, Cout, device="cuda", generator=g).bfloat16() if residual else None
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device="cuda").bfloat16()
    st = lib.irp_conv2d_nhwc(ptr(x), ptr(w), ptr(bias), ptr(res), ptr(out), B, H, W, Cin, Cout, k, stride, int(relu),
                             stream())
    _lib.check(st, "irp_conv2d_nhwc")
    torch.cuda.synchronize()
    ref = conv_ref(x, w, bias, res, stride, relu)
    o = out.float()
    nan = torch.isnan(o).sum().item()
    err = (o - ref).abs()
    scale = ref.abs().max().item()
    rel = (err.max() / scale).item() if nan == 0 else float("nan")
    ok = nan == 0 and rel < 2e-2
    print(f"[conv{tag}] B{B} {H}x{W} Cin{Cin} Cout{Cout} k{k} s{stride} relu{int(relu)} res{int(residual)}: "
          f"max_err/max_ref={rel:.3e} nan={nan} {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        e = torch.nan_to_num(err, nan=1e9).reshape(-1, Cout)
        rows_bad = (e.max(dim=1).values > 2e-2 * scale).nonzero().flatten()
        cols_bad = (e.max(dim=0).values > 2e-2 * scale).nonzero().flatten()
        print(f"   bad rows: {rows_bad.numel()}/{e.shape[0]} first {rows_bad[:16].tolist()}")
        print(f"   bad cols: {cols_bad.numel()}/{Cout} first {cols_bad[:16].tolist()}")
        torch.save({"x": x.cpu(), "w": w.cpu(), "bias": bias.cpu(), "out": out.cpu(), "ref": ref.cpu()},
                   os.path.join(OUT, f"fail_conv_{B}_{H}_{Cin}_{Cout}_{k}_{stride}.pt"))
    return ok


def t_conv_flat(lib):
    ok = True
    ok &= run_conv(lib, 1, 16, 16, 64, 64, 1, 1, relu=False)
    ok &= run_conv(lib, 1, 16, 16, 256, 128, 1, 1, relu=False)
    ok &= run_conv(lib, 3, 7, 7, 128, 256, 1, 1)
    ok &= run_conv(lib, 8, 56, 56, 64, 256, 1, 1, residual=True)
    ok &= run_conv(lib, 4, 14, 14, 1024, 256, 1, 1)
    return ok


def t_conv_3x3(lib):
    ok = True
    ok &= run_conv(lib, 2, 8, 8, 64, 64, 3, 1, relu=False)
    ok &= run_conv(lib, 2, 56, 56, 64, 64, 3, 1)
    ok &= run_conv(lib, 8, 28, 28, 128, 128, 3, 1)
    ok &= run_conv(lib, 32, 14, 14, 256, 256, 3, 1)
    ok &= run_conv(lib, 128, 7, 7, 512, 512, 3, 1)
    ok &= run_conv(lib, 6, 7, 7, 512, 512, 3, 1)
    ok &= run_conv(lib, 3, 14, 14, 256, 256, 3, 1)
    return ok


def t_conv_s2(lib):
    ok = True
    ok &= run_conv(lib, 2, 56, 56, 128, 128, 3, 2)
    ok &= run_conv(lib, 8, 28, 28, 256, 256, 3, 2)
    ok &= run_conv(lib, 32, 14, 14, 512, 512, 3, 2)
    ok &= run_conv(lib, 2, 56, 56, 256, 512, 1, 2, relu=False)
    ok &= run_conv(lib, 8, 14, 14, 1024, 2048, 1, 2, relu=False)
    return ok


class Net:
    """Random-init torchvision ResNet-50 trunk loaded into an irp_resnet50 handle."""

    def __init__(self, lib, max_batch, seed=1234):
        import torchvision
        self.lib = lib
        torch.manual_seed(seed)
        m = torchvision.models.resnet50(weights=None)
        self.torch_model = torch.nn.Sequential(*list(m.children())[:-1]).cuda().eval()
        self.full = m.cuda().eval()
        h = C.c_void_p()
        _lib.check(lib.irp_resnet50_create(C.byref(h), max_batch), "create")
        self.h = h
        convs = [(m.conv1, m.bn1)]
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                convs += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2), (blk.conv3, blk.bn3)]
                if blk.downsample is not None:
                    convs.append((blk.downsample[0], blk.downsample[1]))
        assert len(convs) == 53
        self.convs = convs
        for i, (cv, bn) in enumerate(convs):
            _lib.check(lib.irp_resnet50_load_conv(h, i, ptr(cv.weight.detach().contiguous()), ptr(bn.weight.detach()),
                                                  ptr(bn.bias.detach()), ptr(bn.running_mean), ptr(bn.running_var),
                                                  C.c_float(bn.eps), stream()), f"load_conv {i}")
        torch.cuda.synchronize()


def make_input(B, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, 3, 224, 224, device="cuda", generator=g)
    xb = x.bfloat16()
    xp = torch.zeros(B, 230, 230, 4, device="cuda", dtype=torch.bfloat16)
    xp[:, 3:227, 3:227, :3] = xb.permute(0, 2, 3, 1)
    return xb.float(), xp


def t_stem(lib):
    ok = True
    B = 4
    net = Net(lib, B)
    x, xp = make_input(B)
    emb = torch.empty(B, 2048, device="cuda")
    cap = torch.full((B, 112, 112, 64), float("nan"), device="cuda").bfloat16()
    _lib.check(lib.irp_resnet50_embed_capture(net.h, ptr(xp), B, ptr(emb), 0, ptr(cap), cap.numel(), stream()),
               "embed_capture")
    torch.cuda.synchronize()
    cv, bn = net.convs[0]
    with torch.no_grad():
        ref = F.relu(bn(cv(x))).permute(0, 2, 3, 1)
    err = (cap.float() - ref).abs()
    rel = (err.max() / ref.abs().max()).item()
    nan = torch.isnan(cap.float()).sum().item()
    print(f"[stem mode={os.environ.get('IRP_STEM_MODE', 'default')}] max_err/max_ref={rel:.3e} nan={nan} "
          f"{'OK' if rel < 2e-2 and nan == 0 else 'FAIL'}", flush=True)
    ok &= rel < 2e-2 and nan == 0
    if not ok:
        torch.save({"cap": cap.cpu(), "ref": ref.cpu()}, os.path.join(OUT, "fail_stem.pt"))
    return ok


def t_resnet(lib):
    B = 8
    net = Net(lib, B)
    x, xp = make_input(B)
    emb = torch.empty(B, 2048, device="cuda")
    ok = True
    with torch.no_grad():
        feats = {}
        m = net.full
        t = m.maxpool(m.relu(m.bn1(m.conv1(x))))
        feats[0] = None
        idx = 1
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                o1 = blk.relu(blk.bn1(blk.conv1(t)))
                o2 = blk.relu(blk.bn2(blk.conv2(o1)))
                o3 = blk.bn3(blk.conv3(o2))
                idt = t
                n = 3
                if blk.downsample is not None:
                    idt = blk.downsample(t)
                    feats[idx + 3] = idt
                    n = 4
                t = blk.relu(o3 + idt)
                feats[idx], feats[idx + 1], feats[idx + 2] = o1, o2, t
                idx += n
        ref = net.torch_model(x).flatten(1)
    for ci in [1, 2, 3, 4, 11, 12, 13, 14, 24, 25, 43, 44, 45, 46, 52]:
        r = feats[ci].permute(0, 2, 3, 1)
        cap = torch.full(r.shape, float("nan"), device="cuda").bfloat16().contiguous()
        _lib.check(lib.irp_resnet50_embed_capture(net.h, ptr(xp), B, ptr(emb), ci, ptr(cap), cap.numel(), stream()),
                   "embed_capture")
        torch.cuda.synchronize()
        rel = ((cap.float() - r).abs().max() / r.abs().max()).item()
        print(f"[resnet] conv {ci:2d} out {tuple(r.shape)} max_err/max_ref={rel:.3e}", flush=True)
    _lib.check(lib.irp_resnet50_embed(net.h, ptr(xp), B, ptr(emb), stream()), "embed")
    torch.cuda.synchronize()
    cos = F.cosine_similarity(emb, ref, dim=1)
    linf = ((emb - ref).abs().max(dim=1).values / ref.abs().max(dim=1).values)
    print(f"[resnet] embed cos min={cos.min().item():.6f} Linf-rel max={linf.max().item():.3e}", flush=True)
    ok &= cos.min().item() > 0.999 and linf.max().item() < 2e-2
    # timing at batch 256
    del net
    B = 256
    net = Net(lib, B)
    x, xp = make_input(B)
    emb = torch.empty(B, 2048, device="cuda")
    for _ in range(3):
        lib.irp_resnet50_embed(net.h, ptr(xp), B, ptr(emb), stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 10
    for _ in range(iters):
        lib.irp_resnet50_embed(net.h, ptr(xp), B, ptr(emb), stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"[resnet] batch {B}: {ms:.3f} ms -> {B / ms * 1e3:.0f} img/s, {8.1743e9 * B / ms / 1e9:.1f} TFLOP/s",
          flush=True)
    with torch.no_grad():
        tm = net.torch_model.to(memory_format=torch.channels_last).bfloat16()
        xx = x.bfloat16().contiguous(memory_format=torch.channels_last)
        for _ in range(3):
            tm(xx)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            tm(xx)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / iters
    print(f"[resnet] torch eager cuDNN bf16 channels_last batch {B}: {ms2:.3f} ms -> {B / ms2 * 1e3:.0f} img/s",
          flush=True)
    return ok


def t_preprocess(lib):
    from PIL import Image
    import torchvision
    tfm = torchvision.models.ResNet50_Weights.DEFAULT.transforms()
    rng = np.random.default_rng(0)
    sizes = [(224, 224), (300, 400), (480, 640), (150, 200), (400, 300), (57, 60), (1000, 700), (233, 232),
             (231, 500), (232, 232), (640, 232)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    offs, pos = [], 0
    for im in imgs:
        offs.append(pos)
        pos += (im.size + 127) // 128 * 128
    buf = np.zeros(pos, np.uint8)
    for o, im in zip(offs, imgs):
        buf[o:o + im.size] = im.reshape(-1)
    hw = np.array(sizes, np.int32)
    taps = max(_lib.geometry(h, w)[4] for h, w in sizes)
    d_pix = torch.from_numpy(buf).cuda()
    d_off = torch.tensor(offs, dtype=torch.int64).cuda()
    d_hw = torch.from_numpy(hw).cuda()
    n = len(imgs)
    ws_bytes = lib.irp_preprocess_workspace_bytes(n, taps)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    ref = torch.stack([tfm(Image.fromarray(im)) for im in imgs]).bfloat16()
    ok = True
    out = torch.full((n, 3, 224, 224), float("nan"), device="cuda").bfloat16()
    _lib.check(lib.irp_preprocess(ptr(d_pix), ptr(d_off), ptr(d_hw), n, taps, ptr(ws), ws_bytes, ptr(out),
                                  _lib.LAYOUT_NCHW, stream()), "preprocess")
    torch.cuda.synchronize()
    for i in range(n):
        neq = (out[i].cpu().view(torch.int16) != ref[i].view(torch.int16)).sum().item()
        print(f"[preprocess NCHW] {sizes[i]} mismatches={neq}", flush=True)
        ok &= neq == 0
    out2 = torch.full((n, 230, 230, 4), float("nan"), device="cuda").bfloat16()
    _lib.check(lib.irp_preprocess(ptr(d_pix), ptr(d_off), ptr(d_hw), n, taps, ptr(ws), ws_bytes, ptr(out2),
                                  _lib.LAYOUT_NHWC4P, stream()), "preprocess")
    torch.cuda.synchronize()
    exp = torch.zeros(n, 230, 230, 4, dtype=torch.bfloat16)
    exp[:, 3:227, 3:227, :3] = ref.permute(0, 2, 3, 1)
    neq = (out2.cpu().view(torch.int16) != exp.view(torch.int16)).sum().item()
    print(f"[preprocess NHWC4P] mismatches={neq}", flush=True)
    ok &= neq == 0
    # timing: 256 images 300x400
    B = 256
    h, w = 300, 400
    per = (h * w * 3 + 127) // 128 * 128
    d_pix = torch.randint(0, 256, (B * per,), dtype=torch.uint8, device="cuda")
    d_off = (torch.arange(B, dtype=torch.int64) * per).cuda()
    d_hw = torch.tensor([[h, w]] * B, dtype=torch.int32).cuda()
    taps = _lib.geometry(h, w)[4]
    ws_bytes = lib.irp_preprocess_workspace_bytes(B, taps)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(B, 230, 230, 4, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        lib.irp_preprocess(ptr(d_pix), ptr(d_off), ptr(d_hw), B, taps, ptr(ws), ws_bytes, ptr(out), 1, stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lib.irp_preprocess(ptr(d_pix), ptr(d_off), ptr(d_hw), B, taps, ptr(ws), ws_bytes, ptr(out), 1, stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    alg = B * (3 * 290 * 290 + 301056)
    print(f"[preprocess] {B}x{h}x{w}: {ms:.3f} ms, {B / ms * 1e3:.0f} img/s, algorithmic {alg / ms / 1e6:.0f} GB/s",
          flush=True)
    return ok


def synth_features(n, d=2048, seed=0):
    """Embedding-like matrix: a few dominant directions, a slowly decaying tail, positive mean."""
    rng = np.random.default_rng(seed)
    r = 256
    basis = np.linalg.qr(rng.standard_normal((d, r)))[0]
    sv = np.concatenate([[150.0, 60.0, 30.0], 12.0 * np.arange(1, r - 2) ** -0.6])
    x = (rng.standard_normal((n, r)) * sv) @ basis.T + 0.05 * rng.standard_normal((n, d)) + 3.0
    return np.maximum(x, 0).astype(np.float32)  # relu-like, many exact zeros


def gpu_pca(lib, x_t, k, shift_t):
    n, d = x_t.shape
    cnt = torch.zeros(1, dtype=torch.float64, device="cuda")
    ssum = torch.zeros(d, dtype=torch.float64, device="cuda")
    scat = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    wsb = lib.irp_cov_workspace_bytes(n, d)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    _lib.check(lib.irp_cov_accumulate(ptr(x_t), n, d, ptr(shift_t), ptr(cnt), ptr(ssum), ptr(scat), ptr(ws), wsb,
                                      stream()), "cov_accumulate")
    mean = torch.empty(d, dtype=torch.float64, device="cuda")
    comps = torch.empty(k, d, dtype=torch.float64, device="cuda")
    ev = torch.empty(k + 1, dtype=torch.float64, device="cuda")
    wsb2 = lib.irp_pca_fit_workspace_bytes(d, k)
    ws2 = torch.empty(wsb2, dtype=torch.uint8, device="cuda")
    _lib.check(lib.irp_pca_fit(ptr(cnt), ptr(ssum), ptr(scat), ptr(shift_t), d, k, ptr(mean), ptr(comps), ptr(ev),
                               ptr(ws2), wsb2, stream()), "pca_fit")
    z = torch.empty(n, k, dtype=torch.float32, device="cuda")
    _lib.check(lib.irp_pca_transform(ptr(x_t), n, d, ptr(mean), ptr(comps), k, ptr(z), stream()), "pca_transform")
    torch.cuda.synchronize()
    return mean, comps, ev, z, (cnt, ssum, scat)


def t_pca(lib):
    from oracle import pca_ref
    ok = True
    for (n, k) in [(1500, 50), (700, 20)]:
        x = synth_features(n, seed=n)
        x_t = torch.from_numpy(x).cuda()
        shift_t = x_t[:256].mean(0).contiguous()
        mean, comps, ev, z, (cnt, ssum, scat) = gpu_pca(lib, x_t, k, shift_t)
        ref = pca_ref.pca_fit(x, k)
        xs = x.astype(np.float64) - shift_t.cpu().numpy().astype(np.float64)
        s_ref = xs.T @ xs
        s_gpu = np.triu(scat.cpu().numpy())
        s_gpu = s_gpu + np.triu(s_gpu, 1).T
        ds = s_gpu - s_ref
        print(f"[pca] scatter: max|dS|/max|S|={np.abs(ds).max() / np.abs(s_ref).max():.2e}, ||dS||_2/(n-1)="
              f"{np.linalg.norm(ds, 2) / (n - 1):.2e}, gap(k)={ref.eigenvalues[k - 1] - ref.eigenvalues[k]:.3e}, "
              f"mean signed rel err={np.mean(ds / (np.abs(s_ref) + 1e-30)):.2e}", flush=True)
        ang = pca_ref.subspace_angle(comps.cpu().numpy(), ref.components)
        ev_rel = np.abs(ev[:k].cpu().numpy() - ref.explained_variance).max() / ref.explained_variance[0]
        ev_rel_each = (np.abs(ev[:k].cpu().numpy() - ref.explained_variance) / ref.explained_variance).max()
        tv_rel = abs(ev[k].item() - ref.eigenvalues.sum()) / ref.eigenvalues.sum()
        mean_err = np.abs(mean.cpu().numpy() - ref.mean).max()
        cosv = np.abs((comps.cpu().numpy() * ref.components).sum(1))
        sign_ok = bool(((comps.cpu().numpy() * ref.components).sum(1) > 0).all())
        zref = pca_ref.pca_transform(x, ref.mean, ref.components)
        zerr = np.abs(z.cpu().numpy() - zref).max() / np.abs(zref).max()
        orth = np.abs(comps.cpu().numpy() @ comps.cpu().numpy().T - np.eye(k)).max()
        good = ang < 1e-3 and ev_rel_each < 1e-4 and sign_ok and zerr < 1e-3
        print(f"[pca] n={n} k={k}: subspace angle={ang:.3e} rad, eval rel err(max)={ev_rel_each:.2e} "
              f"(vs top {ev_rel:.2e}), total var rel={tv_rel:.2e}, mean err={mean_err:.2e}, min|cos|={cosv.min():.6f}, "
              f"signs {'ok' if sign_ok else 'BAD'}, orth err={orth:.2e}, Z rel err={zerr:.2e} "
              f"{'OK' if good else 'FAIL'}", flush=True)
        ok &= good
    # timing at 27k rows
    n, k = 27000, 50
    x_t = torch.from_numpy(synth_features(4096, seed=7)).cuda().repeat(7, 1)[:n].contiguous()
    x_t += 0.01 * torch.randn_like(x_t)
    shift_t = x_t[:256].mean(0).contiguous()
    d = 2048
    cnt = torch.zeros(1, dtype=torch.float64, device="cuda")
    ssum = torch.zeros(d, dtype=torch.float64, device="cuda")
    scat = torch.zeros(d, d, dtype=torch.float64, device="cuda")
    wsb = lib.irp_cov_workspace_bytes(n, d)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    mean = torch.empty(d, dtype=torch.float64, device="cuda")
    comps = torch.empty(k, d, dtype=torch.float64, device="cuda")
    ev = torch.empty(k + 1, dtype=torch.float64, device="cuda")
    wsb2 = lib.irp_pca_fit_workspace_bytes(d, k)
    ws2 = torch.empty(wsb2, dtype=torch.uint8, device="cuda")
    z = torch.empty(n, k, dtype=torch.float32, device="cuda")
    ev_t = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for it in range(2):
        cnt.zero_(); ssum.zero_(); scat.zero_()
        ev_t[0].record()
        lib.irp_cov_accumulate(ptr(x_t), n, d, ptr(shift_t), ptr(cnt), ptr(ssum), ptr(scat), ptr(ws), wsb, stream())
        ev_t[1].record()
        lib.irp_pca_fit(ptr(cnt), ptr(ssum), ptr(scat), ptr(shift_t), d, k, ptr(mean), ptr(comps), ptr(ev), ptr(ws2),
                        wsb2, stream())
        ev_t[2].record()
        lib.irp_pca_transform(ptr(x_t), n, d, ptr(mean), ptr(comps), k, ptr(z), stream())
        ev_t[3].record()
        torch.cuda.synchronize()
    print(f"[pca] n={n}: cov {ev_t[0].elapsed_time(ev_t[1]):.2f} ms, fit(eig) {ev_t[1].elapsed_time(ev_t[2]):.2f} ms, "
          f"transform {ev_t[2].elapsed_time(ev_t[3]):.2f} ms", flush=True)
    return ok


def gpu_lof(lib, z_t, groups_t, n_groups, k, cont):
    n, d = z_t.shape
    scores = torch.empty(n, dtype=torch.float64, device="cuda")
    offs = torch.empty(n_groups, dtype=torch.float64, device="cuda")
    flags = torch.empty(n, dtype=torch.uint8, device="cuda")
    wsb = lib.irp_lof_workspace_bytes(n, d, k)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    _lib.check(lib.irp_lof(ptr(z_t), n, d, ptr(groups_t), n_groups, k, C.c_double(cont), ptr(scores), ptr(offs),
                           ptr(flags), ptr(ws), wsb, stream()), "lof")
    torch.cuda.synchronize()
    return scores, offs, flags


def t_lof(lib):
    from oracle import lof_ref
    from sklearn.neighbors import LocalOutlierFactor
    ok = True
    rng = np.random.default_rng(3)
    n, d, G = 3000, 50, 10
    y = rng.integers(0, G, n)
    centers = rng.standard_normal((G, d)) * 4
    z = (centers[y] + rng.standard_normal((n, d)) * rng.uniform(0.5, 2.0, (n, 1))).astype(np.float32)
    z[:12] = z[20]  # duplicates
    z_t = torch.from_numpy(z).cuda()
    # global
    sc, off, fl = gpu_lof(lib, z_t, None, 1, 75, 0.03)
    lof = LocalOutlierFactor(n_neighbors=75, contamination=0.03)
    pred = lof.fit_predict(z) == -1
    s_ref = lof.negative_outlier_factor_.astype(np.float64)
    rel = np.abs(sc.cpu().numpy() - s_ref).max() / np.abs(s_ref).max()
    band = np.abs(s_ref - lof.offset_) < 1e-3 * abs(lof.offset_)
    mism = ((fl.cpu().numpy() != 0) != pred) & ~band
    o_ref = lof_ref.lof_scores(z, 75)
    print(f"[lof global] score max rel err vs sklearn={rel:.2e}, vs oracle={np.abs(sc.cpu().numpy()-o_ref).max():.2e}, "
          f"offset {off[0].item():.8f} vs {lof.offset_:.8f}, flagged {int(fl.sum())} vs {int(pred.sum())}, "
          f"mismatches outside band={int(mism.sum())} {'OK' if mism.sum() == 0 and rel < 1e-5 else 'FAIL'}", flush=True)
    ok &= mism.sum() == 0 and rel < 1e-5
    # per class
    g_t = torch.from_numpy(y.astype(np.int32)).cuda()
    sc, off, fl = gpu_lof(lib, z_t, g_t, G, 30, 0.05)
    bad = 0
    worst = 0.0
    for c in range(G):
        m = y == c
        lof = LocalOutlierFactor(n_neighbors=30, contamination=0.05)
        pred = lof.fit_predict(z[m]) == -1
        s_ref = lof.negative_outlier_factor_.astype(np.float64)
        worst = max(worst, np.abs(sc.cpu().numpy()[m] - s_ref).max() / np.abs(s_ref).max())
        band = np.abs(s_ref - lof.offset_) < 1e-3 * abs(lof.offset_)
        bad += int((((fl.cpu().numpy()[m] != 0) != pred) & ~band).sum())
    print(f"[lof per-class] worst score rel err={worst:.2e}, mismatches outside band={bad} "
          f"{'OK' if bad == 0 and worst < 1e-5 else 'FAIL'}", flush=True)
    ok &= bad == 0 and worst < 1e-5
    # small groups (k clipped) + centroid scorer
    n2 = 400
    y2 = np.concatenate([np.zeros(20, np.int64), np.ones(31, np.int64), rng.integers(2, 5, n2 - 51)])
    z2 = rng.standard_normal((n2, 16)).astype(np.float32)
    sc, off, fl = gpu_lof(lib, torch.from_numpy(z2).cuda(), torch.from_numpy(y2.astype(np.int32)).cuda(), 5, 30, 0.05)
    import warnings
    bad = 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for c in range(5):
            m = y2 == c
            lof = LocalOutlierFactor(n_neighbors=30, contamination=0.05)
            pred = lof.fit_predict(z2[m]) == -1
            band = np.abs(lof.negative_outlier_factor_ - lof.offset_) < 1e-3 * abs(lof.offset_)
            bad += int((((fl.cpu().numpy()[m] != 0) != pred) & ~band).sum())
    print(f"[lof small groups] mismatches outside band={bad} {'OK' if bad == 0 else 'FAIL'}", flush=True)
    ok &= bad == 0
    dist = torch.empty(n, dtype=torch.float64, device="cuda")
    zs = torch.empty(n, dtype=torch.float64, device="cuda")
    thr = torch.empty(G, dtype=torch.float64, device="cuda")
    fl2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    wsb = lib.irp_centroid_workspace_bytes(n, d, G)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    _lib.check(lib.irp_centroid_zscore(ptr(z_t), n, d, ptr(g_t), G, C.c_double(0.05), ptr(dist), ptr(zs), ptr(thr),
                                       ptr(fl2), ptr(ws), wsb, stream()), "centroid")
    torch.cuda.synchronize()
    rd, rz, rt, rf = lof_ref.centroid_zscore(z, y, G, 0.05)
    e1 = np.abs(dist.cpu().numpy() - rd).max()
    e2 = np.abs(zs.cpu().numpy() - rz).max()
    e3 = np.abs(thr.cpu().numpy() - rt).max()
    fm = int(((fl2.cpu().numpy() != 0) != rf).sum())
    print(f"[centroid] dist err={e1:.2e} zscore err={e2:.2e} thr err={e3:.2e} flag mismatches={fm} "
          f"{'OK' if fm == 0 and e1 < 1e-9 else 'FAIL'}", flush=True)
    ok &= fm == 0 and e1 < 1e-9
    # timing 27k x 50
    n = 27000
    y = rng.integers(0, G, n)
    z = (centers[y] + rng.standard_normal((n, d))).astype(np.float32)
    z_t = torch.from_numpy(z).cuda()
    g_t = torch.from_numpy(y.astype(np.int32)).cuda()
    e0, e1_, e2_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for _ in range(2):
        e0.record()
        gpu_lof(lib, z_t, g_t, G, 30, 0.05)
        e1_.record()
        gpu_lof(lib, z_t, None, 1, 75, 0.03)
        e2_.record()
        torch.cuda.synchronize()
    print(f"[lof] n={n} d={d}: per-class {e0.elapsed_time(e1_):.2f} ms, global {e1_.elapsed_time(e2_):.2f} ms", flush=True)
    return ok


TESTS = {"conv_flat": t_conv_flat, "conv_3x3": t_conv_3x3, "conv_s2": t_conv_s2, "stem": t_stem, "resnet": t_resnet,
         "preprocess": t_preprocess, "pca": t_pca, "lof": t_lof}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    lib = _lib.init(0)
    print("device:", torch.cuda.get_device_name(0), flush=True)
    names = list(TESTS) if which == "all" else [which]
    rc = 0
    for nme in names:
        t0 = time.time()
        try:
            good = TESTS[nme](lib)
        except Exception as ex:  # noqa: BLE001
            print(f"[{nme}] EXCEPTION {type(ex).__name__}: {ex}", flush=True)
            good = False
        print(f"== {nme}: {'PASS' if good else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
        rc |= 0 if good else 1
    sys.exit(rc)
