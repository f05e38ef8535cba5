"""GPU parity: the CUDA path (through the C ABI / irp_b200 custom ops / drop-in functions) against the oracle,
the committed golden fixtures, and size-independent properties at the BASELINE sizes.

Tolerances (BASELINE.json north_star):
  preprocess   bit-exact with the reference transform after its float32 result is rounded to bf16
  embeddings   cosine >= 0.999 per image and max|a-b|/max|b| <= 2e-2 (bf16 operands, fp32 accumulate)
  PCA          top-k subspace angle <= 1e-3 rad vs the exact fp64 solver, same component signs
  LOF flags    identical except rows whose score is within 1e-3 (relative) of the threshold
"""
import ctypes as C
import os
import warnings

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from oracle import lof_ref, pca_ref, pil_resample, stage_ref, synth

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COS_MIN = 0.999
LINF_REL_MAX = 2e-2
ANGLE_MAX = 1e-3
BAND = 1e-3


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.fixture(scope="module")
def ops_mod(lib):
    from irp_b200 import ops
    return ops


@pytest.fixture(scope="module")
def trunk(lib):
    from irp_b200.stage import ResNet50Trunk
    return ResNet50Trunk(stage_ref.full_resnet50(seed=1234), torch.device("cuda:0"), max_batch=64)


# =============================================================================================== A1 preprocess
def _preprocess(ops_mod, images, layout):
    from irp_b200.stage import pack_images
    p = pack_images(images).to("cuda:0")
    return ops_mod.preprocess(p.pixels, p.offsets, p.hw, p.max_taps, layout)


def _expected_bf16(images):
    return torch.from_numpy(np.stack([pil_resample.transform(im) for im in images])).bfloat16()


def test_preprocess_golden_bit_exact(ops_mod):
    g = load_golden("preprocess.npz")
    images = [np.random.default_rng(int(s)).integers(0, 256, (int(h), int(w), 3), dtype=np.uint8)
              for (h, w), s in zip(g["sizes"], g["seeds"])]
    out = _preprocess(ops_mod, images, 0).cpu()
    exp = torch.from_numpy(np.stack([pil_resample.normalize(c) for c in g["crops"]])).bfloat16()
    assert torch.equal(out.view(torch.int16), exp.view(torch.int16))


@pytest.mark.parametrize("sizes", [
    [(224, 224), (232, 232), (233, 232), (232, 640)],          # no resize / pure crop / one-axis resize
    [(57, 60), (120, 500), (500, 120), (150, 200)],             # upscales, extreme aspect ratios
    [(1000, 700), (2400, 1800), (231, 500), (3000, 2900)],      # many-tap downscales
])
def test_preprocess_matches_oracle_ragged_batches(ops_mod, sizes):
    rng = np.random.default_rng(len(sizes) + sizes[0][0])
    images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in sizes]
    exp = _expected_bf16(images)
    out = _preprocess(ops_mod, images, 0).cpu()
    assert torch.equal(out.view(torch.int16), exp.view(torch.int16))
    padded = _preprocess(ops_mod, images, 1).cpu()
    want = torch.zeros(len(images), 230, 230, 4, dtype=torch.bfloat16)
    want[:, 3:227, 3:227, :3] = exp.permute(0, 2, 3, 1)
    assert torch.equal(padded.view(torch.int16), want.view(torch.int16))


def test_preprocess_single_image_and_transform_callable(lib):
    from PIL import Image
    from functions import data_curation as dc
    img = np.random.default_rng(5).integers(0, 256, (310, 415, 3), dtype=np.uint8)
    t = dc.B200Transform("cuda:0")(Image.fromarray(img))
    assert t.shape == (3, 224, 224) and t.dtype == torch.float32
    exp = torch.from_numpy(pil_resample.transform(img)).bfloat16().float()
    assert torch.equal(t.cpu(), exp)


def test_preprocess_full_size_properties(ops_mod):
    """Config-2-shaped batch (256 mixed-resolution images): constant images stay constant (weights sum to one),
    borders/pad channel are exactly zero, and the result does not depend on the batch an image travels in."""
    hw = synth.mixed_resolution_sizes(256, seed=0)
    rng = np.random.default_rng(0)
    images = []
    for i, (h, w) in enumerate(hw):
        if i % 2 == 0:
            images.append(np.full((h, w, 3), rng.integers(0, 256, 3, dtype=np.uint8), dtype=np.uint8))
        else:
            images.append(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
    out = _preprocess(ops_mod, images, 1)
    assert (out[:, :3].float().abs().sum() + out[:, 227:].float().abs().sum() + out[:, :, :3].float().abs().sum()
            + out[:, :, 227:].float().abs().sum() + out[..., 3].float().abs().sum()).item() == 0.0
    for i in range(0, 256, 2):
        inner = out[i, 3:227, 3:227, :3].float().cpu()
        colour = pil_resample.normalize(images[i][:1, :1])[:, 0, 0]  # float32 [3]
        want = torch.from_numpy(colour).bfloat16().float()
        assert torch.equal(inner, want.expand(224, 224, 3)), f"constant image {i} {hw[i]}"
    solo = _preprocess(ops_mod, [images[37]], 1)
    assert torch.equal(solo[0].view(torch.int16), out[37].view(torch.int16))


# =============================================================================================== A2 convolutions
def _conv_ref(x, w, bias, res, stride, relu):
    k = w.shape[1]
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias, stride=stride, padding=k // 2)
    y = y.permute(0, 2, 3, 1)
    if res is not None:
        y = y + res.float()
    return y.relu() if relu else y


@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,relu,residual", [
    (1, 16, 64, 64, 1, 1, False, False),      # smallest flat GEMM
    (3, 7, 128, 256, 1, 1, True, False),      # M = 147: ragged last tile
    (8, 56, 64, 256, 1, 1, True, True),       # layer1 conv3 + residual
    (4, 14, 1024, 256, 1, 1, True, False),    # long K
    (2, 56, 64, 64, 3, 1, True, False),       # layer1 3x3
    (8, 28, 128, 128, 3, 1, True, False),     # layer2 3x3: patch-resident CTA-pair kernel, ragged 4 x 2 tiles
    (1, 28, 128, 128, 3, 1, False, False),    # one image: 8 tiles, no ReLU
    (3, 20, 128, 128, 3, 1, True, False),     # odd tile count (phantom tile in the last pair), ragged both ways
    (67, 28, 128, 128, 3, 1, True, False),    # several tiles per CTA pair (ring / accumulator recycling)
    (32, 14, 256, 256, 3, 1, True, False),
    (128, 7, 512, 512, 3, 1, True, False),    # (1,1,128) boxes
    (6, 7, 512, 512, 3, 1, True, False),      # (7,7,2) boxes, 98 of 128 rows
    (3, 14, 256, 256, 3, 1, True, False),     # batch not a multiple of the box
    (2, 56, 128, 128, 3, 2, True, False),     # stride-2 3x3 through the parity views
    (8, 28, 256, 256, 3, 2, True, False),
    (2, 56, 256, 512, 1, 2, False, False),    # stride-2 downsample
    (8, 14, 1024, 2048, 1, 2, False, False),
])
def test_conv2d_matches_torch(lib, B, H, Cin, Cout, k, stride, relu, residual):
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + Cin + k)
    x = torch.randn(B, H, H, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(Cout, k, k, Cin, device="cuda", generator=g) / (k * k * Cin) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device="cuda", generator=g)
    Ho = (H + 2 * (k // 2) - k) // stride + 1
    res = torch.randn(B, Ho, Ho, Cout, device="cuda", generator=g).bfloat16() if residual else None
    out = torch.full((B, Ho, Ho, Cout), float("nan"), device="cuda").bfloat16()
    from irp_b200 import _lib
    _lib.check(lib.irp_conv2d_nhwc(_ptr(x), _ptr(w), _ptr(bias), _ptr(res), _ptr(out), B, H, H, Cin, Cout, k, stride,
                                   int(relu), _stream()), "irp_conv2d_nhwc")
    torch.cuda.synchronize()
    ref = _conv_ref(x, w, bias, res, stride, relu)
    assert not torch.isnan(out.float()).any()
    # bf16 output rounding (2^-9) + fp32 accumulation order
    assert ((out.float() - ref).abs().max() / ref.abs().max()).item() < 6e-3


@pytest.mark.parametrize("rows,K1,N1,N2", [
    (128, 64, 128, 64),          # one tile, one pass
    (300, 64, 128, 64),          # 3 M tiles: ragged last tile + a phantom tile in the last CTA pair
    (1000, 64, 256, 64),         # layer1 junction, ragged
    (2 * 3136, 64, 256, 128),    # layer1 -> layer2 junction
    (3 * 784, 128, 512, 128),    # layer2 junction
    (2 * 784, 128, 512, 256),    # layer2 -> layer3 junction
    (5 * 196, 256, 1024, 256),   # layer3 junction (8 passes)
    (40000, 64, 256, 64),        # many tiles per CTA pair: ring / accumulator recycling
])
def test_conv1x1_chain_matches_torch(lib, rows, K1, N1, N2):
    """conv3 + residual + ReLU chained with the next block's conv1 + ReLU (conv_chain.cuh) against fp32 torch, with
    the intermediate rounded to bf16 exactly where the kernel rounds it."""
    from irp_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(rows + K1 + N1 + N2)
    t2 = torch.randn(rows, K1, device="cuda", generator=g).bfloat16()
    w3 = (torch.randn(N1, K1, device="cuda", generator=g) / K1 ** 0.5).bfloat16()
    b3 = torch.randn(N1, device="cuda", generator=g)
    res = torch.randn(rows, N1, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(N2, N1, device="cuda", generator=g) / N1 ** 0.5).bfloat16()
    b1 = torch.randn(N2, device="cuda", generator=g)
    y = torch.full((rows, N1), float("nan"), device="cuda").bfloat16()
    t1 = torch.full((rows, N2), float("nan"), device="cuda").bfloat16()
    _lib.check(lib.irp_conv1x1_chain(_ptr(t2), _ptr(w3), _ptr(b3), _ptr(res), _ptr(y), _ptr(w1), _ptr(b1), _ptr(t1),
                                     rows, K1, N1, N2, _stream()), "irp_conv1x1_chain")
    torch.cuda.synchronize()
    y_ref = (t2.float() @ w3.float().t() + b3 + res.float()).relu()
    assert not torch.isnan(y.float()).any() and not torch.isnan(t1.float()).any()
    assert ((y.float() - y_ref).abs().max() / y_ref.abs().max()).item() < 6e-3
    t1_ref = (y.float() @ w1.float().t() + b1).relu()       # from the kernel's own bf16 intermediate
    assert ((t1.float() - t1_ref).abs().max() / t1_ref.abs().max()).item() < 6e-3


@pytest.mark.parametrize("rows,K1,K2,N1,N2", [
    (256, 64, 64, 256, 64),        # layer1's first block, one tile pair
    (3136 * 5 + 77, 64, 64, 256, 64),   # ragged tail, several tiles per pair at a small batch
    (40000, 64, 64, 256, 64),      # ring / accumulator recycling
    (1000, 128, 256, 512, 128),    # general K1 != K2
])
def test_conv1x1_chain_with_folded_shortcut_matches_torch(lib, rows, K1, K2, N1, N2):
    """conv3 with the block's stride-1 shortcut conv folded into the same accumulator (no residual tensor), ReLU,
    chained with the next block's conv1 + ReLU, against fp32 torch."""
    from irp_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(rows + K1 + K2 + N1 + N2)
    t2 = torch.randn(rows, K1, device="cuda", generator=g).bfloat16()
    x = torch.randn(rows, K2, device="cuda", generator=g).bfloat16()
    w3 = (torch.randn(N1, K1, device="cuda", generator=g) / K1 ** 0.5).bfloat16()
    wds = (torch.randn(N1, K2, device="cuda", generator=g) / K2 ** 0.5).bfloat16()
    wcat = torch.cat([w3, wds], 1).contiguous()
    bias = torch.randn(N1, device="cuda", generator=g)
    w1 = (torch.randn(N2, N1, device="cuda", generator=g) / N1 ** 0.5).bfloat16()
    b1 = torch.randn(N2, device="cuda", generator=g)
    y = torch.full((rows, N1), float("nan"), device="cuda").bfloat16()
    t1 = torch.full((rows, N2), float("nan"), device="cuda").bfloat16()
    _lib.check(lib.irp_conv1x1_chain_ds(_ptr(t2), _ptr(x), _ptr(wcat), _ptr(bias), _ptr(y), _ptr(w1), _ptr(b1),
                                        _ptr(t1), rows, K1, K2, N1, N2, _stream()), "irp_conv1x1_chain_ds")
    torch.cuda.synchronize()
    y_ref = (t2.float() @ w3.float().t() + x.float() @ wds.float().t() + bias).relu()
    assert not torch.isnan(y.float()).any() and not torch.isnan(t1.float()).any()
    assert ((y.float() - y_ref).abs().max() / y_ref.abs().max()).item() < 6e-3
    t1_ref = (y.float() @ w1.float().t() + b1).relu()
    assert ((t1.float() - t1_ref).abs().max() / t1_ref.abs().max()).item() < 6e-3


@pytest.mark.parametrize("batch,cout", [(1, 128), (3, 256), (7, 2048), (64, 2048), (256, 2048)])
def test_conv1x1_pool_matches_torch(lib, batch, cout):
    """The trunk's last kernel: conv3 (512 -> Cout @7x7) + bias + residual + ReLU + global average pool, fp32 out,
    against fp32 torch; ragged last image triple, several triples per CTA, and bitwise repeatability (the pooled sum
    is a fixed-order chain inside one thread)."""
    from irp_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(batch * 7 + cout)
    rn = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
    t2 = rn(batch * 49, 512).relu().bfloat16()
    w = (rn(cout, 512) / 22).bfloat16()
    b = rn(cout) * 0.5
    res = rn(batch * 49, cout).relu().bfloat16()
    out = torch.full((batch, cout), float("nan"), device="cuda")
    call = lambda o: _lib.check(lib.irp_conv1x1_pool(_ptr(t2), _ptr(w), _ptr(b), _ptr(res), _ptr(o), batch, 512, cout,
                                                     _stream()), "irp_conv1x1_pool")
    call(out)
    torch.cuda.synchronize()
    ref = (t2.float() @ w.float().t() + b + res.float()).relu().view(batch, 49, cout).mean(1)
    assert not torch.isnan(out).any()
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 2e-3
    out2 = torch.empty_like(out)
    call(out2)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)


def test_conv2d_rejects_unsupported_shapes(lib):
    x = torch.zeros(1, 8, 8, 48, device="cuda", dtype=torch.bfloat16)
    st = lib.irp_conv2d_nhwc(_ptr(x), _ptr(x), _ptr(x), None, _ptr(x), 1, 8, 8, 48, 64, 1, 1, 0, _stream())
    assert st == 1 and b"multiples of 64" in lib.irp_last_error()
    st = lib.irp_conv2d_nhwc(_ptr(x), _ptr(x), _ptr(x), None, _ptr(x), 1, 8, 8, 64, 64, 5, 1, 0, _stream())
    assert st == 1


# =============================================================================================== A2 trunk
def _golden_images():
    g = load_golden("embeddings.npz")
    hw = g["hw"]
    sizes = hw[:, 0].astype(np.int64) * hw[:, 1] * 3
    offs = np.concatenate([[0], np.cumsum(sizes)])
    return g, [g["pixels"][offs[i]:offs[i + 1]].reshape(hw[i, 0], hw[i, 1], 3) for i in range(len(hw))]


def _embedding_metrics(a, b):
    cos = (a * b).sum(1) / np.linalg.norm(a, axis=1) / np.linalg.norm(b, axis=1)
    linf = np.abs(a - b).max(1) / np.abs(b).max(1)
    return cos.min(), linf.max()


def test_embeddings_match_reference_golden(trunk):
    from irp_b200.stage import OutlierStage, pack_images
    g, images = _golden_images()
    stage = OutlierStage(trunk, batch_size=5)
    feats = stage.embed_packed(pack_images(images), from_host=True).cpu().numpy()
    cos, linf = _embedding_metrics(feats, g["features"])
    assert cos >= COS_MIN and linf <= LINF_REL_MAX, (cos, linf)


def test_embed_lanes_match_single_lane_bitwise(trunk):
    """Batches alternating between two / three trunk handles on their own streams (CudaBackend(lanes=...)) give the
    same bits as one handle on one stream, from device-resident and from pinned host input."""
    from irp_b200.stage import CudaBackend, OutlierStage, pack_images
    from oracle import synth
    rng = np.random.default_rng(5)
    sizes = [(224, 224), (300, 400), (180, 150), (260, 233), (500, 310)]
    images = [synth.smooth_image(rng, *sizes[i % 5], i % 3) for i in range(37)]
    host = pack_images(images)
    dev = host.to(trunk.device, non_blocking=False)
    ref = OutlierStage(CudaBackend(trunk, lanes=1), batch_size=8).embed_packed(dev)
    torch.cuda.synchronize()
    for lanes in (2, 3):
        stage = OutlierStage(CudaBackend(trunk, lanes=lanes), batch_size=8)
        for src, from_host in ((dev, False), (host, True)):
            got = stage.embed_packed(src, from_host=from_host)
            torch.cuda.synchronize()
            assert torch.equal(got, ref), (lanes, from_host, float((got - ref).abs().max()))


def test_trunk_matches_torchvision_per_layer(lib, trunk):
    """Every conv's fused output (bias/ReLU/residual epilogue) against torchvision evaluated in fp32 on the GPU.
    Index 0 is compared after the max pool (the stem kernel fuses it; the unpooled tensor never exists).  With a
    capture every conv runs as its own launch: layer1's first shortcut conv, which the product path folds into the
    junction kernel, is covered by test_conv1x1_chain_with_folded_shortcut_matches_torch and, end to end, by the
    golden-embedding test."""
    from irp_b200 import _lib
    from irp_b200.stage import conv_bn_pairs
    m = stage_ref.full_resnet50(seed=1234).cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    B = 4
    x = torch.randn(B, 3, 224, 224, device="cuda", generator=g).bfloat16()
    xp = torch.zeros(B, 230, 230, 4, device="cuda", dtype=torch.bfloat16)
    xp[:, 3:227, 3:227, :3] = x.permute(0, 2, 3, 1)
    feats = {}
    with torch.no_grad():
        t = m.maxpool(m.relu(m.bn1(m.conv1(x.float()))))
        feats[0] = t
        idx = 1
        for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
            for blk in layer:
                o1 = blk.relu(blk.bn1(blk.conv1(t)))
                o2 = blk.relu(blk.bn2(blk.conv2(o1)))
                idt, n = t, 3
                if blk.downsample is not None:
                    idt, n = blk.downsample(t), 4
                    feats[idx + 3] = idt
                t = blk.relu(blk.bn3(blk.conv3(o2)) + idt)
                feats[idx], feats[idx + 1], feats[idx + 2] = o1, o2, t
                idx += n
        ref_embed = torch.flatten(m.avgpool(t), 1)
    assert idx == 53 and len(conv_bn_pairs(m)) == 53
    emb = torch.empty(B, 2048, device="cuda")
    worst = 0.0
    for ci in range(53):
        r = feats[ci].permute(0, 2, 3, 1).contiguous()
        cap = torch.full(r.shape, float("nan"), device="cuda").bfloat16()
        _lib.check(lib.irp_resnet50_embed_capture(C.c_void_p(trunk.handle), _ptr(xp), B, _ptr(emb), ci, _ptr(cap),
                                                  cap.numel(), _stream()), "embed_capture")
        torch.cuda.synchronize()
        rel = ((cap.float() - r).abs().max() / r.abs().max()).item()
        worst = max(worst, rel)
        assert rel < LINF_REL_MAX, f"conv {ci}: {rel}"
    cos, linf = _embedding_metrics(emb.cpu().numpy(), ref_embed.cpu().numpy())
    assert cos >= COS_MIN and linf <= LINF_REL_MAX


def test_embedding_is_batch_invariant_at_full_batch(lib):
    """Property at the BASELINE batch size (256): an image's embedding does not depend on its batch or slot."""
    from irp_b200.stage import ResNet50Trunk
    big = ResNet50Trunk(stage_ref.full_resnet50(seed=1234), torch.device("cuda:0"), max_batch=256)
    g = torch.Generator(device="cuda").manual_seed(3)
    xp = torch.zeros(256, 230, 230, 4, device="cuda", dtype=torch.bfloat16)
    xp[:, 3:227, 3:227, :3] = torch.randn(256, 224, 224, 3, device="cuda", generator=g).bfloat16()
    full = big.embed(xp)
    perm = torch.randperm(256, device="cuda", generator=g)
    shuffled = big.embed(xp[perm].contiguous())
    assert torch.equal(shuffled, full[perm])
    part = big.embed(xp[100:117].contiguous())
    cos, linf = _embedding_metrics(part.cpu().numpy(), full[100:117].cpu().numpy())
    assert cos > 0.99999 and linf < 1e-2  # a different tile shape only changes the fp32 summation order
    assert torch.isfinite(full).all() and (full >= 0).all()
    big.close()


def test_model_callable_like_the_reference(lib):
    from functions import data_curation as dc
    model, transform = dc.initialize_model("cuda:0", weights=None, seed=1234, max_batch=8)
    ref_model, _ = stage_ref.initialize_model("cuda:0", seed=1234)
    x = torch.randn(3, 3, 224, 224, device="cuda")
    with torch.no_grad():
        out = model(x)
        ref = ref_model(x.bfloat16().float())
    assert out.shape == (3, 2048, 1, 1) and out.dtype == torch.float32
    cos, linf = _embedding_metrics(out.flatten(1).cpu().numpy(), ref.flatten(1).cpu().numpy())
    assert cos >= COS_MIN and linf <= LINF_REL_MAX
    assert "B200Transform" in repr(transform)


# =============================================================================================== A3 PCA
def _gpu_pca(ops_mod, x, k, solver=None, first_check=0, want_info=False):
    xt = torch.from_numpy(x).cuda()
    n, d = xt.shape
    shift = xt[: min(256, n)].mean(0).contiguous()
    acc = torch.zeros(1 + d + d * d, dtype=torch.float64, device="cuda")
    count, total, scatter = acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d)
    ops_mod.cov_accumulate(xt, shift, count, total, scatter)
    info = None
    if solver is None:
        mean, comps, evals = ops_mod.pca_fit(count, total, scatter, shift, k)
    else:
        mean, comps, evals, info = ops_mod.pca_fit_ex(count, total, scatter, shift, k, solver, first_check)
    z = ops_mod.pca_transform(xt, mean, comps)
    out = (mean.cpu().numpy(), comps.cpu().numpy(), evals.cpu().numpy(), z.cpu().numpy())
    return out + (info,) if want_info else out


def test_pca_matches_golden(ops_mod):
    g = load_golden("pca.npz")
    x = synth.embedding_like(int(g["n"]), int(g["d"]), seed=int(g["seed"]))
    k = int(g["k"])
    mean, comps, evals, z = _gpu_pca(ops_mod, x, k)
    assert pca_ref.subspace_angle(comps, g["components"]) <= ANGLE_MAX
    assert ((comps * g["components"]).sum(1) > 0.999).all()  # sign convention + per-component match
    np.testing.assert_allclose(evals[:k], g["explained_variance"], rtol=1e-4)
    np.testing.assert_allclose(evals[:k] / evals[k], g["explained_variance_ratio"], rtol=1e-4)
    np.testing.assert_allclose(mean, g["mean"], atol=1e-5)
    np.testing.assert_allclose(z[:64], g["z_head"], atol=2e-3 * np.abs(g["z_head"]).max())


@pytest.mark.parametrize("n,k", [(1500, 50), (300, 128), (64, 8), (2500, 1)])
def test_pca_matches_exact_oracle(ops_mod, n, k):
    x = synth.embedding_like(n, 2048, seed=n + k)
    mean, comps, evals, z = _gpu_pca(ops_mod, x, k)
    ref = pca_ref.pca_fit(x, k)
    assert pca_ref.subspace_angle(comps, ref.components) <= ANGLE_MAX
    assert np.abs(comps @ comps.T - np.eye(k)).max() < 1e-9
    np.testing.assert_allclose(evals[:k], ref.explained_variance, rtol=2e-4, atol=1e-6 * ref.explained_variance[0])
    np.testing.assert_allclose(evals[k], ref.eigenvalues.sum(), rtol=1e-5)
    zref = pca_ref.pca_transform(x, ref.mean, ref.components)
    if k > 1:
        # compare inside the subspace: rotate our scores onto the oracle's basis
        r = comps @ ref.components.T
        np.testing.assert_allclose(z.astype(np.float64) @ r, zref, atol=2e-3 * np.abs(zref).max())


def test_pca_partial_sums_combine_like_one_pass(ops_mod):
    """Sharded accumulation (the multi-GPU path on one device): G partial accumulators summed == one pass."""
    x = synth.embedding_like(2000, 2048, seed=9)
    xt = torch.from_numpy(x).cuda()
    d = 2048
    shift = xt[:256].mean(0).contiguous()
    one = torch.zeros(1 + d + d * d, dtype=torch.float64, device="cuda")
    ops_mod.cov_accumulate(xt, shift, one[:1], one[1:1 + d], one[1 + d:].view(d, d))
    total = torch.zeros_like(one)
    for lo, hi in [(0, 700), (700, 701), (701, 2000)]:  # ragged shards, including a single-row one
        part = torch.zeros_like(one)
        ops_mod.cov_accumulate(xt[lo:hi].contiguous(), shift, part[:1], part[1:1 + d], part[1 + d:].view(d, d))
        total += part
    assert total[0].item() == 2000.0
    np.testing.assert_allclose(total[1:1 + d].cpu().numpy(), one[1:1 + d].cpu().numpy(), rtol=1e-12, atol=1e-9)
    a = torch.triu(total[1 + d:].view(d, d)).cpu().numpy()
    b = torch.triu(one[1 + d:].view(d, d)).cpu().numpy()
    assert np.abs(a - b).max() <= 2e-6 * np.abs(b).max()


def test_pca_full_size_properties(ops_mod):
    """BASELINE size (27 000 x 2048, k=50): orthonormal components, eigen-residual, variance bookkeeping."""
    n, k = 27000, 50
    base = torch.from_numpy(synth.embedding_like(3000, 2048, seed=1)).cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    xt = (base.repeat(9, 1) + 0.3 * torch.randn(n, 2048, device="cuda", generator=g)).contiguous()
    d = 2048
    shift = xt[:256].mean(0).contiguous()
    acc = torch.zeros(1 + d + d * d, dtype=torch.float64, device="cuda")
    ops_mod.cov_accumulate(xt, shift, acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d))
    mean, comps, evals = ops_mod.pca_fit(acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d), shift, k)
    x64 = xt.double()
    mu = x64.mean(0)
    cov = (x64 - mu).T @ (x64 - mu) / (n - 1)
    assert torch.allclose(mean, mu, atol=1e-5)
    eye = torch.eye(k, dtype=torch.float64, device="cuda")
    assert (comps @ comps.T - eye).abs().max().item() < 1e-9
    resid = (cov @ comps.T - comps.T * evals[:k]).norm(dim=0) / evals[0]
    assert resid.max().item() < 1e-5
    ev = evals[:k]
    assert (ev[:-1] >= ev[1:]).all() and (ev >= 0).all()
    assert abs(evals[k].item() - torch.trace(cov).item()) / torch.trace(cov).item() < 1e-5
    z = ops_mod.pca_transform(xt, mean, comps)
    zref = (x64 - mu) @ comps.T
    assert (z.double() - zref).abs().max().item() < 1e-3 * zref.abs().max().item()
    np.testing.assert_allclose(z.double().var(0, unbiased=True).cpu().numpy(), ev.cpu().numpy(), rtol=1e-3)


def test_pca_lanczos_extends_until_converged_and_matches_eigh(ops_mod):
    """Power-law spectrum (lambda_i ~ 1/i, 2 % gaps around the 50th) with the first convergence check forced too
    early (56 steps for 50 pairs): the Krylov iteration must extend itself, and the result must still agree with a
    dense fp64 eigendecomposition; bitwise repeatable (every reduction on the path has a fixed order)."""
    from irp_b200 import _lib
    rng = np.random.default_rng(77)
    d, n, k = 2048, 6000, 50
    scale = (np.arange(1, d + 1, dtype=np.float64) ** -0.5).astype(np.float32)
    x = (rng.standard_normal((n, d), dtype=np.float32) * scale)[:, rng.permutation(d)].copy()
    mean, comps, evals, _, info = _gpu_pca(ops_mod, x, k, _lib.PCA_SOLVER_LANCZOS, 56, want_info=True)
    assert info["solver"] == _lib.PCA_SOLVER_LANCZOS and info["checks"] >= 2 and info["lanczos_steps"] > 56, info
    ref = pca_ref.pca_fit(x, k)
    assert pca_ref.subspace_angle(comps, ref.components) <= ANGLE_MAX
    assert np.abs(comps @ comps.T - np.eye(k)).max() < 1e-9
    np.testing.assert_allclose(evals[:k], ref.explained_variance, rtol=2e-4)
    for _ in range(3):
        mean2, comps2, evals2, _ = _gpu_pca(ops_mod, x, k, _lib.PCA_SOLVER_LANCZOS, 56)
        assert np.array_equal(mean, mean2) and np.array_equal(comps, comps2) and np.array_equal(evals, evals2)


def test_pca_lanczos_and_householder_agree(ops_mod):
    """The Krylov path and the exact Householder path (taken for small / rank-deficient problems) on one input."""
    from irp_b200 import _lib
    x = synth.embedding_like(1200, 2048, seed=3)
    _, ca, ea, _, ia = _gpu_pca(ops_mod, x, 50, _lib.PCA_SOLVER_LANCZOS, want_info=True)
    _, cb, eb, _, ib = _gpu_pca(ops_mod, x, 50, _lib.PCA_SOLVER_HOUSEHOLDER, want_info=True)
    assert ia["solver"] == _lib.PCA_SOLVER_LANCZOS and ib["solver"] == _lib.PCA_SOLVER_HOUSEHOLDER
    np.testing.assert_allclose(ea, eb, rtol=1e-10)
    assert (np.abs((ca * cb).sum(1)) > 1 - 1e-10).all()
    assert ((ca * cb).sum(1) > 0).all()  # same sign convention


def test_cov_accumulate_is_bitwise_repeatable_and_chunk_ordered(ops_mod):
    """The scatter matrix, the column sums and the trace are reduced in a fixed order (no floating-point atomics):
    ten accumulations of the same rows give bit-identical fp64 accumulators and eigenvalues."""
    x = torch.from_numpy(synth.embedding_like(5000, 2048, seed=21)).cuda()
    d = 2048
    shift = x[:256].mean(0).contiguous()
    ref_acc, ref_ev = None, None
    for _ in range(10):
        acc = torch.zeros(1 + d + d * d, dtype=torch.float64, device="cuda")
        ops_mod.cov_accumulate(x, shift, acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d))
        _, _, ev = ops_mod.pca_fit(acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d), shift, 50)
        if ref_acc is None:
            ref_acc, ref_ev = acc.clone(), ev.clone()
        assert torch.equal(acc, ref_acc) and torch.equal(ev, ref_ev)


# =============================================================================================== A4 scoring
def _flags_equal_outside_band(flags, ref_flags, ref_scores, ref_offset):
    near = np.abs(ref_scores - ref_offset) <= BAND * abs(ref_offset)
    return int(((flags != ref_flags) & ~near).sum())


def test_lof_matches_reference_golden(ops_mod):
    g = load_golden("lof.npz")
    z, y = synth.clustered_points(int(g["n"]), int(g["d"]), int(g["classes"]), seed=int(g["seed"]))
    zt = torch.from_numpy(z).cuda()
    scores, offs, flags = ops_mod.lof(zt, None, 1, 75, 0.03)
    np.testing.assert_allclose(scores.cpu().numpy(), g["global_scores"], rtol=2e-6)
    assert abs(offs[0].item() - float(g["global_offset"])) < 2e-6 * abs(float(g["global_offset"]))
    assert _flags_equal_outside_band(flags.cpu().numpy().astype(bool), g["global_outliers"], g["global_scores"],
                                     float(g["global_offset"])) == 0
    # per class: sklearn label encoding = sorted class names = ids here
    ids = torch.from_numpy(y.astype(np.int32)).cuda()
    _, _, cflags = ops_mod.lof(zt, ids, int(g["classes"]), 30, 0.05)
    bad = 0
    for c in range(int(g["classes"])):
        m = y == c
        ref_flags, near = stage_ref.lof_band(z[m], 30, 0.05, BAND)
        assert np.array_equal(ref_flags, g["class_outliers"][m])
        bad += int(((cflags.cpu().numpy()[m].astype(bool) != ref_flags) & ~near).sum())
    assert bad == 0


def test_detect_outliers_drop_in_matches_reference_golden(lib):
    from functions import data_curation as dc
    g = load_golden("lof.npz")
    z, y = synth.clustered_points(int(g["n"]), int(g["d"]), int(g["classes"]), seed=int(g["seed"]))
    labels = np.array([f"cls{c:02d}" for c in y])
    cls_out, glob_out = dc.detect_outliers(z, labels)
    assert cls_out.dtype == bool and glob_out.dtype == bool and cls_out.shape == (len(z),)
    # no score lies inside the band for this fixture, so the sets must be identical
    assert np.array_equal(glob_out, g["global_outliers"])
    assert np.array_equal(cls_out, g["class_outliers"])
    # the reference's production input: 2-D points, classes smaller than n_neighbors (k clipped with a warning)
    z2, y2 = synth.clustered_points(int(g["n2"]), int(g["d2"]), int(g["classes2"]), seed=int(g["seed2"]))
    labels2 = np.array([f"c{c:02d}" for c in y2])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        cls2, glob2 = dc.detect_outliers(z2, labels2)
    assert any("n_neighbors" in str(x.message) for x in w)
    assert np.array_equal(glob2, g["global_outliers2"])
    assert np.array_equal(cls2, g["class_outliers2"])


def test_lof_duplicates_and_unsorted_groups(ops_mod):
    z, y = synth.clustered_points(1200, 16, 5, seed=8)
    z[:40] = z[100]          # 40 exact duplicates (> k): lrd hits the 1e-10 guard
    y[:40] = y[100]
    zt = torch.from_numpy(z).cuda()
    ids = torch.from_numpy(y.astype(np.int32)).cuda()
    scores, _, flags = ops_mod.lof(zt, ids, 5, 30, 0.05)
    bad = 0
    for c in range(5):
        m = y == c
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_flags, near = stage_ref.lof_band(z[m], 30, 0.05, BAND)
        bad += int(((flags.cpu().numpy()[m].astype(bool) != ref_flags) & ~near).sum())
    assert bad == 0
    assert torch.isfinite(scores).all()


def test_lof_sharded_multi_two_problems_equal_the_single_calls(ops_mod):
    """ops.lof_sharded_multi (per-class + global problem in lockstep, their neighbour searches side by side on two
    streams) on one part with an identity all-reduce: bit-equal to two irp_lof calls, run twice for stream hazards."""
    z, y = synth.clustered_points(4000, 20, 5, seed=3)
    zt = torch.from_numpy(z).cuda()
    ids = torch.from_numpy(y.astype(np.int32)).cuda()
    problems = [(ids, 5, 30, 0.05), (None, 1, 75, 0.03)]
    refs = [ops_mod.lof(zt, g, ng, k, c) for (g, ng, k, c) in problems]
    for _ in range(2):
        res = ops_mod.lof_sharded_multi(zt, problems, 0, 1, lambda t: None)
        torch.cuda.synchronize()
        for (s, o, f), (rs, ro, rf) in zip(res, refs):
            assert torch.equal(s, rs) and torch.equal(o, ro) and torch.equal(f, rf)


@pytest.mark.parametrize("n_parts", [2, 3, 8])
def test_lof_sharded_over_parts_is_bitwise_the_single_call(lib, ops_mod, n_parts):
    """Multi-GPU form of irp_lof on ONE GPU: the parts (ranks) are run one after the other, their three fp64
    vectors summed exactly as the all-reduce would; scores / offsets / flags must equal irp_lof bit for bit."""
    from irp_b200 import _lib
    z, y = synth.clustered_points(3000, 24, 6, seed=11)
    zt = torch.from_numpy(z).cuda()
    for ids, n_groups, k, cont in ((torch.from_numpy(y.astype(np.int32)).cuda(), 6, 30, 0.05), (None, 1, 75, 0.03)):
        ref_scores, ref_offsets, ref_flags = ops_mod.lof(zt, ids, n_groups, k, cont)
        n, d = zt.shape
        ws_bytes = lib.irp_lof_workspace_bytes(n, d, k)
        wss = [torch.empty(ws_bytes, dtype=torch.uint8, device="cuda") for _ in range(n_parts)]
        vec = lambda: torch.zeros(n, dtype=torch.float64, device="cuda")
        total = vec()
        for p in range(n_parts):
            t = vec()
            _lib.check(lib.irp_lof_knn_part(_ptr(zt), n, d, _ptr(ids), n_groups, k, p, n_parts, _ptr(t), _ptr(wss[p]),
                                            ws_bytes, _stream()), "knn_part")
            total += t
        kdist = total
        total = vec()
        for p in range(n_parts):
            t = vec()
            _lib.check(lib.irp_lof_lrd_part(n, n_groups, k, p, n_parts, _ptr(kdist), _ptr(t), _ptr(wss[p]), ws_bytes,
                                            _stream()), "lrd_part")
            total += t
        lrd = total
        total = vec()
        for p in range(n_parts):
            t = vec()
            _lib.check(lib.irp_lof_score_part(n, n_groups, k, p, n_parts, _ptr(lrd), _ptr(t), _ptr(wss[p]), ws_bytes,
                                              _stream()), "score_part")
            total += t
        scores, offsets = vec(), torch.empty(n_groups, dtype=torch.float64, device="cuda")
        flags = torch.empty(n, dtype=torch.uint8, device="cuda")
        _lib.check(lib.irp_lof_finish(n, n_groups, k, C.c_double(cont), _ptr(total), _ptr(scores), _ptr(offsets),
                                      _ptr(flags), _ptr(wss[0]), ws_bytes, _stream()), "finish")
        torch.cuda.synchronize()
        assert torch.equal(scores, ref_scores) and torch.equal(offsets, ref_offsets) and torch.equal(flags, ref_flags)


def test_lof_full_size_properties(ops_mod):
    """BASELINE size (27 000 x 50): flag counts follow the contamination, flags are permutation-equivariant, and
    the per-class pass equals running each class alone."""
    n, d, G = 27000, 50, 10
    rng = np.random.default_rng(0)
    y = synth.class_assignment(n, seed=0)
    centers = rng.standard_normal((G, d)) * 4
    z = (centers[y] + rng.standard_normal((n, d)) * rng.uniform(0.5, 2.0, (n, 1))).astype(np.float32)
    zt, ids = torch.from_numpy(z).cuda(), torch.from_numpy(y).cuda()
    gs, goff, gf = ops_mod.lof(zt, None, 1, 75, 0.03)
    cs, coff, cf = ops_mod.lof(zt, ids, G, 30, 0.05)
    assert abs(int(gf.sum()) - 0.03 * n) <= 2
    for c in range(G):
        m = ids == c
        assert abs(int(cf[m].sum()) - 0.05 * int(m.sum())) <= 2
        solo_s, solo_o, solo_f = ops_mod.lof(zt[m].contiguous(), None, 1, 30, 0.05)
        assert torch.equal(solo_f, cf[m]) and torch.allclose(solo_s, cs[m], rtol=1e-12, atol=0)
        assert abs(solo_o[0].item() - coff[c].item()) < 1e-12
    perm = torch.randperm(n, device="cuda")
    ps, _, pf = ops_mod.lof(zt[perm].contiguous(), None, 1, 75, 0.03)
    assert torch.allclose(ps, gs[perm], rtol=1e-9, atol=0)
    assert int((pf != gf[perm]).sum()) <= 1  # only a tie exactly at the threshold could flip
    assert (gs <= 0).all() and gs.max().item() < -0.5


def test_centroid_scorer_matches_its_oracle(ops_mod):
    z, y = synth.clustered_points(5000, 50, 10, seed=4)
    zt, ids = torch.from_numpy(z).cuda(), torch.from_numpy(y.astype(np.int32)).cuda()
    dist, zs, thr, flags = ops_mod.centroid_zscore(zt, ids, 10, 0.05)
    rd, rz, rt, rf = lof_ref.centroid_zscore(z, y, 10, 0.05)
    np.testing.assert_allclose(dist.cpu().numpy(), rd, rtol=1e-12)
    np.testing.assert_allclose(zs.cpu().numpy(), rz, atol=1e-10)
    np.testing.assert_allclose(thr.cpu().numpy(), rt, rtol=1e-12)
    assert np.array_equal(flags.cpu().numpy().astype(bool), rf)
    dist1, _, thr1, flags1 = ops_mod.centroid_zscore(zt, None, 1, 0.03)
    rd1, _, rt1, rf1 = lof_ref.centroid_zscore(z, None, 1, 0.03)
    np.testing.assert_allclose(dist1.cpu().numpy(), rd1, rtol=1e-12)
    assert np.array_equal(flags1.cpu().numpy().astype(bool), rf1)


# =============================================================================================== drop-in stage
def test_process_image_directory_drop_in(lib, tmp_path, capsys):
    from PIL import Image
    from functions import data_curation as dc
    g, images = _golden_images()
    for name, img in zip(g["names"], images):
        p = tmp_path / str(name)
        p.parent.mkdir(exist_ok=True)
        Image.fromarray(img).save(p)
    (tmp_path / "class0" / "broken.png").write_bytes(b"not an image")
    (tmp_path / "README.txt").write_text("stray file")
    model, transform = dc.initialize_model("cuda:0", weights=None, seed=int(g["seed"]), max_batch=8)
    feats, labels, paths = dc.process_image_directory(str(tmp_path), "cuda:0", transform, batch_size=5, model=model)
    assert "Skipped" in capsys.readouterr().out          # reference prints and continues (:681-682)
    assert feats.shape == (12, 2048) and feats.dtype == np.float32
    assert labels.shape == (12,) and paths.shape == (12,)
    rel = [os.path.relpath(p, tmp_path) for p in paths]
    order = [rel.index(str(n)) for n in g["names"]]
    assert list(labels[order]) == list(g["labels"])
    cos, linf = _embedding_metrics(feats[order], g["features"])
    assert cos >= COS_MIN and linf <= LINF_REL_MAX
    # a host-side transform (not the fused one) takes the per-image path and still matches
    _, ref_tfm = stage_ref.initialize_model("cpu")
    feats2, _, paths2 = dc.process_image_directory(str(tmp_path), "cuda:0", ref_tfm, batch_size=4, model=model)
    rel2 = [os.path.relpath(p, tmp_path) for p in paths2]
    cos, linf = _embedding_metrics(feats2[[rel2.index(str(n)) for n in g["names"]]], g["features"])
    assert cos >= COS_MIN and linf <= LINF_REL_MAX


def test_create_pca_embeddings_returns_a_working_sklearn_pca(lib):
    from functions import data_curation as dc
    x = synth.embedding_like(800, 2048, seed=21)
    labels = np.array([f"c{i % 7}" for i in range(800)])
    z, le, pca = dc.create_pca_embeddings(x, labels, pca_components=50)
    zr, ref = stage_ref.pca_exact(x, 50)
    assert z.shape == (800, 50) and z.dtype == np.float32 and list(le.classes_) == sorted(set(labels))
    assert pca_ref.subspace_angle(pca.components_, ref.components_) <= ANGLE_MAX
    np.testing.assert_allclose(pca.explained_variance_ratio_, ref.explained_variance_ratio_, rtol=1e-3)
    np.testing.assert_allclose(pca.singular_values_, ref.singular_values_, rtol=1e-3)
    np.testing.assert_allclose(pca.noise_variance_, ref.noise_variance_, rtol=1e-3)
    np.testing.assert_allclose(pca.transform(x[:10]), z[:10], atol=1e-2)  # sklearn's own transform on our fit
    with pytest.raises(ValueError):
        dc.create_pca_embeddings(x[:20], labels[:20], pca_components=50)
    with pytest.raises(ImportError):
        dc.create_embeddings(x, labels)  # umap-learn is absent here; the reference fails at import time


# ------------------------------------------------------------------ N4 UMAP graph construction (parity unpinned)
@pytest.mark.parametrize("n,d,k", [(600, 20, 15), (3000, 50, 15), (257, 8, 30), (40, 3, 5)])
def test_knn_graph_and_fuzzy_weights_match_oracle(lib, n, d, k):
    """irp_knn_graph / irp_umap_fuzzy_weights against oracle/umap_graph_ref.py (umap-learn 0.5.7 restated; parity
    unpinned).  Neighbours: distances within 1e-6 relative, index rows equal wherever a row's distances are all
    distinct.  sigma within 1e-3 relative (the bisection stops on |psum - log2 k| < 1e-5, reached a step earlier or
    later depending on the last bits of expf), rho exact, edge weights within 1e-4."""
    from irp_b200 import ops
    from oracle import umap_graph_ref as ug
    rng = np.random.default_rng(n + d)
    x = rng.normal(size=(n, d)).astype(np.float32) * (1.0 + 3.0 * rng.random((n, 1)).astype(np.float32))
    x[3] = x[4]  # exact duplicates: zero distances next to the row itself
    idx_r, dist_r = ug.nearest_neighbors(x, k)
    xt = torch.from_numpy(x).cuda()
    idx_t, dist_t = ops.knn_graph(xt, k)
    idx, dist = idx_t.cpu().numpy(), dist_t.cpu().numpy()
    assert np.array_equal(idx[:, 0], np.arange(n)) and np.all(dist[:, 0] == 0)
    assert np.abs(dist - dist_r).max() <= 1e-6 * max(1.0, dist_r.max())
    # rows 3 and 4 are the same point: as neighbours they are interchangeable (when only one of them fits below the
    # k-th distance, which one is kept is a tie) -> compare with 4 renamed to 3
    canon = lambda a: np.where(a == 4, 3, a)  # noqa: E731
    distinct = np.array([np.unique(r).size == k for r in dist_r])
    assert np.array_equal(canon(idx[distinct]), canon(idx_r[distinct]))
    for i in np.flatnonzero(~distinct):
        assert set(canon(idx[i]).tolist()) == set(canon(idx_r[i]).tolist())
    assert np.all(np.diff(dist, axis=1) >= 0)
    # weights on the ORACLE's arrays (stage-wise parity)
    sig_r, rho_r = ug.smooth_knn_dist(dist_r, float(k))
    _, _, vals_r = ug.compute_membership_strengths(idx_r, dist_r, sig_r, rho_r)
    sig, rho, vals = ops.umap_fuzzy_weights(torch.from_numpy(idx_r).cuda(), torch.from_numpy(dist_r).cuda())
    assert np.array_equal(rho.cpu().numpy(), rho_r)
    assert np.abs(sig.cpu().numpy() - sig_r).max() <= 1e-3 * sig_r.max()
    assert np.abs(vals.cpu().numpy().reshape(-1) - vals_r).max() <= 1e-4
    # the defining equation on the device result itself
    s = sig.cpu().numpy().astype(np.float64)
    psum = np.exp(-np.maximum(dist_r[:, 1:].astype(np.float64) - rho_r[:, None], 0) / s[:, None]).sum(1)
    assert np.abs(psum - np.log2(k)).max() < 1e-3


def test_fuzzy_simplicial_set_mirror_matches_oracle_graph(lib):
    from irp_b200 import umap_graph
    from oracle import umap_graph_ref as ug
    rng = np.random.default_rng(11)
    x = rng.normal(size=(900, 16)).astype(np.float32)
    g, sig, rho = umap_graph.fuzzy_simplicial_set(x, 15)
    gr, sig_r, rho_r = ug.fuzzy_simplicial_set(x, 15)
    assert abs(g.tocsr() - g.tocsr().T).max() == 0
    assert abs(g.tocsr() - gr.tocsr()).max() <= 2e-4 and np.array_equal(rho, rho_r)
    idx, dist = umap_graph.nearest_neighbors(x, 15)
    assert idx.dtype == np.int32 and dist.dtype == np.float32 and idx.shape == (900, 15)


def test_whole_stage_against_reference_route(trunk):
    """Config-1-shaped run (reduced to 160 images): every stage fed the oracle's previous-stage output."""
    from irp_b200.stage import OutlierStage, pack_images
    images, labels = synth.config1_images(160, 10, seed=0)
    ids = np.array([int(l[5:]) for l in labels], np.int32)
    stage = OutlierStage(trunk, batch_size=64, pca_components=20)
    res = stage.run(pack_images(images), torch.from_numpy(ids), 10, from_host=True)
    ref_feats = stage_ref.embed_arrays(images, batch_size=32, seed=1234)
    cos, linf = _embedding_metrics(res.features.cpu().numpy(), ref_feats)
    assert cos >= COS_MIN and linf <= LINF_REL_MAX
    # PCA on the ORACLE features (stage-wise parity)
    xt = torch.from_numpy(ref_feats).cuda()
    pca = stage.fit_pca(xt)
    zr, ref = stage_ref.pca_exact(ref_feats, 20)
    assert pca_ref.subspace_angle(pca.components.cpu().numpy(), ref.components_) <= ANGLE_MAX
    # scoring on the ORACLE projection
    zt = torch.from_numpy(zr.astype(np.float32)).cuda()
    cf, gf, cs, gs = stage.detect(zt, torch.from_numpy(ids).cuda(), 10)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gref, gnear = stage_ref.lof_band(zr.astype(np.float32), 75, 0.03, BAND)
        assert int(((gf.cpu().numpy() != gref) & ~gnear).sum()) == 0
        for c in range(10):
            m = ids == c
            cref, cnear = stage_ref.lof_band(zr.astype(np.float32)[m], 30, 0.05, BAND)
            assert int(((cf.cpu().numpy()[m] != cref) & ~cnear).sum()) == 0
    assert res.class_outliers.shape == (160,) and res.z.shape == (160, 20)


def test_scale_out_config_shape_batch512_pca128_global_only(lib):
    """BASELINE.json configs[3] scaled down (1M -> 2 560 images): 224x224 inputs, batch 512, PCA(128), global
    scorer only.  Embeddings of a sample against the reference route, PCA(128) and the global LOF flags against the
    oracle on the same features."""
    from irp_b200.stage import OutlierStage, ResNet50Trunk, pack_images
    n, k = 2560, 128
    images, _ = synth.config1_images(n, 10, seed=3)
    trunk512 = ResNet50Trunk(stage_ref.full_resnet50(seed=1234), torch.device("cuda:0"), max_batch=512)
    try:
        stage = OutlierStage(trunk512, batch_size=512, pca_components=k)
        assert stage.batch_size == 512
        feats = stage.embed_packed(pack_images(images), from_host=True)
        sample = list(range(0, n, 40))  # 64 images spread over all five batches
        ref_feats = stage_ref.embed_arrays([images[i] for i in sample], batch_size=32, seed=1234)
        cos, linf = _embedding_metrics(feats[sample].cpu().numpy(), ref_feats)
        assert cos >= COS_MIN and linf <= LINF_REL_MAX
        pca = stage.fit_pca(feats)
        z = stage.transform(feats, pca)
        x = feats.cpu().numpy()
        ref = pca_ref.pca_fit(x, k)
        assert pca_ref.subspace_angle(pca.components.cpu().numpy(), ref.components) <= ANGLE_MAX
        zref = pca_ref.pca_transform(x, ref.mean, ref.components).astype(np.float32)
        r = pca.components.cpu().numpy() @ ref.components.T
        np.testing.assert_allclose(z.cpu().numpy().astype(np.float64) @ r, zref, atol=2e-3 * np.abs(zref).max())
        # global scorer on the ORACLE projection (stage-wise parity)
        zt = torch.from_numpy(zref).cuda()
        gs, goff, gf = stage.backend.lof(zt, None, 1, 75, 0.03)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gref, gnear = stage_ref.lof_band(zref, 75, 0.03, BAND)
        assert int(((gf.bool().cpu().numpy() != gref) & ~gnear).sum()) == 0
        assert abs(int(gf.sum()) - round(0.03 * n)) <= 2
    finally:
        trunk512.close()


# =============================================================================================== N1 classifier
def _val_preprocess(ops_mod, images, layout):
    from irp_b200 import _lib
    from irp_b200.stage import pack_images
    p = pack_images(images, transform=_lib.TRANSFORM_VAL_256).to("cuda:0")
    return ops_mod.preprocess_ex(p.pixels, p.offsets, p.hw, p.max_taps, layout, _lib.TRANSFORM_VAL_256)


def test_val_transform_golden_and_ragged_sizes_bit_exact(ops_mod):
    """functions/dataload.py:51-56 (Resize((256,256)), CenterCrop(224), ToTensor, Normalize) on the device: equal to
    the reference's float32 output rounded to bf16, for the golden images and for assorted aspect ratios
    (upscales, 4x downscales, already-256 inputs)."""
    from oracle import classifier_ref
    g = load_golden("classifier.npz")
    images, _ = classifier_ref.synthetic_eval_set(int(g["n"]), int(g["num_classes"]), seed=int(g["data_seed"]))
    out = _val_preprocess(ops_mod, images[:4], 0).cpu()
    assert torch.equal(out, torch.from_numpy(g["x_head"]).bfloat16())
    sizes = [(256, 256), (300, 400), (97, 640), (1024, 300), (224, 224), (31, 29), (600, 800)]
    imgs = [np.random.default_rng(50 + i).integers(0, 256, (h, w, 3), dtype=np.uint8) for i, (h, w) in enumerate(sizes)]
    want = torch.from_numpy(np.stack([pil_resample.val_transform(im) for im in imgs])).bfloat16()
    assert torch.equal(_val_preprocess(ops_mod, imgs, 0).cpu(), want)
    padded = _val_preprocess(ops_mod, imgs, 1).cpu()
    assert torch.equal(padded[:, 3:227, 3:227, :3], want.permute(0, 2, 3, 1))
    assert padded[:, :3].abs().sum() == 0 and padded[:, :, :3].abs().sum() == 0 and padded[..., 3].abs().sum() == 0


@pytest.mark.parametrize("B,C", [(1, 10), (37, 10), (256, 10), (300, 3), (64, 37)])
def test_classifier_head_matches_torch(ops_mod, B, C):
    g = torch.Generator(device="cuda").manual_seed(B * 100 + C)
    feats = torch.randn(B, 2048, device="cuda", generator=g).abs() * 3
    w1 = torch.randn(512, 2048, device="cuda", generator=g) / 45
    b1 = torch.randn(512, device="cuda", generator=g)
    w2 = torch.randn(C, 512, device="cuda", generator=g) / 22
    b2 = torch.randn(C, device="cuda", generator=g)
    logits, pred = ops_mod.classifier_head(feats, w1, b1, w2, b2)
    ref = torch.relu(feats.double() @ w1.double().t() + b1.double()) @ w2.double().t() + b2.double()
    assert (logits.double() - ref).abs().max().item() < 1e-4 * ref.abs().max().item()
    assert torch.equal(pred.long(), logits.argmax(1))
    labels = torch.randint(0, C, (B,), device="cuda", generator=g)
    labels[0] = C - 1
    weights = torch.rand(C, device="cuda", generator=g) + 0.5
    for w in (None, weights):
        stats = ops_mod.cross_entropy_stats(logits, labels, w).cpu().numpy()
        ce = F.cross_entropy(logits.double(), labels, weight=None if w is None else w.double(), reduction="mean")
        assert abs(stats[0] / stats[1] - ce.item()) < 1e-9 * max(1.0, abs(ce.item()))
        assert stats[2] == float((logits.argmax(1) == labels).sum().item())


def test_classifier_matches_reference_golden_and_evaluate_full_drop_in(lib, capsys):
    """AnimalClassifier inference + evaluate_full (functions/model.py:9-41, functions/train.py:192-238) through the
    drop-in against the outputs of the reference's own code (tests/golden/classifier.npz)."""
    from functions import train as b200_train
    from irp_b200 import _lib
    from irp_b200.classifier import B200Classifier
    from irp_b200.stage import pack_images
    from oracle import classifier_ref
    g = load_golden("classifier.npz")
    n, c, b = int(g["n"]), int(g["num_classes"]), int(g["batch"])
    ref_model = classifier_ref.build_classifier(c, seed=int(g["seed"]))
    images, labels = classifier_ref.synthetic_eval_set(n, c, seed=int(g["data_seed"]))
    model = B200Classifier(ref_model, "cuda:0", max_batch=32)
    try:
        # 1. the reference's calling convention: normalised float tensors from the DataLoader
        batches = classifier_ref.val_batches(images, labels, b)
        logits = torch.cat([model(x) for x, _ in batches]).cpu().numpy()
        ref = g["logits"]
        scale = np.abs(ref).max()
        assert np.abs(logits - ref).max() <= 2e-2 * scale
        cos = (logits * ref).sum(1) / np.linalg.norm(logits, axis=1) / np.linalg.norm(ref, axis=1)
        assert cos.min() >= 0.999
        top2 = np.sort(ref, 1)[:, -2:]
        decided = (top2[:, 1] - top2[:, 0]) > 4e-2 * scale  # rows whose argmax cannot flip inside the tolerance
        assert np.array_equal(logits.argmax(1)[decided], g["preds"][decided])
        loss, acc, preds, labs = b200_train.evaluate_full(model, batches, torch.nn.CrossEntropyLoss(),
                                                          disable_progress=True)
        assert "Evaluated on 48 samples" in capsys.readouterr().out
        assert isinstance(preds, list) and len(preds) == n and np.array_equal(np.asarray(labs), g["labels"])
        assert abs(loss - float(g["loss"])) <= 2e-2 * max(1.0, float(g["loss"]))
        assert np.array_equal(np.asarray(preds)[decided], g["preds"][decided])
        assert abs(acc - float(g["acc"])) <= 100.0 * (~decided).sum() / n + 1e-9
        # 2. the fused route: decoded uint8 images, val_transform on the device
        lg2, pr2 = model.predict_packed(pack_images(images, transform=_lib.TRANSFORM_VAL_256))
        assert np.abs(lg2.cpu().numpy() - logits).max() <= 1e-5 * scale  # same pixels, same kernels
        assert torch.equal(pr2.long().cpu(), torch.from_numpy(logits.argmax(1)))
    finally:
        model.close()


# =============================================================================================== N2 Lanczos resize
def test_wds_lanczos_resize_bit_exact(ops_mod, lib):
    """resize_and_crop_image (functions/data_curation.py:883-913; Pillow LANCZOS, support 3): the device output is
    the reference's image byte for byte -- golden fixtures, the oracle on further sizes (up to a 2.7x downscale on
    the two-pass path, 4.5x through the band kernel), and the PIL-in / PIL-out drop-in with an RGBA input."""
    from irp_b200 import _lib
    from irp_b200.stage import pack_images
    from oracle.make_golden import wds_input
    g = load_golden("wds_resize.npz")
    imgs = [wds_input(int(s), int(h), int(w), smooth=(i % 2 == 0))
            for i, ((h, w), s) in enumerate(zip(g["sizes"], g["seeds"]))]

    def run(images):
        p = pack_images(images, transform=_lib.TRANSFORM_WDS_LANCZOS).to("cuda:0")
        return ops_mod.preprocess_ex(p.pixels, p.offsets, p.hw, p.max_taps, _lib.LAYOUT_U8_HWC,
                                     _lib.TRANSFORM_WDS_LANCZOS).cpu().numpy()

    out = run(imgs)
    for i in range(len(imgs)):
        assert np.array_equal(out[i], g["crops"][i]), tuple(g["sizes"][i])
    sizes = [(260, 333), (90, 120), (224, 230), (512, 380), (600, 601), (1010, 1400), (31, 500)]
    more = [np.random.default_rng(70 + i).integers(0, 256, (h, w, 3), dtype=np.uint8) for i, (h, w) in enumerate(sizes)]
    out = run(more)
    for i, im in enumerate(more):
        assert np.array_equal(out[i], pil_resample.wds_transform_u8(im)), sizes[i]
    # drop-in: PIL image in, PIL image out, RGBA composited on white like the reference
    from PIL import Image
    from functions import data_curation as dc
    rgba = np.random.default_rng(5).integers(0, 256, (180, 240, 4), dtype=np.uint8)
    pil = Image.fromarray(rgba, "RGBA")
    got = dc.resize_and_crop_image(pil)
    bg = Image.new("RGB", pil.size, (255, 255, 255))
    bg.paste(pil, mask=pil.split()[3])
    assert got.size == (224, 224) and got.mode == "RGB"
    assert np.array_equal(np.asarray(got), pil_resample.wds_transform_u8(np.asarray(bg)))


def test_bilinear_many_tap_images_bit_exact(ops_mod):
    """Downscales by 2.5x-13x (7 to 27 horizontal / vertical taps): the fused kernel's generic horizontal loop and
    its narrower bands (16 ... 2 output rows per item) -- all bit-exact with the reference transform."""
    sizes = [(600, 640), (650, 900), (700, 700), (1200, 1210), (1700, 1650), (233, 3000), (2600, 2500), (3000, 240)]
    imgs = [np.random.default_rng(90 + i).integers(0, 256, (h, w, 3), dtype=np.uint8) for i, (h, w) in enumerate(sizes)]
    assert torch.equal(_preprocess(ops_mod, imgs, 0).cpu(), _expected_bf16(imgs))
    assert torch.equal(_preprocess(ops_mod, imgs, 1).cpu()[:, 3:227, 3:227, :3], _expected_bf16(imgs).permute(0, 2, 3, 1))


# =============================================================================================== N3 duplicate hash
def test_image_hash_matches_reference_golden(ops_mod, lib):
    """compute_image_hash on the device (bicubic 64x64 resize + MD5) against the hex digests the UNMODIFIED reference
    function produced (tests/golden/hash.npz), through the batched op and through the PIL drop-in; then sizes that
    need the narrow bands / the band-kernel fallback (downscales up to 45x), against the oracle restatement."""
    from PIL import Image
    from functions import data_curation as dc
    from irp_b200 import _lib
    from irp_b200.stage import pack_images
    from oracle.make_golden import wds_input
    g = load_golden("hash.npz")
    imgs = [wds_input(int(s), int(h), int(w), smooth=(i % 2 == 0))
            for i, ((h, w), s) in enumerate(zip(g["sizes"], g["seeds"]))]
    part = pack_images(imgs, transform=_lib.TRANSFORM_HASH_64).to("cuda")
    small = ops_mod.preprocess_ex(part.pixels, part.offsets, part.hw, part.max_taps, _lib.LAYOUT_U8_HWC,
                                  _lib.TRANSFORM_HASH_64).cpu().numpy()
    for img, got in zip(imgs, small):
        assert np.array_equal(got, pil_resample.hash_resize_u8(img)), img.shape
    digests = ops_mod.image_hashes(part.pixels, part.offsets, part.hw, part.max_taps).cpu().numpy()
    assert [d.tobytes().hex() for d in digests] == [str(x) for x in g["hexdigests"]]
    pil = [Image.fromarray(im) for im in imgs]
    assert dc.compute_image_hashes(pil) == [str(x) for x in g["hexdigests"]]
    near = imgs[4].copy()
    near[150, 200, 1] ^= 0x40
    assert dc.compute_image_hash(Image.fromarray(near)) == str(g["near_hex"])
    # grayscale / RGBA: resized in their own mode on the host like the reference, digest on the device
    import hashlib
    for mode in ("L", "RGBA"):
        im = Image.fromarray(imgs[3]).convert(mode)
        want = hashlib.md5(im.copy().resize((64, 64)).convert("RGB").tobytes()).hexdigest()
        assert dc.compute_image_hash(im) == want
    big = [np.random.default_rng(70 + i).integers(0, 256, (h, w, 3), dtype=np.uint8)
           for i, (h, w) in enumerate([(900, 64), (64, 1500), (1400, 1900), (2900, 2400)])]
    bp = pack_images(big, transform=_lib.TRANSFORM_HASH_64).to("cuda")
    dg = ops_mod.image_hashes(bp.pixels, bp.offsets, bp.hw, bp.max_taps).cpu().numpy()
    assert [d.tobytes().hex() for d in dg] == [pil_resample.image_hash(b) for b in big]


def test_md5_rows_matches_hashlib(ops_mod):
    """RFC 1321 on the device: lengths around the 56 / 64-byte padding boundaries, an empty row, the hash's 12 288."""
    import hashlib
    rng = np.random.default_rng(3)
    for nbytes in (0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 128, 1000, 12288):
        rows = rng.integers(0, 256, (5, nbytes), dtype=np.uint8)
        if nbytes == 0:
            got = ops_mod.md5_rows(torch.zeros((5, 0), dtype=torch.uint8, device="cuda")).cpu().numpy()
        else:
            got = ops_mod.md5_rows(torch.from_numpy(rows).cuda()).cpu().numpy()
        assert [g.tobytes().hex() for g in got] == [hashlib.md5(r.tobytes()).hexdigest() for r in rows], nbytes
