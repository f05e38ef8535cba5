"""Developer tool: PCA fit, Lanczos vs Householder (subprocess per mode: the switch is read once per process)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    sys.path.insert(0, os.path.join(ROOT, "image-recognition-pipeline_b200")); sys.path.insert(0, ROOT)
    from irp_b200 import ops
    from oracle import synth
    out = sys.argv[2]
    d = 2048
    res = {}
    for case, (n, k) in enumerate([(27000, 50), (27000, 128), (4000, 50), (1500, 50)]):
        x = torch.from_numpy(synth.embedding_like(min(n, 3000), d, seed=1 + case)).cuda()
        if n > 3000:
            x = x.repeat(n // 3000, 1).contiguous()
            x += 0.3 * torch.randn(x.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
        shift = x[:256].mean(0).contiguous()
        acc = torch.zeros(1 + d + d * d, dtype=torch.float64, device="cuda")
        ops.cov_accumulate(x, shift, acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d))
        def fit():
            return ops.pca_fit(acc[:1], acc[1:1 + d], acc[1 + d:].view(d, d), shift, k)
        for _ in range(2): fit()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): mean, comps, ev = fit()
        e1.record(); torch.cuda.synchronize()
        print(f"case n={n} k={k}: fit {e0.elapsed_time(e1) / 5:.2f} ms", flush=True)
        res[f"c{case}"] = comps.cpu().numpy(); res[f"e{case}"] = ev.cpu().numpy()
    np.savez(out, **res)
else:
    import numpy as np
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    outs = {}
    for mode in ("lanczos", "householder"):
        env = dict(os.environ)
        if mode == "householder": env["IRP_PCA_HOUSEHOLDER"] = "1"
        path = os.path.join(ROOT, "gpurun_out", f"pca_{mode}.npz")
        print("==", mode, flush=True)
        subprocess.run([sys.executable, __file__, "child", path], env=env, check=True)
        outs[mode] = np.load(path)
    a, b = outs["lanczos"], outs["householder"]
    for case in range(4):
        ca, cb = a[f"c{case}"], b[f"c{case}"]
        ea, eb = a[f"e{case}"], b[f"e{case}"]
        s = np.linalg.svd(ca @ cb.T, compute_uv=False)
        ang = float(np.arccos(np.clip(s.min(), -1, 1)))
        dots = np.abs(np.sum(ca * cb, axis=1))
        print(f"case {case}: max rel eval diff {np.max(np.abs(ea - eb) / np.abs(eb).max()):.3e}, "
              f"subspace angle {ang:.3e}, min |<u_i,v_i>| {dots.min():.12f}, "
              f"signs equal {bool(np.all(np.sum(ca * cb, axis=1) > 0))}, "
              f"orth err {np.abs(ca @ ca.T - np.eye(len(ca))).max():.2e}")
