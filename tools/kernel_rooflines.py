#!/usr/bin/env python
"""Per-kernel achieved bandwidth / FLOP rate of the hot path, each kernel timed alone with CUDA events (3 warm-ups,
inputs larger than the 126 MB L2 so every timed launch streams from HBM), against MEASURED_PEAKS.json.
Algorithmic bytes per ray-level follow SURVEY.md §8d / DESIGN.md §4.   Usage: python tools/kernel_rooflines.py [B]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from durf_b200 import _lib as L, ops, synthetic as S  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    N = 128
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else dict(hbm_gbs=6650.0, bf16_tflops=1590.0)
    hbm = peaks["hbm_gbs"]
    rng = np.random.default_rng(S.SEED)
    rays, c2w = S.random_rays(rng, B, far=40.0)
    dev = "cuda"
    o, d, r = [torch.from_numpy(np.asarray(a)).to(dev) for a in (rays.origins, rays.directions, rays.radii)]
    near, far = torch.from_numpy(rays.near).to(dev), torch.from_numpy(rays.far).to(dev)
    out = []

    def report(name, ms, bytes_per_ray, note=""):
        gbs = B * bytes_per_ray / (ms * 1e-3) / 1e9
        out.append(dict(kernel=name, ms=round(ms, 4), rays=B, algorithmic_bytes_per_ray=bytes_per_ray, achieved_gbs=round(gbs, 1),
                        frac_of_measured_hbm=round(gbs / hbm, 3), note=note))
        print(f"{name:38s} {ms:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / hbm:5.1f}% of measured HBM ({hbm:.0f} GB/s)  {note}")

    t = ops.raymarch(o, d, r, N, near=near, far=far, contract=True, bf16_tiles=True)['t_vals']
    report("raymarch_fwd (bf16 tiles, contract)", timed(lambda: ops.raymarch(o, d, r, N, t_vals=t, contract=True, bf16_tiles=True)),
           48 + 516 + 128 * 128, "48 B ray + 516 B t_vals in, 16 KB tile image out")
    Bs = B // 4
    report("raymarch_fwd (fp32 features, parity)", timed(lambda: ops.raymarch(o[:Bs], d[:Bs], r[:Bs], N, t_vals=t[:Bs], contract=True)) * 4,
           48 + 516 + 128 * 60 * 4, "exact reference sequence (accurate sinf per feature): instruction-bound by design")
    raw_rgb = torch.randn(B, N, 3, device=dev)
    raw_den = torch.randn(B, N, device=dev)
    report("composite_fwd", timed(lambda: ops.composite(raw_rgb, raw_den, t, d)), 128 * 16 + 516 + 12 + 512 + 20 + 1024,
           "raw rgb/density + t_vals in; weights, t_mids, t_dists, rgb/depth/acc out")
    comp = ops.composite(raw_rgb, raw_den, t, d)
    g_rgb, g_dep, g_w = torch.randn(B, 3, device=dev), torch.randn(B, device=dev), torch.randn(B, N, device=dev)
    report("composite_bwd", timed(lambda: ops.composite_bwd(raw_rgb, raw_den, t, d, g_rgb, g_dep, g_w)), 128 * 16 + 516 + 12 + 512 + 16 + 128 * 16,
           "forward inputs + d_weights in, d_raw_rgb/d_raw_density out")
    w = comp['weights']
    report("resample (deterministic)", timed(lambda: ops.resample(t, w)), 1544)
    u = torch.rand(B, N + 1, device=dev)
    report("resample (randomized)", timed(lambda: ops.resample(t, w, u_rand=u)), 1544 + 516)
    K = 2
    centers, ext = S.boxes_in_view(rng, c2w, K)
    box = torch.from_numpy(centers[0]).to(dev)
    ext_t = torch.from_numpy(ext).to(dev)
    report("obb_frontend_fwd (K=2)", timed(lambda: ops.obb_frontend(o, d, box, ext_t)), 24 + 28 + 12 * K)
    vd = torch.from_numpy(rays.viewdirs).to(dev)
    report("viewdir_enc", timed(lambda: ops.viewdir_enc(vd, 4)), 12 + 27 * 4)
    n = 594308 + 2 * 168836 + 60
    p = torch.randn(n, device=dev); g = torch.randn(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
    ss = torch.zeros(1, device=dev)
    ms = timed(lambda: (ops.grad_sanitize(g, 0.1, 1.0, ss), ops.adam_step(p, g, m, v, ss, max_norm=1.0, lr=1e-3, step=0)))
    print(f"{'grad_sanitize + adam (0.93 M params)':38s} {ms:8.3f} ms  (launch-latency bound: 33 MB of traffic)")
    out.append(dict(kernel="grad_sanitize+adam", ms=round(ms, 4), params=n))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kernel_rooflines.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
