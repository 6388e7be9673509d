"""Shared fixtures: one synthetic scene, materialised for the oracle (torch CPU) and for the CUDA path."""
import numpy as np
import torch

from durf_b200 import synthetic as S
from oracle import durf_oracle as O


def scene(B=256, K=2, seed=7, bias_scale=0.05, far=40.0, behind=False, N=128):
    rng = np.random.default_rng(seed)
    rays, c2w = S.random_rays(rng, B, far=far)
    centers, ext = S.boxes_in_view(rng, c2w, K, behind=behind)
    mlp = S.glorot_mlp(rng, 60, 256, bias_scale)
    box_mlps = [S.glorot_mlp(rng, 63, 128, bias_scale) for _ in range(K)]
    tg = S.targets(rng, B)
    t_rand = rng.uniform(size=(B, N + 1)).astype(np.float32)
    u_rand = rng.uniform(size=(B, N + 1)).astype(np.float32)
    return dict(rays=rays, c2w=c2w, centers=centers, ext=ext, mlp=mlp, box_mlps=box_mlps, targets=tg, t_rand=t_rand,
                u_rand=u_rand, B=B, K=K)


def oracle_params(sc, dtype=torch.float32):
    cv = lambda layers: [(torch.from_numpy(k).to(dtype), torch.from_numpy(b).to(dtype)) for k, b in layers]
    return dict(mlp=cv(sc['mlp']), box_mlps=[cv(m) for m in sc['box_mlps']], box_centers=torch.from_numpy(sc['centers']).to(dtype))


def oracle_rays(sc, dtype=torch.float32):
    return O.Rays(*[torch.from_numpy(np.asarray(a)).to(dtype) for a in sc['rays']])


def cuda_rays(sc):
    from durf_b200.utils import Rays
    return Rays(*[torch.from_numpy(np.asarray(a)).cuda() for a in sc['rays']])


def cuda_variables(sc, model):
    from durf_b200.obbpose_model import Variables
    v = Variables.allocate(model, sc['K'], sc['centers'].shape[0], 'cuda')
    v.load_mlp('MLP_0', sc['mlp'])
    for k, m in enumerate(sc['box_mlps']):
        v.load_mlp(f'BoxMLP_{k}', m)
    v.box_centers.copy_(torch.from_numpy(sc['centers']).cuda())
    v.mark_dirty()
    return v


def flat_oracle_grads(sc, variables, grads_by_name):
    """Lay oracle gradients (dict name -> list[(dk, db)] / tensor) out like Variables.flat."""
    flat = torch.zeros(variables.flat.numel(), dtype=torch.float64)
    for name, (off, n) in variables.slots.items():
        g = grads_by_name.get(name)
        if g is None:
            continue
        if name == 'box_centers':
            flat[off:off + n] = g.reshape(-1).double()
        else:
            flat[off:off + n] = torch.cat([torch.cat([dk.reshape(-1), db.reshape(-1)]) for dk, db in g]).double()
    return flat


def assert_close(got, want, rtol=1e-5, atol_scale=1.0, what=""):
    """|a-b| <= rtol * max(|b|, atol_scale)  (SURVEY §7 tolerance form)."""
    got = got.detach().double().cpu()
    want = want.detach().double().cpu()
    assert got.shape == want.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    tol = rtol * torch.clamp(want.abs(), min=atol_scale)
    err = (got - want).abs()
    bad = err > tol
    if bad.any():
        i = torch.argmax(err - tol)
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} beyond tol; worst |d|={float(err.flatten()[i]):.3e} "
                             f"want={float(want.flatten()[i]):.6e} got={float(got.flatten()[i]):.6e}")


def unswizzle_tiles(tiles: torch.Tensor, rows: int, F: int) -> torch.Tensor:
    """bf16 128x64 SWIZZLE_128B tile images -> [rows, F] float32."""
    t = tiles.view(torch.bfloat16).reshape(-1, 128 * 64).float().cpu()
    ntiles = t.shape[0]
    r = torch.arange(128)
    out = torch.empty(ntiles, 128, 64)
    for c in range(8):
        off = (r // 8) * 512 + (r % 8) * 64 + ((c ^ (r % 8)) * 8)      # element offsets (bytes / 2)
        idx = off[:, None] + torch.arange(8)[None, :]
        out[:, :, c * 8:(c + 1) * 8] = t[:, idx.reshape(-1)].reshape(ntiles, 128, 8)
    return out.reshape(ntiles * 128, 64)[:rows, :F]


def sampler_pdf_cdf(bins, weights):
    """fp64 restatement of the pdf/cdf construction of math.sorted_piecewise_constant_pdf (math.py:237-251)."""
    w = weights.double().cpu()
    ws = w.sum(-1, keepdim=True)
    pad = torch.clamp(1e-5 - ws, min=0.0)
    w = w + pad / w.shape[-1]
    pdf = w / (ws + pad)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.clamp(torch.cumsum(pdf[..., :-1], -1), max=1.0),
                     torch.ones_like(pdf[..., :1])], -1)
    return pdf, cdf


def cdf_at(bins, cdf, t):
    """Piecewise-linear CDF F(t) (fp64) of the histogram (bins, cdf) at positions t [..., S]."""
    b = bins.double().cpu().contiguous()
    t = t.double().cpu().contiguous()
    j = torch.clamp(torch.searchsorted(b, t, right=True) - 1, 0, b.shape[-1] - 2)
    b0, b1 = torch.gather(b, -1, j), torch.gather(b, -1, j + 1)
    c0, c1 = torch.gather(cdf, -1, j), torch.gather(cdf, -1, j + 1)
    frac = torch.where(b1 > b0, (t - b0) / torch.clamp(b1 - b0, min=1e-300), torch.zeros_like(t))
    return c0 + torch.clamp(frac, 0.0, 1.0) * (c1 - c0)


def assert_samples_close(got, want, bins, weights, what="sampler", pos_rtol=1e-5, cdf_ulps=64):
    """Inverse-CDF samples are ill-conditioned in position where a bin holds little mass (a 1-ulp difference in the
    fp32 cumsum moves a sample by ulp * width / mass), but well-conditioned in CDF space.  A sample passes if it agrees
    in position (pos_rtol * max(|t|,1)) OR in CDF space (cdf_ulps * eps32): the summation order of the scan (warp scan
    here, XLA's associative scan in the reference, sequential in the oracle) is the only thing that differs."""
    _, cdf = sampler_pdf_cdf(bins, weights)
    got64, want64 = got.double().cpu(), want.double().cpu()
    pos_ok = (got64 - want64).abs() <= pos_rtol * torch.clamp(want64.abs(), min=1.0)
    dF = (cdf_at(bins, cdf, got64) - cdf_at(bins, cdf, want64)).abs()
    cdf_ok = dF <= cdf_ulps * 1.1920929e-07
    bad = ~(pos_ok | cdf_ok)
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{bad.numel()} samples differ in position and in CDF space; "
                                 f"worst dF={float(dF[bad].max()):.3e}")
    return float(pos_ok.double().mean())


def _bf16_ste(x):
    """Round to bf16 in the forward pass, identity in the backward pass (straight-through)."""
    return x + (x.to(torch.bfloat16).to(x.dtype) - x).detach()


def mlp_apply_bf16_emulated(params, topo, x, condition):
    """The oracle's MLP (oracle.mlp_apply, obbpose_model.py:305-354) with the tensor-core path's quantisation points made
    explicit: GEMM operands (input features, trunk/bottleneck activations, kernels of the tensor-core layers) are rounded
    to bf16, accumulation, biases, the density / rgb heads and the 27 view inputs stay fp32.  Test infrastructure: it
    separates "bf16 arithmetic" from "kernel bug" when the CUDA gradients are compared (ReLU masks then agree)."""
    B, N, F = x.shape
    x = _bf16_ste(x.reshape(-1, F))
    inputs = x
    li = 0
    q = _bf16_ste
    h = x
    for i in range(topo.depth):
        k, b = params[li]; li += 1
        pre = h @ q(k) + b
        a = torch.relu(pre)
        h_fp32 = a                     # the density head reads the fp32 ReLU output of the last trunk layer
        h = q(a)
        if i % topo.skip == 0 and i > 0:
            h = torch.cat([h, inputs], dim=-1)
    k, b = params[li]; li += 1
    raw_density = (h_fp32 @ k + b).reshape(-1, N, 1)
    k, b = params[li]; li += 1
    bott = q(h @ q(k) + b)
    k, b = params[li]; li += 1
    W = topo.width
    cond = condition[:, None, :].expand(B, N, condition.shape[-1]).reshape(-1, condition.shape[-1])
    c = torch.relu(bott @ q(k[:W]) + cond @ k[W:] + b)
    k, b = params[li]
    raw_rgb = (c @ k + b).reshape(-1, N, 3)
    return raw_rgb, raw_density
