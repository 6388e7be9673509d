"""Mirror of the reference's internal/math.py pieces on the hot path: the inverse-CDF sampler (device) and the
host-side schedules (plain Python scalars in the reference as well)."""
import math as _m

import torch

from . import ops


def sorted_piecewise_constant_pdf(key, bins, weights, num_samples, randomized):
    """math.py:222-284.  bins [B,N+1], weights [B,N] (N <= 128) -> samples [B,num_samples].
    `key`: u_rand [B,num_samples] U[0,1) when randomized (else drawn on the device)."""
    u = None
    if randomized:
        u = key if torch.is_tensor(key) else torch.rand(bins.shape[0], num_samples, device=bins.device)
    return ops.resample(bins, weights, u_rand=u, blurpool=False, num_samples=num_samples)


def learning_rate_decay(step, lr_init, lr_final, max_steps, lr_delay_steps=0, lr_delay_mult=1):
    """math.py:156-190."""
    if lr_delay_steps > 0:
        delay_rate = lr_delay_mult + (1 - lr_delay_mult) * _m.sin(0.5 * _m.pi * min(max(step / lr_delay_steps, 0), 1))
    else:
        delay_rate = 1.
    t = min(max(step / max_steps, 0), 1)
    return delay_rate * _m.exp(_m.log(lr_init) * (1 - t) + _m.log(lr_final) * t)


def freq_alpha_rate(step, alpha_init, alpha_final, alpha_delay_steps, alpha_max_steps):
    """math.py:193-219."""
    if step < alpha_delay_steps:
        return alpha_init
    if step < alpha_max_steps:
        return (step - alpha_delay_steps) / (alpha_max_steps - alpha_delay_steps) * alpha_final
    return alpha_final


def mse_to_psnr(mse):
    """math.py:49-51."""
    return -10. / _m.log(10.) * torch.log(torch.as_tensor(mse))


def psnr_to_mse(psnr):
    """math.py:54-56."""
    return torch.exp(-0.1 * _m.log(10.) * torch.as_tensor(psnr))


def compute_avg_error(psnr, ssim, lpips):
    """math.py:59-63: geometric mean of MSE, sqrt(1 - SSIM) and LPIPS."""
    vals = torch.stack([psnr_to_mse(psnr), torch.sqrt(1 - torch.as_tensor(ssim)), torch.as_tensor(lpips)]).to(torch.float64)
    return torch.exp(torch.mean(torch.log(vals)))


def compute_ssim(img0, img1, max_val, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03, return_map=False):
    """math.py:66-137 (modelled after tf.image.ssim): images [..., H, W, C]; a separable Gaussian window applied as two
    'valid' 1-D convolutions per channel; returns the mean SSIM per image, or the map [..., H-fs+1, W-fs+1, C].
    Evaluation-time metric (host API of the render path's output); runs on whatever device the images are on."""
    img0, img1 = torch.as_tensor(img0), torch.as_tensor(img1)
    dt = img0.dtype if img0.dtype in (torch.float32, torch.float64) else torch.float32
    img0, img1 = img0.to(dt), img1.to(dt)
    hw = filter_size // 2
    shift = (2 * hw - filter_size + 1) / 2
    f_i = ((torch.arange(filter_size, dtype=dt, device=img0.device) - hw + shift) / filter_sigma) ** 2
    filt = torch.exp(-0.5 * f_i)
    filt = filt / filt.sum()
    lead = img0.shape[:-3]
    H, W, Cn = img0.shape[-3:]

    def blur(z):      # [..., H, W, C] -> [..., H-fs+1, W-fs+1, C]
        z = z.reshape(-1, H, W, Cn).permute(0, 3, 1, 2).reshape(-1, 1, H, W)
        # the Gaussian is symmetric, so correlation (conv2d) equals the reference's convolution
        z = torch.nn.functional.conv2d(z, filt.view(1, 1, 1, -1))
        z = torch.nn.functional.conv2d(z, filt.view(1, 1, -1, 1))
        h, w = z.shape[-2:]
        return z.reshape(-1, Cn, h, w).permute(0, 2, 3, 1).reshape(*lead, h, w, Cn)

    mu0, mu1 = blur(img0), blur(img1)
    mu00, mu11, mu01 = mu0 * mu0, mu1 * mu1, mu0 * mu1
    sigma00 = torch.clamp(blur(img0 ** 2) - mu00, min=0.)
    sigma11 = torch.clamp(blur(img1 ** 2) - mu11, min=0.)
    sigma01 = blur(img0 * img1) - mu01
    sigma01 = torch.sign(sigma01) * torch.minimum(torch.sqrt(sigma00 * sigma11), torch.abs(sigma01))
    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    ssim_map = ((2 * mu01 + c1) * (2 * sigma01 + c2)) / ((mu00 + mu11 + c1) * (sigma00 + sigma11 + c2))
    return ssim_map if return_map else ssim_map.mean(dim=(-3, -2, -1))


def linear_to_srgb(linear):
    """math.py:140-145."""
    linear = torch.as_tensor(linear)
    eps = torch.finfo(torch.float32).eps
    srgb0 = 323 / 25 * linear
    srgb1 = (211 * torch.clamp(linear, min=eps) ** (5 / 12) - 11) / 200
    return torch.where(linear <= 0.0031308, srgb0, srgb1)


def srgb_to_linear(srgb):
    """math.py:148-153."""
    srgb = torch.as_tensor(srgb)
    eps = torch.finfo(torch.float32).eps
    linear0 = 25 / 323 * srgb
    linear1 = torch.clamp((200 * srgb + 11) / 211, min=eps) ** (12 / 5)
    return torch.where(srgb <= 0.04045, linear0, linear1)
