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
