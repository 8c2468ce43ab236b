"""Oracle restatement of the PnP-Flow restoration loop (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/pnpflow/methods/pnp_flow.py:
  learning_rate_strat   :29-37      grad_datafit        :39-45
  interpolation_step    :47-48      denoiser            :50-52
  lr pre-scale          :60-66      measurement         :77-80
  x0 = H_adj(ones)      :93         hot loop            :102-121
and the PSNR definition of pnpflow/utils.py:560-577,594-610 (images mapped (x+1)/2, data_range 1, per image).
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Union

import torch


def learning_rate(lr: float, t: torch.Tensor, gamma_style: str, alpha: float) -> Union[torch.Tensor, float]:
    """pnp_flow.py:29-37; unknown styles fall back to constant lr exactly like the reference's dict.get."""
    t = t.view(-1, 1, 1, 1)
    if gamma_style == '1_minus_t':
        return lr * (1 - t)
    if gamma_style == 'sqrt_1_minus_t':
        return lr * torch.sqrt(1 - t)
    if gamma_style == 'alpha_1_minus_t':
        return lr * (1 - t) ** alpha
    return lr


def grad_datafit(x, y, H, H_adj, sigma_noise: float, noise_type: str = 'gaussian'):
    """pnp_flow.py:39-45."""
    if noise_type == 'gaussian':
        return H_adj(H(x) - y) / (sigma_noise ** 2)
    if noise_type == 'laplace':
        return H_adj(2 * torch.heaviside(H(x) - y, torch.zeros_like(H(x))) - 1) / sigma_noise
    raise ValueError('Noise type not supported')


def synthesize_measurement(clean: torch.Tensor, H, sigma_noise: float, batch_idx: int,
                           noise_type: str = 'gaussian') -> torch.Tensor:
    """pnp_flow.py:77-80 (gaussian): y = H(clean); torch.manual_seed(batch); y += sigma * randn_like(y).
    pnp_flow.py:81-85 (laplace): y = H(clean) + Laplace(0, sigma).sample() from the global generator (not re-seeded)."""
    y = H(clean.clone())
    if noise_type == 'gaussian':
        torch.manual_seed(batch_idx)
        y = y + torch.randn_like(y) * sigma_noise
    elif noise_type == 'laplace':
        y = y + torch.distributions.laplace.Laplace(torch.zeros_like(y), sigma_noise * torch.ones_like(y)).sample()
    else:
        raise ValueError('Noise type not supported')
    return y


def pnp_flow_restore(model_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], y: torch.Tensor,
                     degradation, sigma_noise: float, *, steps_pnp: int = 100, lr_pnp: float = 1.0,
                     alpha: float = 1.0, gamma_style: str = 'alpha_1_minus_t', num_samples: int = 5,
                     noise_type: str = 'gaussian', noise: Optional[Iterable[torch.Tensor]] = None,
                     trace: Optional[Callable[[int, torch.Tensor], None]] = None) -> torch.Tensor:
    """The T-step loop of PNP_FLOW.solve_ip for one batch (pnp_flow.py:93,102-121); returns the final x.

    ``noise``: optional iterable yielding one eps tensor per (step, draw) in loop order; when None the
    global torch generator is consumed with one ``randn_like`` per draw, exactly like the reference (:48).
    """
    H, H_adj = degradation.H, degradation.H_adj
    if noise_type == 'gaussian':
        lr = sigma_noise ** 2 * lr_pnp                 # :60-62 (sigma^2 cancels against grad_datafit's division)
    elif noise_type == 'laplace':
        lr = sigma_noise * lr_pnp                      # :64-66
    else:
        raise ValueError('Noise type not supported')
    steps, delta = steps_pnp, 1 / steps_pnp
    noise_it = iter(noise) if noise is not None else None
    x = H_adj(torch.ones_like(y))                      # :93
    with torch.no_grad():
        for iteration in range(int(steps)):
            t1 = torch.ones(len(x), device=x.device) * delta * iteration      # :107-108
            lr_t = learning_rate(lr, t1, gamma_style, alpha)
            z = x - lr_t * grad_datafit(x, y, H, H_adj, sigma_noise, noise_type)
            x_new = torch.zeros_like(x)
            tb = t1.view(-1, 1, 1, 1)
            for _ in range(num_samples):
                eps = next(noise_it) if noise_it is not None else torch.randn_like(z)
                z_tilde = tb * z + eps * (1 - tb)                              # :47-48
                x_new += z_tilde + (1 - tb) * model_fn(z_tilde, t1)            # :50-52,118
            x_new /= num_samples
            x = x_new
            if trace is not None:
                trace(iteration, x)
    return x


def psnr(rec: torch.Tensor, clean: torch.Tensor) -> torch.Tensor:
    """Per-image PSNR in dB on (x+1)/2 with data_range 1 (utils.py:560-577 postprocess, :594-610)."""
    a = (rec.double() + 1) / 2
    b = (clean.double() + 1) / 2
    mse = ((a - b) ** 2).flatten(1).mean(dim=1)
    return 10 * torch.log10(1.0 / mse)
