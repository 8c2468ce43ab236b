"""Forward-only sibling of the PnP-Flow path that shares the U-Net engine: Euler sampling of the
flow-matching ODE dx/dt = v_theta(x, t), mirror of ``FLOW_MATCHING.generate_samples(integration_method="euler")``
(pnpflow/train_flow_matching.py:170-198; the reference delegates the time stepping to torchdiffeq's fixed-grid Euler solver).

Every step is one C-ABI call (``pnpf_euler_step``: time fill, U-Net evaluation, x += dt * v); t0 and dt are call arguments, so
the steps are launched eagerly (the ~190 launches of an evaluation are asynchronous and the U-Net dominates).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from .engine import UNetEngine


def generate_samples(engine: UNetEngine, n_samples: int = 16, batch_size: Optional[int] = None, integration_steps: int = 100,
                     tmax: float = 1, x0: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Images [n_samples, C, H, W] (fp32, CUDA).  ``x0``: optional latents (else N(0, I) from ``generator`` / the global one,
    one ``torch.randn`` per batch like :187-188).  Same batching rule as the reference (:176-181)."""
    lib = _lib.load()
    dev = engine.device
    Cc, Hh = engine.cfg["input_channels"], engine.cfg["input_height"]
    if batch_size is None:
        batch_size = n_samples
    batches = [batch_size] * (n_samples // batch_size)
    if n_samples % batch_size:
        batches += [n_samples % batch_size]
    # time grid in fp32 like torch.linspace(0, tmax, int(tmax * steps)) (:184-185); dt_k = t_{k+1} - t_k in fp32
    tp = torch.linspace(0, tmax, int(tmax * integration_steps)).numpy().astype(np.float32)
    out, off = [], 0
    with torch.no_grad(), torch.cuda.device(dev):
        for b in batches:
            engine.ensure_batch(b)
            if x0 is not None:
                x = x0[off:off + b].to(dev, torch.float32).contiguous().clone()
            else:
                x = torch.randn(b, Cc, Hh, Hh, device=dev, generator=generator)
            off += b
            v = torch.empty_like(x)
            tdev = torch.empty(b, device=dev)
            for k in range(len(tp) - 1):
                _lib.check(lib.pnpf_euler_step(engine._h, x.data_ptr(), float(tp[k]), float(np.float32(tp[k + 1] - tp[k])), b, Cc, Hh, Hh,
                                               tdev.data_ptr(), v.data_ptr(), x.data_ptr(), _lib.stream_ptr()))
            out.append(x)
    return torch.cat(out, dim=0)
