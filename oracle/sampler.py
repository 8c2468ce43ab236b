"""Oracle restatement of the forward-only flow-matching sampler (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/pnpflow/train_flow_matching.py:
  FLOW_MATCHING.generate_samples   :170-198   x0 ~ N(0, I); time_points = linspace(0, tmax, int(tmax * integration_steps));
                                              traj = odeint(cnf(model), x0, time_points, method=integration_method); traj[-1]
  cnf.forward                      :252-262   f(t, x) = model(x, t.repeat(x.shape[0]))
The ODE solver is the third-party `torchdiffeq` (listed un-pinned in the reference's pyproject.toml, absent from this image and
from /root/reference).  Its published fixed-grid Euler method (torchdiffeq/_impl/fixed_grid.py, class Euler:
`_step_func(func, t0, dt, t1, y0) -> dt * func(t0, y0)`; FixedGridODESolver.integrate: `y1 = y0 + dy` over consecutive grid
points when the output times ARE the grid) is restated here; parity is "unpinned" against torchdiffeq itself.
"""
from __future__ import annotations

from typing import Callable

import torch


def euler_sample(model_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], x0: torch.Tensor,
                 integration_steps: int = 100, tmax: float = 1) -> torch.Tensor:
    """traj[-1] of generate_samples(integration_method='euler') for one batch with the given latent x0."""
    time_points = torch.linspace(0, tmax, int(tmax * integration_steps), device=x0.device)      # :184-185
    y = x0
    with torch.no_grad():
        for t0, t1 in zip(time_points[:-1], time_points[1:]):
            dt = t1 - t0
            y = y + dt * model_fn(y, t0.repeat(y.shape[0]))                                      # Euler step; cnf.forward :258-262
    return y
