"""pnpflow_b200 — B200-native (sm_100a) engine for PnP-Flow's ``method=pnp_flow`` restoration hot path.

Public surface (mirrors the reference's plugin API):
    PNP_FLOW(model, device, args).run_method(...)      method plugin      (pnpflow/methods/pnp_flow.py)
    Denoising, BoxInpainting, RandomInpainting, PaintbrushInpainting, GaussianDeblurring, Superresolution
                                                         operator plugins   (pnpflow/degradations.py)
    UNetEngine(model_or_cfg, state_dict)(x, t)          velocity prior     (pnpflow/models.py UNet)
    restore(engine, y, degradation, sigma, ...) -> x     the loop as a function (the reference only writes files)
    generate_samples(engine, n, ...)                     Euler sampling of the flow ODE with the same engine
                                                         (pnpflow/train_flow_matching.py:170-198)
All device work happens in libpnpflow_sm100a.so (hand-written CUDA behind the C ABI of include/pnpflow_b200.h);
importing this package never imports ``oracle`` and there is no CPU / PyTorch fallback.
"""
from .degradations import (BoxInpainting, Degradation, Denoising, GaussianDeblurring, PaintbrushInpainting,  # noqa: F401
                           RandomInpainting, Superresolution, as_engine_operator)
from .engine import UNetEngine  # noqa: F401
from .sampler import generate_samples  # noqa: F401
from .method import PNP_FLOW, PnPFlowSession, gamma_schedule, psnr, restore, step_time  # noqa: F401

__all__ = ["PNP_FLOW", "PnPFlowSession", "restore", "UNetEngine", "Degradation", "Denoising", "BoxInpainting", "RandomInpainting",
           "PaintbrushInpainting", "GaussianDeblurring", "Superresolution", "as_engine_operator", "gamma_schedule", "psnr", "step_time", "generate_samples"]
