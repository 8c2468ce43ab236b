"""Import the UNMODIFIED reference (/root/reference) in the build container (TEST INFRASTRUCTURE).

The reference's `pnpflow.utils` imports metric/plot packages that are absent here (matplotlib,
torchmetrics, ignite, deepinv, lpips; utils.py:16,19,20,29,33).  None of them touches the arithmetic of
the pnp_flow path, so empty stub modules are injected before import.  The reference tree
does not exist on the GPU box: callers must check ``available()`` and skip.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("PNPFLOW_REFERENCE_ROOT", "/root/reference")

_STUBS = ["matplotlib", "matplotlib.pyplot", "torchmetrics", "torchmetrics.functional",
          "torchmetrics.functional.image", "ignite", "ignite.metrics", "deepinv", "lpips", "gdown"]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "pnpflow"))


def load():
    """Returns (pnpflow.utils, pnpflow.degradations, pnpflow.methods.pnp_flow, pnpflow.models)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    img = sys.modules["torchmetrics.functional.image"]
    if not hasattr(img, "peak_signal_noise_ratio"):
        img.peak_signal_noise_ratio = None
    ign = sys.modules["ignite.metrics"]
    if not hasattr(ign, "SSIM"):
        ign.SSIM = None
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pnpflow.utils as U
        import pnpflow.degradations as D
        import pnpflow.methods.pnp_flow as M
        import pnpflow.models as N
    return U, D, M, N


def build_reference_unet(cfg, state_dict):
    """Instantiate the reference nn.Module UNet for an oracle UNetConfig and load ``state_dict`` into it."""
    _, _, _, N = load()
    net = N.UNet(input_channels=cfg.input_channels, input_height=cfg.input_height, ch=cfg.ch,
                 ch_mult=tuple(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks,
                 attn_resolutions=tuple(cfg.attn_resolutions), resamp_with_conv=True)
    net.load_state_dict(state_dict, strict=True)
    net.eval()
    return net


class RefArgs:
    """Attribute bag with the `args` keys PNP_FLOW reads."""
    def __init__(self, **kw):
        d = dict(method='pnp_flow', model='ot', noise_type='gaussian', gamma_style='alpha_1_minus_t', alpha=1.0,
                 steps_pnp=100, lr_pnp=1.0, num_samples=5, max_batch=1, compute_time=False, compute_memory=False,
                 save_results=False, batch=0, save_path='/tmp/pnpflow_ref', eval_split='test',
                 dict_cfg_method={}, problem='denoising', dataset='celeba', dim_image=128, num_channels=3)
        d.update(kw)
        self.__dict__.update(d)


def run_reference_solve_ip(net, clean_batches, degradation, sigma_noise, args):
    """Drive the unmodified PNP_FLOW.solve_ip and capture the final x of every batch.

    solve_ip returns nothing (results leave through utils.save_images/compute_*, pnp_flow.py:128-163), so the
    sinks on the `pnpflow.utils` module object are patched: with save_results=True the final
    ``utils.save_images(clean, noisy, restored, ...)`` call (:156) hands us both y and x.
    """
    import torch
    U, _, M, _ = load()
    captured = []
    saved = {k: getattr(U, k) for k in ("save_images", "compute_psnr", "compute_ssim", "compute_lpips",
                                         "compute_average_psnr", "compute_average_ssim", "compute_average_lpips")}

    def _noop(*a, **k):
        return None

    def _grab(clean_img, noisy_img, restored_img, args_, H_adj, iter='final'):
        captured.append((noisy_img.detach().clone(), restored_img.detach().clone()))

    try:
        for k in saved:
            setattr(U, k, _noop)
        U.save_images = _grab
        args.save_results = True
        args.max_batch = len(clean_batches)
        os.makedirs(args.save_path, exist_ok=True)
        args.save_path_ip = args.save_path
        method = M.PNP_FLOW(net, torch.device('cpu'), args)
        loader = [(c, torch.zeros(len(c))) for c in clean_batches]
        # steps//10 == 0 would raise ZeroDivisionError in should_save_image
        assert args.steps_pnp >= 10
        method.solve_ip(loader, degradation, sigma_noise)
    finally:
        for k, v in saved.items():
            setattr(U, k, v)
    return captured
