"""PNP_FLOW: drop-in for the reference's method plugin (pnpflow/methods/pnp_flow.py:10-188) on the sm_100a engine.

Same constructor ``PNP_FLOW(model, device, args)``, same ``run_method / solve_ip`` and the same public helper
methods and ``args`` keys.  ``model`` may be the reference's ``pnpflow.models.UNet`` module (its
state_dict is repacked into the engine), an ``oracle``-style ``(cfg, state_dict)`` pair, or a ready ``UNetEngine``.

Per outer step (pnp_flow.py:102-121) the engine runs
    1. one fused data-fidelity kernel          z = x - gamma_t * A^T(Ax - y)            (:29-45,111-112)
       (noise_type='laplace': A^T(2*heaviside(Ax - y, 0) - 1), :42-43)
    2. one interpolation kernel for all S draws  z~_s = t z + (1-t) eps_s               (:47-48)
    3. ONE U-Net evaluation on the S*B batch    v_s = v_theta(z~_s, t)                  (:19-21)  [CUDA-graph replay]
    4. one push+average kernel                 x = (1/S) sum_s (z~_s + (1-t) v_s)       (:50-52,114-121)
The noise eps_s comes from ``torch.randn_like`` in the reference's call order (one call per draw), so a run with the
same seed consumes the same Philox stream as the reference's GPU path; ``noise=`` injects explicit draws instead.
"""
from __future__ import annotations

import ctypes as C
import os
from time import perf_counter
from typing import Callable, Iterable, Optional

import numpy as np
import torch

from . import _lib
from .degradations import as_engine_operator
from .engine import UNetEngine


def gamma_schedule(lr_pnp: float, t: float, gamma_style: str, alpha: float) -> float:
    """lr_t / sigma^2 of the reference (pnp_flow.py:29-37 with the sigma^2 of :60-62 cancelled against :41),
    evaluated in fp32 like the reference's tensor arithmetic."""
    omt = np.float32(1.0) - np.float32(t)
    if gamma_style == '1_minus_t':
        g = omt
    elif gamma_style == 'sqrt_1_minus_t':
        g = np.sqrt(omt)
    elif gamma_style == 'alpha_1_minus_t':
        g = np.power(omt, np.float32(alpha))
    else:                                           # 'constant' and unknown styles (dict.get default, :37)
        g = np.float32(1.0)
    return float(np.float32(lr_pnp) * np.float32(g))


def step_time(delta: float, it: int) -> float:
    """t of iteration ``it`` exactly as the reference forms it (pnp_flow.py:107-108): ``torch.ones(B) * delta * iteration`` is an
    fp32 tensor times python scalars, i.e. fp32(fp32(delta) * fp32(it)) — NOT fp32(double(delta * it))."""
    return float(np.float32(np.float32(delta) * np.float32(it)))


def psnr(rec: torch.Tensor, clean: torch.Tensor) -> torch.Tensor:
    """Per-image PSNR (dB) on (x+1)/2, data_range 1 — the reference's definition (utils.py:560-577,594-610)."""
    a = (rec.double() + 1) / 2
    b = (clean.to(rec.device).double() + 1) / 2
    return 10 * torch.log10(1.0 / ((a - b) ** 2).flatten(1).mean(dim=1))


def restore(engine: UNetEngine, y: torch.Tensor, degradation, sigma_noise: float, *, steps_pnp: int = 100,
            lr_pnp: float = 1.0, alpha: float = 1.0, gamma_style: str = 'alpha_1_minus_t', num_samples: int = 5,
            noise_type: str = 'gaussian', noise: Optional[Iterable[torch.Tensor]] = None,
            trace: Optional[Callable[[int, torch.Tensor], None]] = None, use_cuda_graph: bool = True,
            whole_step_call: bool = False) -> torch.Tensor:
    """The T-step PnP-Flow loop for one batch of measurements ``y`` (CUDA fp32); returns the restored x.

    Mirrors pnp_flow.py:93,102-121.  ``noise``: optional iterable of eps tensors [B,C,H,W], one per (step, draw).
    """
    sess = PnPFlowSession(engine, degradation, tuple(y.shape), steps_pnp=steps_pnp, lr_pnp=lr_pnp, alpha=alpha,
                          gamma_style=gamma_style, num_samples=num_samples, noise_type=noise_type,
                          use_cuda_graph=use_cuda_graph, device=y.device, whole_step_call=whole_step_call)
    return sess.run(y, noise=noise, trace=trace)


class PnPFlowSession:
    """Static buffers + captured U-Net graph for repeated PnP-Flow steps on measurements of a fixed shape.

    ``step(x, y, it)`` is ONE iteration of pnp_flow.py:107-121 at t = it/steps_pnp: data-fidelity kernel, S noise
    draws, one interpolation kernel, one U-Net evaluation on the S*B batch, one push+average kernel."""

    def __init__(self, engine: UNetEngine, degradation, y_shape, *, steps_pnp=100, lr_pnp=1.0, alpha=1.0,
                 gamma_style='alpha_1_minus_t', num_samples=5, noise_type='gaussian', use_cuda_graph=True, device="cuda",
                 whole_step_call=False):
        if noise_type not in ('gaussian', 'laplace'):
            raise ValueError('Noise type not supported')                     # pnp_flow.py:45,68,87
        self.noise_type = noise_type
        self.lib = _lib.load()
        self.engine = engine
        self.op = as_engine_operator(degradation)
        self.dev = torch.device(device)
        self.steps, self.lr_pnp, self.alpha, self.gamma_style = int(steps_pnp), lr_pnp, alpha, gamma_style
        self.delta = 1 / steps_pnp
        self.S = S = int(num_samples)
        self.y_shape = tuple(y_shape)
        B, Cc = y_shape[0], y_shape[1]
        Hh = Ww = engine.cfg["input_height"]
        self.shape = (B, Cc, Hh, Ww)
        self.n = B * Cc * Hh * Ww
        # whole_step_call: run every iteration through the single C entry point pnpf_step (what a non-Python host would call)
        # instead of sequencing data-fit / interpolate / U-Net graph replay / push from here; eager launches, same kernels
        self.whole_step_call = bool(whole_step_call) and hasattr(self.op, "descriptor") and type(self.op).__name__ != "_PythonOperator"
        if self.whole_step_call:
            use_cuda_graph = False
        self.use_cuda_graph = use_cuda_graph
        with torch.cuda.device(self.dev):
            self.z = torch.empty(self.shape, device=self.dev)
            self.eps = torch.empty((S,) + self.shape, device=self.dev)
            self.xbuf = [torch.empty(self.shape, device=self.dev) for _ in range(2)]
            if use_cuda_graph:
                zt, self.tb, v, self.replay = engine.graphed(S * B)
            else:
                zt = torch.empty((S * B, Cc, Hh, Ww), device=self.dev)
                v = torch.empty_like(zt)
                self.tb = torch.empty(S * B, device=self.dev)
        self.zt, self.v = zt, v
        self._flip = 0
        # kernels of OUR library launched by one step (bench.py's gpu_launches): data-fit (2 for blur), interp,
        # every U-Net op except the stats memset, push
        self.launches_per_step = (2 if type(self.op).__name__ == "GaussianDeblurring" else 1) + 1 + (engine.num_launches - 1) + 1

    def initial_state(self, y):
        return self.op.H_adj(torch.ones_like(y))                             # pnp_flow.py:93

    def run(self, y, noise=None, trace=None):
        """All ``steps_pnp`` iterations for one batch of measurements (pnp_flow.py:93,102-121); returns the final x."""
        y = y.contiguous().float()
        x = self.initial_state(y)
        noise_it = iter(noise) if noise is not None else None
        for it in range(self.steps):
            x = self.step(x, y, it, noise_it)
            if trace is not None:
                trace(it, x)
        return x.clone()

    def step(self, x, y, it: int, noise_it=None):
        S, n = self.S, self.n
        sp = _lib.stream_ptr
        t = step_time(self.delta, it)                                        # :107-108
        gamma = gamma_schedule(self.lr_pnp, t, self.gamma_style, self.alpha)
        with torch.no_grad(), torch.cuda.device(self.dev):
            if self.whole_step_call:
                for s in range(S):
                    self.eps[s].copy_(next(noise_it) if noise_it is not None else torch.randn_like(self.z))
                B, Cc, Hh, Ww = self.shape
                opd, _keep = self.op.descriptor(B, Cc, Hh, Ww, self.dev)
                x_new = self.xbuf[self._flip]
                self._flip ^= 1
                _lib.check(self.lib.pnpf_step(self.engine._h, C.byref(opd), 1 if self.noise_type == 'laplace' else 0, x.data_ptr(),
                                              y.data_ptr(), self.eps.data_ptr(), t, gamma, S, B, Cc, Hh, Ww, self.z.data_ptr(),
                                              self.zt.data_ptr(), self.tb.data_ptr(), self.v.data_ptr(), x_new.data_ptr(), sp()))
                return x_new
            self.op.datafit_step(x, y, gamma, out=self.z, noise_type=self.noise_type)
            for s in range(S):
                if noise_it is not None:
                    self.eps[s].copy_(next(noise_it))
                else:
                    self.eps[s].copy_(torch.randn_like(self.z))              # one Philox call per draw, like :48
            _lib.check(self.lib.pnpf_interp(self.z.data_ptr(), self.eps.data_ptr(), t, self.zt.data_ptr(), n, S, sp()))
            self.tb.fill_(t)
            if self.use_cuda_graph:
                self.replay()
            else:
                self.engine.forward(self.zt, self.tb, out=self.v)
            x_new = self.xbuf[self._flip]
            if x_new.data_ptr() == x.data_ptr():
                self._flip ^= 1
                x_new = self.xbuf[self._flip]
            self._flip ^= 1
            _lib.check(self.lib.pnpf_push_accum(self.zt.data_ptr(), self.v.data_ptr(), t, S, x_new.data_ptr(), n, sp()))
        return x_new


class PNP_FLOW(object):
    """Reference-compatible method plugin (pnp_flow.py:10-188)."""

    def __init__(self, model, device, args):
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.args = args
        self.method = args.method
        self.coupling = self.args.model
        if self.coupling not in {"ot", "indep"}:
            raise ValueError("pnpflow_b200 implements the Flow-Matching U-Net prior (args.model in {'ot','indep'}); "
                             f"got {self.coupling!r}")
        if isinstance(model, UNetEngine):
            self.model = model
        elif isinstance(model, tuple):
            self.model = UNetEngine(model[0], model[1], device=self.device)
        else:
            self.model = UNetEngine(model, None, device=self.device)
        self.results = []                     # (clean, noisy, restored) per batch — the reference only writes files

    # ---- public helpers with the reference's signatures -----------------------------------------------------
    def model_forward(self, x, t):
        return self.model(x, t)               # pnp_flow.py:19-21 ('ot' / 'indep')

    def learning_rate_strat(self, lr, t):
        t = t.view(-1, 1, 1, 1)
        style, a = self.args.gamma_style, self.args.alpha
        if style == '1_minus_t':
            return lr * (1 - t)
        if style == 'sqrt_1_minus_t':
            return lr * torch.sqrt(1 - t)
        if style == 'alpha_1_minus_t':
            return lr * (1 - t) ** a
        return lr

    def grad_datafit(self, x, y, H, H_adj):
        if self.args.noise_type == 'gaussian':
            return H_adj(H(x) - y) / (self.args.sigma_noise ** 2)
        elif self.args.noise_type == 'laplace':
            return H_adj(2 * torch.heaviside(H(x) - y, torch.zeros_like(H(x))) - 1) / self.args.sigma_noise
        else:
            raise ValueError('Noise type not supported')

    def interpolation_step(self, x, t):
        eps = torch.randn_like(x)
        x = x.contiguous()
        tt = float(t.flatten()[0]) if torch.is_tensor(t) else float(t)
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pnpf_interp(x.data_ptr(), eps.data_ptr(), tt, out.data_ptr(), x.numel(), 1, _lib.stream_ptr()))
        return out

    def denoiser(self, x, t):
        v = self.model_forward(x, t)
        tt = float(t.flatten()[0])
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pnpf_push_accum(x.contiguous().data_ptr(), v.data_ptr(), tt, 1, out.data_ptr(), x.numel(),
                                                   _lib.stream_ptr()))
        return out

    def should_save_image(self, iteration, steps):
        return iteration % (steps // 10) == 0

    # ---- the solver ------------------------------------------------------------------------------------------
    def solve_ip(self, test_loader, degradation, sigma_noise, H_funcs=None):
        a = self.args
        a.sigma_noise = sigma_noise
        if a.noise_type == 'gaussian':
            lr_eff = a.lr_pnp                       # gamma_t = lr_pnp (1-t)^alpha: sigma^2 cancels (:41 vs :61)
            a.lr_pnp = sigma_noise ** 2 * a.lr_pnp  # keep the reference's observable side effect on args (:61)
        elif a.noise_type == 'laplace':
            lr_eff = a.lr_pnp                       # lr_t * grad = sigma lr_pnp g(t) * A^T(sign)/sigma: sigma cancels (:43 vs :65)
            a.lr_pnp = sigma_noise * a.lr_pnp       # observable side effect on args (:65)
        else:
            raise ValueError('Noise type not supported')
        op = as_engine_operator(degradation)
        loader = iter(test_loader)
        save = bool(getattr(a, 'save_results', False))
        times = []
        sess = None
        for batch in range(a.max_batch):
            (clean_img, labels) = next(loader)
            a.batch = batch
            clean_dev = clean_img.clone().to(self.device).float()
            noisy_img = op.H(clean_dev)
            if a.noise_type == 'gaussian':
                torch.manual_seed(batch)                                     # :79
                noisy_img = noisy_img + torch.randn_like(noisy_img) * sigma_noise
            else:                                                            # :81-85 (unseeded in the reference too)
                noisy_img = noisy_img + torch.distributions.laplace.Laplace(
                    torch.zeros_like(noisy_img), sigma_noise * torch.ones_like(noisy_img)).sample()
            # buffers + the captured U-Net graph are set up OUTSIDE the timed region, once per measurement shape (the
            # reference's timer, :104-126, only covers the iterations)
            if sess is None or sess.y_shape != tuple(noisy_img.shape):
                sess = PnPFlowSession(self.model, op, tuple(noisy_img.shape), steps_pnp=a.steps_pnp, lr_pnp=lr_eff, alpha=a.alpha,
                                      gamma_style=a.gamma_style, num_samples=a.num_samples, noise_type=a.noise_type,
                                      device=self.device)
            trace = None
            if save:
                def trace(it, x, _clean=clean_dev, _noisy=noisy_img):         # :128-139: iteration 0, every 50th, every steps//10
                    if it % 50 == 0 or self.should_save_image(it, a.steps_pnp):
                        self._write_psnr(_clean, _noisy, x, op, it)
            if getattr(a, 'compute_time', False):
                torch.cuda.synchronize()
                t0 = perf_counter()
            if getattr(a, 'compute_memory', False):
                torch.cuda.reset_peak_memory_stats(self.device)
            x = sess.run(noisy_img, trace=trace)
            if getattr(a, 'compute_time', False):
                torch.cuda.synchronize()
                times.append(perf_counter() - t0)
                self._append_stat('time_stats.txt', {"batch": batch, "time_per_batch": times[-1]})
            if getattr(a, 'compute_memory', False):
                self._append_stat('memory_stats.txt', {"batch": batch, "max_allocated": torch.cuda.max_memory_allocated(self.device)})
            self.results.append((clean_img.cpu(), noisy_img.cpu(), x.cpu()))
            if save:
                self._write_psnr(clean_dev, noisy_img, x, op, a.steps_pnp - 1)   # :156-158 (iter = last loop index)
        return self.results

    def _write_psnr(self, clean, noisy, rec, op, it):
        """The reference's PSNR sink (utils.py:594-625): '{iter} {batch-mean PSNR}' lines appended to psnr_rec_batch{b}.txt and
        psnr_noisy_batch{b}.txt (same names and format, so utils.compute_average_psnr parses them).  SSIM / LPIPS sinks are
        out of scope."""
        noisy_full = op.H_adj(noisy) if tuple(noisy.shape) != tuple(clean.shape) else noisy     # utils.py:603-606
        b = self.args.batch
        self._append_lines(f'psnr_rec_batch{b}.txt', [f'{it} {psnr(rec, clean).mean().item()}'])
        self._append_lines(f'psnr_noisy_batch{b}.txt', [f'{it} {psnr(noisy_full, clean).mean().item()}'])

    def _append_stat(self, fname, d):
        path = getattr(self.args, 'save_path_ip', None)
        if path:
            os.makedirs(path, exist_ok=True)
            with open(os.path.join(path, fname), "a") as f:
                f.write(str(d) + '\n')

    def _append_lines(self, fname, lines):
        path = getattr(self.args, 'save_path_ip', None)
        if path:
            os.makedirs(path, exist_ok=True)
            with open(os.path.join(path, fname), "a") as f:
                f.write('\n'.join(lines) + '\n')

    def run_method(self, data_loaders, degradation, sigma_noise, H_funcs=None):
        folder = ""
        for key, value in getattr(self.args, 'dict_cfg_method', {}).items():      # utils.py:1112-1120
            folder = os.path.join(folder, f"{key}={value}")
        self.args.save_path_ip = os.path.join(self.args.save_path, folder)
        os.makedirs(self.args.save_path_ip, exist_ok=True)
        return self.solve_ip(data_loaders[self.args.eval_split], degradation, sigma_noise, H_funcs)
