"""Operator plugins A / A^T on the sm_100a engine — mirror of the reference's pnpflow/degradations.py.

Same class names, constructor arguments, attributes and duck-typed surface (``H(x)``, ``H_adj(x)``;
degradations.py:6-12), so code written against the reference's operators keeps working; inside the PnP-Flow loop
the engine does not call H/H_adj separately but fuses ``x - gamma * A^T(Ax - y)`` into one kernel
(``datafit_step``), driven by the ``pnpf_operator`` descriptor each class exports through ``descriptor()``.

Host-side mask generation (numpy legacy RNG seeded 42 / python ``random`` + cv2 lines; reference utils.py:339-361,
904-969) happens ONCE per (B,H,W) and is cached on the device as uint8 — the reference regenerates it on the host
for every H/H_adj call.

``as_engine_operator(obj)`` also accepts the *reference's own* degradation objects (recognised by class name and
attributes); unknown Degradation subclasses are wrapped so that their Python H/H_adj are called between engine
kernels (operator-API fallback, not a CPU fallback: tensors stay on the GPU).
"""
from __future__ import annotations

import ctypes as C
import random as _pyrandom
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib


def _check_cuda(x: torch.Tensor):
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4):
        raise RuntimeError("pnpflow_b200 operators need CUDA fp32 [B,C,H,W] tensors (no CPU fallback)")


class Degradation:
    """Base class: subclasses provide ``descriptor(B, C, H, W, device)`` -> (_lib.OperatorC, keep-alive list)."""
    kind = None

    def descriptor(self, B, C, H, W, device):
        raise NotImplementedError()

    def out_shape(self, B, C, H, W):
        return (B, C, H, W)

    # -- reference surface (degradations.py:8-12)
    def H(self, x):
        _check_cuda(x)
        x = x.contiguous()
        B, Cc, Hh, Ww = x.shape
        op, _keep = self.descriptor(B, Cc, Hh, Ww, x.device)
        y = torch.empty(self.out_shape(B, Cc, Hh, Ww), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pnpf_apply_H(C.byref(op), x.data_ptr(), y.data_ptr(), B, Cc, Hh, Ww, _lib.stream_ptr()))
        return y

    def H_adj(self, y):
        _check_cuda(y)
        y = y.contiguous()
        B, Cc, Hh, Ww = self.full_shape(*y.shape)
        op, _keep = self.descriptor(B, Cc, Hh, Ww, y.device)
        x = torch.empty((B, Cc, Hh, Ww), device=y.device, dtype=torch.float32)
        with torch.cuda.device(y.device):
            _lib.check(_lib.load().pnpf_apply_H_adj(C.byref(op), y.data_ptr(), x.data_ptr(), B, Cc, Hh, Ww, _lib.stream_ptr()))
        return x

    def full_shape(self, B, C, Hy, Wy):
        return (B, C, Hy, Wy)

    # -- fused data-fidelity step  z = x - gamma * A^T(Ax - y)   (pnp_flow.py:39-41,111-112)
    def datafit_step(self, x, y, gamma: float, out=None, noise_type: str = 'gaussian'):
        """z = x - gamma * A^T r,  r = Ax - y (gaussian, pnp_flow.py:41) or 2*heaviside(Ax - y, 0) - 1 (laplace, :43)."""
        _check_cuda(x)
        if noise_type not in ('gaussian', 'laplace'):
            raise ValueError('Noise type not supported')                     # pnp_flow.py:45
        B, Cc, Hh, Ww = x.shape
        op, _keep = self.descriptor(B, Cc, Hh, Ww, x.device)
        z = out if out is not None else torch.empty_like(x)
        lib = _lib.load()
        fn = lib.pnpf_datafit_step if noise_type == 'gaussian' else lib.pnpf_datafit_step_laplace
        with torch.cuda.device(x.device):
            _lib.check(fn(C.byref(op), x.data_ptr(), y.data_ptr(), z.data_ptr(), float(gamma), B, Cc, Hh, Ww,
                          _lib.stream_ptr()))
        return z


def _op(kind, **kw):
    o = _lib.OperatorC()
    o.kind = kind
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Denoising(Degradation):
    """degradations.py:15-20."""
    def descriptor(self, B, C, H, W, device):
        return _op(_lib.OP_IDENTITY), []


class BoxInpainting(Degradation):
    """degradations.py:23-32; the mask is computed from pixel coordinates inside the kernel (utils.py:327-336)."""
    def __init__(self, half_size_mask):
        self.half_size_mask = half_size_mask

    def descriptor(self, B, C, H, W, device):
        return _op(_lib.OP_BOX, half_size=int(self.half_size_mask)), []


class _CachedMask(Degradation):
    def __init__(self):
        self._cache: Dict[Tuple, torch.Tensor] = {}

    def _host_mask(self, B, H, W) -> np.ndarray:
        raise NotImplementedError()

    def device_mask(self, B, H, W, device) -> torch.Tensor:
        key = (B, H, W, str(device))
        if key not in self._cache:
            m = np.ascontiguousarray(self._host_mask(B, H, W).astype(np.uint8))
            self._cache[key] = torch.from_numpy(m).to(device)
        return self._cache[key]

    def descriptor(self, B, C, H, W, device):
        m = self.device_mask(B, H, W, device)
        return _op(_lib.OP_MASK, mask=m.data_ptr()), [m]


class RandomInpainting(_CachedMask):
    """degradations.py:35-44 + utils.py:353-361: np.random.seed(42); binomial(1, 1-p, (B,H,W)), shared over channels.
    Image i's mask does not depend on B (sequential stream), so shards may slice a full-batch mask."""
    def __init__(self, p):
        super().__init__()
        self.p = p

    def _host_mask(self, B, H, W):
        rs = np.random.RandomState(42)                # same MT19937 stream as np.random.seed(42) + np.random.binomial
        return rs.binomial(n=1, p=1 - self.p, size=(B, H, W))


class PaintbrushInpainting(_CachedMask):
    """degradations.py:47-52 + utils.py:339-350,904-969."""
    def _host_mask(self, B, H, W):
        import cv2
        if W < 64 or H < 64:
            raise Exception("Width and Height of mask must be at least 64!")
        rng = _pyrandom.Random(42)
        size = int((W + H) * 0.08)
        out = np.zeros((B, H, W), np.uint8)
        for i in range(B):
            img = np.zeros((H, W, 1), np.uint8)
            for _ in range(10):
                x1, x2 = rng.randint(W // 2 - 30, W // 2 + 30), rng.randint(W // 2 - 30, W // 2 + 30)
                y1, y2 = rng.randint(H // 2 - 30, H // 2 + 30), rng.randint(H // 2 - 30, H // 2 + 30)
                thickness = rng.randint(8, size)
                cv2.line(img, (x1, y1), (x2, y2), (255, 255, 255), thickness)
            out[i] = (img[:, :, 0] == 0)
        return out


class GaussianDeblurring(Degradation):
    """degradations.py:55-89 (mode 'fft' = circular convolution).  The reference multiplies FFTs of a zero-padded,
    origin-centred 2-D Gaussian; that kernel is exactly outer(g, g) with g the normalised 1-D Gaussian, so the engine
    runs a separable circular convolution in shared memory (no FFT, no complex intermediates)."""
    def __init__(self, sigma_blur, kernel_size, mode="fft", num_channels=3, dim_image=128, device="cuda"):
        if mode != "fft":
            raise ValueError("only the reference's mode='fft' (circular) blur is implemented")
        self.mode, self.sigma, self.kernel_size, self.device = mode, sigma_blur, kernel_size, device
        r = torch.arange(-kernel_size // 2 + 1., kernel_size // 2 + 1.)
        g = torch.exp(-(r ** 2) / (2 * sigma_blur ** 2)).double()
        self.taps_host = (g / g.sum()).float()
        k2 = torch.exp(-(r[:, None] ** 2 + r[None, :] ** 2) / (2 * sigma_blur ** 2))
        self.kernel = (k2 / k2.sum())                  # attribute kept for API parity (degradations.py:60)
        self._taps: Dict[str, torch.Tensor] = {}
        self._scratch: Dict[Tuple, torch.Tensor] = {}

    def descriptor(self, B, C, H, W, device):
        key = str(device)
        if key not in self._taps:
            self._taps[key] = self.taps_host.to(device)
        sk = (B, C, H, W, key)
        if sk not in self._scratch:
            self._scratch = {sk: torch.empty(B * C * H * W, device=device)}
        taps, scratch = self._taps[key], self._scratch[sk]
        return _op(_lib.OP_BLUR, taps=taps.data_ptr(), ksize=int(self.kernel_size), scratch=scratch.data_ptr()), [taps, scratch]


class Superresolution(Degradation):
    """degradations.py:92-127.  mode None: s-fold decimation / zero-filled upsampling (utils.py:283-310).  mode 'bicubic'
    (:97-109,117-127; never selected by the reference's main.py:165): the reference filters with the 4 sf x 4 sf bicubic kernel
    (utils.py:365-396) through the FFT, circularly, then decimates; the engine evaluates the same separable circular filter
    directly (PNPF_OP_SR_BICUBIC: K^2 MACs per low-resolution pixel forward, 16 per pixel in the adjoint), fused with the
    residual and the gradient step in ``datafit_step``.  The reference constructor's dense (H^2/sf^2 x H^2)
    ``downsampling_matrix`` (1.07 GB at 256^2/sf 4) is never read by pnp_flow and is not built."""
    def __init__(self, sf, dim_image, mode=None, device="cuda"):
        if mode not in (None, "bicubic"):
            raise ValueError("Superresolution: mode must be None or 'bicubic' (degradations.py:93-109)")
        self.sf, self.mode, self.dim_image = sf, mode, dim_image
        self._dev: Dict[Tuple, torch.Tensor] = {}
        if mode == "bicubic":
            self.taps_host = self._bicubic_taps(sf)                          # 1-D factor: kernel = outer(taps, taps)
            self.filter = self._bicubic_filter_image(sf, dim_image)          # attribute kept for API parity (:105-109), CPU

    @staticmethod
    def _bicubic_weights(sf):
        """utils.py:386-395: the un-normalised 1-D bicubic weights on the 4*sf half-integer grid (a = -0.5), float64."""
        x = np.abs(np.arange(start=-2 * sf + 0.5, stop=2 * sf, step=1) / sf)
        a = -0.5
        w = ((a + 2) * np.power(x, 3) - (a + 3) * np.power(x, 2) + 1) * (x <= 1)
        w += (a * np.power(x, 3) - 5 * a * np.power(x, 2) + 8 * a * x - 4 * a) * (x > 1) * (x < 2)
        return w

    @classmethod
    def _bicubic_taps(cls, sf):
        w = cls._bicubic_weights(sf)
        return torch.from_numpy(w / np.sum(w)).float()                       # outer(w, w) / sum(outer) == outer(g, g)

    @classmethod
    def _bicubic_filter_image(cls, sf, dim_image):
        """utils.py:365-396 + degradations.py:97-109: the 4sf x 4sf bicubic filter zero-padded to the image, origin-centred."""
        w = cls._bicubic_weights(sf)
        w = np.outer(w, w)
        k = torch.Tensor(w / np.sum(w)).unsqueeze(0).unsqueeze(0)
        f = torch.zeros((1, 3, dim_image, dim_image))
        f[..., : k.shape[-1], : k.shape[-1]] = k
        return torch.roll(f, shifts=(-(k.shape[-1] - 1) // 2, -(k.shape[-1] - 1) // 2), dims=(2, 3))

    def descriptor(self, B, C, H, W, device):
        if self.mode is None:
            return _op(_lib.OP_SR, sf=int(self.sf)), []
        kt, ks = ("taps", str(device)), ("scratch", str(device), B * C * H * W)
        if kt not in self._dev:
            self._dev[kt] = self.taps_host.to(device)
        if ks not in self._dev:
            self._dev = {k: v for k, v in self._dev.items() if k[0] != "scratch"}
            self._dev[ks] = torch.empty(B * C * (H // self.sf) * (W // self.sf), device=device)
        taps, scratch = self._dev[kt], self._dev[ks]
        return _op(_lib.OP_SR_BICUBIC, sf=int(self.sf), taps=taps.data_ptr(), ksize=4 * int(self.sf), scratch=scratch.data_ptr()), [taps, scratch]

    def out_shape(self, B, C, H, W):
        return (B, C, H // self.sf, W // self.sf)

    def full_shape(self, B, C, Hy, Wy):
        return (B, C, Hy * self.sf, Wy * self.sf)


class _PythonOperator(Degradation):
    """Operator-API fallback for unknown Degradation subclasses: calls the plugin's own torch H / H_adj on the GPU."""
    def __init__(self, inner):
        self.inner = inner

    def H(self, x):
        return self.inner.H(x)

    def H_adj(self, y):
        return self.inner.H_adj(y)

    def datafit_step(self, x, y, gamma, out=None, noise_type='gaussian'):
        r = self.inner.H(x) - y
        if noise_type == 'laplace':
            r = 2 * torch.heaviside(r, torch.zeros_like(r)) - 1
        elif noise_type != 'gaussian':
            raise ValueError('Noise type not supported')
        z = x - gamma * self.inner.H_adj(r)
        if out is not None:
            out.copy_(z)
            return out
        return z


def as_engine_operator(obj) -> Degradation:
    """Map a degradation object (ours, the reference's, or any duck-typed plugin) to an engine operator."""
    if isinstance(obj, Degradation):
        return obj
    name = type(obj).__name__
    if name == "Denoising":
        return Denoising()
    if name == "BoxInpainting":
        return BoxInpainting(obj.half_size_mask)
    if name == "RandomInpainting":
        return RandomInpainting(obj.p)
    if name == "PaintbrushInpainting":
        return PaintbrushInpainting()
    if name == "GaussianDeblurring" and getattr(obj, "mode", None) == "fft":
        return GaussianDeblurring(obj.sigma, obj.kernel_size, "fft")
    if name == "Superresolution" and getattr(obj, "mode", "x") is None:
        return Superresolution(obj.sf, 0)
    if name == "Superresolution" and getattr(obj, "mode", None) == "bicubic":
        return Superresolution(obj.sf, int(obj.filter.shape[-1]), mode="bicubic")
    if hasattr(obj, "H") and hasattr(obj, "H_adj"):
        return _PythonOperator(obj)
    raise TypeError(f"{name} is not a degradation operator (needs H and H_adj)")
