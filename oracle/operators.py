"""Oracle restatement of the degradation operators A / A^T (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/pnpflow/degradations.py:6-127 and the helpers in
/root/reference/pnpflow/utils.py:273-361 (gaussian_2d_kernel, upsample, downsample, square_mask,
paintbrush_mask, random_mask) and :904-969 (MaskGenerator), with the per-problem constants of
/root/reference/main.py:120-179 (``PROBLEMS``/``make_degradation``).

Every operator exposes the reference's duck-typed surface ``H(x)`` / ``H_adj(x)`` (degradations.py:6-12)
plus ``mask(shape)`` for the diagonal operators so tests can compare the engine's cached device masks.
"""
from __future__ import annotations

import random as _pyrandom

import numpy as np
import torch


class Degradation:
    def H(self, x):               # degradations.py:8-9
        raise NotImplementedError()

    def H_adj(self, x):           # degradations.py:11-12
        raise NotImplementedError()


class Denoising(Degradation):
    """degradations.py:15-20: identity."""
    def H(self, x):
        return x

    def H_adj(self, x):
        return x


class _Diagonal(Degradation):
    """Binary-mask operators are self-adjoint and applied as ``mask * x`` (utils.py:336,350,361)."""
    def mask(self, x):
        raise NotImplementedError()

    def H(self, x):
        return self.mask(x) * x

    H_adj = H


class BoxInpainting(_Diagonal):
    """degradations.py:23-32 + utils.py:327-336: zero the centre square [d-h, d+h)^2, d = H//2."""
    def __init__(self, half_size_mask):
        self.half_size_mask = half_size_mask

    def mask(self, x):
        d, h = x.shape[2] // 2, self.half_size_mask
        m = torch.ones_like(x)
        m[:, :, d - h:d + h, d - h:d + h] = 0
        return m


class RandomInpainting(_Diagonal):
    """degradations.py:35-44 + utils.py:353-361: numpy legacy RNG re-seeded to 42 on EVERY call,
    Bernoulli(1-p) keep mask of shape (B,H,W) shared over channels, dtype int64 (promotes with x)."""
    def __init__(self, p):
        self.p = p

    def mask(self, x):
        np.random.seed(42)
        m = np.random.binomial(n=1, p=1 - self.p, size=(x.shape[0], x.shape[2], x.shape[3]))
        return torch.from_numpy(m).to(x.device).unsqueeze(1)


class PaintbrushInpainting(_Diagonal):
    """degradations.py:47-52 + utils.py:339-350 + MaskGenerator utils.py:904-969 (python `random`
    seeded with 42 per call, ten cv2 lines per image near the centre, keep = not painted)."""
    def mask(self, x):
        import cv2
        B, _, Hh, Ww = x.shape
        if Ww < 64 or Hh < 64:
            raise Exception("Width and Height of mask must be at least 64!")      # utils.py:928-929
        rng = _pyrandom.Random(42)                      # == random.seed(42) stream (utils.py:920-921)
        size = int((Ww + Hh) * 0.08)
        m = torch.zeros_like(x)
        for i in range(B):
            img = np.zeros((Hh, Ww, 1), np.uint8)
            for _ in range(10):                         # utils.py:932-938: draw order x1,x2,y1,y2,thickness
                x1, x2 = rng.randint(Ww // 2 - 30, Ww // 2 + 30), rng.randint(Ww // 2 - 30, Ww // 2 + 30)
                y1, y2 = rng.randint(Hh // 2 - 30, Hh // 2 + 30), rng.randint(Hh // 2 - 30, Hh // 2 + 30)
                thickness = rng.randint(8, size)
                cv2.line(img, (x1, y1), (x2, y2), (255, 255, 255), thickness)
            keep = torch.from_numpy(img[:, :, 0] == 0).to(x.device)   # (1-img)-1 == 0  <=>  img == 0
            m[i] = keep
        return m


def gaussian_kernel_1d(sigma: float, size: int) -> torch.Tensor:
    """The reference kernel (utils.py:273-280) is exp(-(x^2+y^2)/2s^2)/sum: exactly outer(g,g) with this g."""
    r = torch.arange(-size // 2 + 1., size // 2 + 1.)
    g = torch.exp(-(r ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def gaussian_kernel_2d(sigma: float, size: int) -> torch.Tensor:
    """utils.py:273-280 verbatim in arithmetic (2-D exp then normalise by the 2-D sum)."""
    r = torch.arange(-size // 2 + 1., size // 2 + 1.)
    xx, yy = torch.meshgrid(r, r, indexing='ij')
    k = torch.exp(-(xx ** 2 + yy ** 2) / (2 * sigma ** 2))
    return k / k.sum()


class GaussianDeblurring(Degradation):
    """degradations.py:55-89, mode 'fft': circular convolution; the 61x61 kernel is zero-padded to
    HxW and rolled by -(k-1)//2 so that its centre sits at the origin (:62-68)."""
    def __init__(self, sigma_blur, kernel_size, mode="fft", num_channels=3, dim_image=128, device="cpu"):
        assert mode == "fft"
        self.sigma, self.kernel_size, self.mode, self.device = sigma_blur, kernel_size, mode, device
        self.kernel = gaussian_kernel_2d(sigma_blur, kernel_size).to(device)
        f = torch.zeros((1, num_channels, dim_image, dim_image), device=device)
        f[..., :kernel_size, :kernel_size] = self.kernel
        s = -(kernel_size - 1) // 2
        self.filter = torch.roll(f, shifts=(s, s), dims=(2, 3))

    def H(self, x):
        return torch.real(torch.fft.ifft2(torch.fft.fft2(x.to(self.device)) * torch.fft.fft2(self.filter)))

    def H_adj(self, x):
        return torch.real(torch.fft.ifft2(torch.fft.fft2(x.to(self.device)) * torch.conj(torch.fft.fft2(self.filter))))


def bicubic_filter(factor: int = 2) -> torch.Tensor:
    """utils.py:365-396 (a = -0.5 cubic kernel sampled at 4*factor points, outer product, normalised), [1,1,4f,4f]."""
    x = np.arange(start=-2 * factor + 0.5, stop=2 * factor, step=1) / factor
    a = -0.5
    x = np.abs(x)
    w = ((a + 2) * np.power(x, 3) - (a + 3) * np.power(x, 2) + 1) * (x <= 1)
    w += (a * np.power(x, 3) - 5 * a * np.power(x, 2) + 8 * a * x - 4 * a) * (x > 1) * (x < 2)
    w = np.outer(w, w)
    w = w / np.sum(w)
    return torch.Tensor(w).unsqueeze(0).unsqueeze(0)


class Superresolution(Degradation):
    """degradations.py:92-127.  mode None: keep the upper-left pixel of every sf x sf patch (utils.py:302-310); adjoint =
    zero-filled upsampling (utils.py:283-299).  mode 'bicubic' (:97-109,117-127): circular convolution with the 4sf x 4sf
    bicubic filter (zero-padded to the image, rolled by -(4sf-1)//2) through the FFT, then decimation; adjoint = zero-filled
    upsampling, then the conjugate filter.  The reference constructor also materialises a dense (H^2/sf^2) x H^2 one-hot
    matrix (:110-111) that pnp_flow never reads; it is not restated."""
    def __init__(self, sf, dim_image, mode=None, device="cpu"):
        assert mode in (None, "bicubic")
        self.sf, self.mode = sf, mode
        if mode == "bicubic":
            k = bicubic_filter(sf).to(device)
            f = torch.zeros((1, 3) + (dim_image, dim_image), device=device)
            f[..., : k.shape[-1], : k.shape[-1]] = k
            self.filter = torch.roll(f, shifts=(-(k.shape[-1] - 1) // 2, -(k.shape[-1] - 1) // 2), dims=(2, 3))

    def _down(self, x):
        return x[..., 0::self.sf, 0::self.sf]

    def _up(self, x):
        z = torch.zeros((x.shape[0], x.shape[1], x.shape[2] * self.sf, x.shape[3] * self.sf)).type_as(x)
        z[..., 0::self.sf, 0::self.sf].copy_(x)
        return z

    def H(self, x):
        if self.mode is None:
            return self._down(x)
        return self._down(torch.real(torch.fft.ifft2(torch.fft.fft2(x) * torch.fft.fft2(self.filter))))

    def H_adj(self, x):
        if self.mode is None:
            return self._up(x)
        return torch.real(torch.fft.ifft2(torch.fft.fft2(self._up(x)) * torch.conj(torch.fft.fft2(self.filter))))


# problem name -> (constructor(dim, channels, device), sigma_noise gaussian, tuned alpha)   main.py:120-179,
# alpha from scripts/script_test.sh:10-20
PROBLEMS = {
    'denoising': (lambda d, c, dev: Denoising(), 0.2, 0.8),
    'inpainting': (lambda d, c, dev: BoxInpainting(20 if d == 128 else 40), 0.05, 0.5),
    'paintbrush_inpainting': (lambda d, c, dev: PaintbrushInpainting(), 0.05, 0.5),
    'random_inpainting': (lambda d, c, dev: RandomInpainting(0.7), 0.01, 0.01),
    'superresolution': (lambda d, c, dev: Superresolution(2 if d == 128 else 4, d, device=dev), 0.05, 0.3),
    'gaussian_deblurring_FFT': (lambda d, c, dev: GaussianDeblurring(1.0 if d == 128 else 3.0, 61, "fft", c, d, dev), 0.05, 0.01),
}


def make_degradation(problem: str, dim_image: int, num_channels: int = 3, device="cpu"):
    """Returns (degradation, sigma_noise, alpha) for a reference problem name (main.py:120-179)."""
    ctor, sigma, alpha = PROBLEMS[problem]
    return ctor(dim_image, num_channels, device), sigma, alpha
