"""Generate tests/golden/*.npz FROM THE UNMODIFIED REFERENCE CODE (run in the build container only).

    python -m oracle.make_golden

Every array named ``*_ref`` below is an output of code imported from /root/reference (through
oracle/ref_shim.py); inputs and weights are regenerated from seeds by the tests (oracle.init_state_dict,
torch.Generator), so the fixtures stay small.  The GPU box has no /root/reference: there the oracle is
pinned by these files (tests/test_oracle_golden.py) and the engine by the oracle.
"""
from __future__ import annotations

import os

import numpy as np
import torch

import oracle
from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SMALL3 = oracle.UNetConfig(3, 32, 32, (1, 2), 1, (16,))      # tiny 3-channel net with one attention level


def golden_inputs(name):
    """Seeded inputs shared by this generator and the tests."""
    if name == 'unet_mnist':
        g = torch.Generator().manual_seed(11)
        return oracle.MNIST_28, dict(seed=3, perturb=0.1), torch.randn(2, 1, 28, 28, generator=g), torch.tensor([0.15, 0.8])
    if name == 'unet_small3':
        g = torch.Generator().manual_seed(12)
        return SMALL3, dict(seed=4, perturb=0.1), torch.randn(2, 3, 32, 32, generator=g), torch.tensor([0.0, 0.55])
    raise KeyError(name)


def operator_cases():
    """(name, reference-ctor(D), oracle-ctor) at 64x64x3, the smallest size all six operators accept."""
    return [
        ('denoising', lambda D: D.Denoising(), lambda: oracle.Denoising()),
        ('box', lambda D: D.BoxInpainting(10), lambda: oracle.BoxInpainting(10)),
        ('random', lambda D: D.RandomInpainting(0.7), lambda: oracle.RandomInpainting(0.7)),
        ('paintbrush', lambda D: D.PaintbrushInpainting(), lambda: oracle.PaintbrushInpainting()),
        ('blur', lambda D: D.GaussianDeblurring(1.0, 61, "fft", 3, 64, "cpu"),
         lambda: oracle.GaussianDeblurring(1.0, 61, "fft", 3, 64, "cpu")),
        ('sr2', lambda D: D.Superresolution(2, 64, device="cpu"), lambda: oracle.Superresolution(2, 64)),
        ('sr2_bicubic', lambda D: D.Superresolution(2, 64, mode="bicubic", device="cpu"), lambda: oracle.Superresolution(2, 64, mode="bicubic")),
    ]


def operator_input():
    g = torch.Generator().manual_seed(13)
    return torch.randn(2, 3, 64, 64, generator=g)


def loop_mnist_inputs():
    g = torch.Generator().manual_seed(1234 + 1)              # manual_seed(1234 + cfg_id)
    clean = torch.rand(4, 1, 28, 28, generator=g) * 2 - 1
    return clean, dict(steps_pnp=20, num_samples=5, alpha=0.8, lr_pnp=1.0), 0.1, dict(seed=0)


def main():
    os.makedirs(OUT, exist_ok=True)
    U, D, M, N = ref_shim.load()
    torch.set_num_threads(1)                                 # fixed reduction order for the fixtures

    for name in ('unet_mnist', 'unet_small3'):
        cfg, wkw, x, t = golden_inputs(name)
        net = ref_shim.build_reference_unet(cfg, oracle.init_state_dict(cfg, **wkw))
        with torch.no_grad():
            v = net(x, t)
        np.savez(os.path.join(OUT, name + '.npz'), v_ref=v.numpy())
        print(name, v.shape, float(v.abs().mean()))

    x = operator_input()
    out = {}
    for name, rctor, _ in operator_cases():
        op = rctor(D)
        y = op.H(x)
        out[name + '_H_ref'] = y.to(torch.float32).numpy()
        out[name + '_Hadj_ref'] = op.H_adj(y).to(torch.float32).numpy()
    np.savez(os.path.join(OUT, 'operators_64.npz'), **out)
    print('operators', list(out))

    clean, kw, sigma, wkw = loop_mnist_inputs()
    cfg = oracle.MNIST_28
    net = ref_shim.build_reference_unet(cfg, oracle.init_state_dict(cfg, **wkw))
    args = ref_shim.RefArgs(num_channels=1, dim_image=28, dataset='mnist', **kw)
    (y_ref, x_ref), = ref_shim.run_reference_solve_ip(net, [clean], D.Denoising(), sigma, args)
    np.savez(os.path.join(OUT, 'loop_mnist.npz'), y_ref=y_ref.numpy(), x_ref=x_ref.numpy())
    print('loop_mnist psnr', oracle.psnr(x_ref, clean).tolist())


if __name__ == '__main__':
    main()
