"""CPU oracle for the PnP-Flow hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch (fp32) restatement of the reference algorithm
(`/root/reference/pnpflow/{methods/pnp_flow,degradations,models,utils}.py`).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it; the product package `pnpflow_b200`
never does, and fails loudly when its CUDA library is missing.

Parity pin: the restatement is checked against the *unmodified reference code*
imported in the build container (`oracle/ref_shim.py`, `tests/test_oracle_vs_reference.py`)
and against golden vectors that were produced by that reference code
(`oracle/make_golden.py` -> `tests/golden/*.npz`).
"""
from .unet import UNetConfig, unet_layer_spec, unet_forward, init_state_dict, CELEBA_128, AFHQ_256, MNIST_28  # noqa: F401
from .operators import (Denoising, BoxInpainting, RandomInpainting, PaintbrushInpainting,  # noqa: F401
                        GaussianDeblurring, Superresolution, make_degradation, PROBLEMS)
from .loop import pnp_flow_restore, learning_rate, psnr  # noqa: F401
from .sampler import euler_sample  # noqa: F401,E402
