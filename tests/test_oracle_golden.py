"""Pin the oracle to golden vectors produced by the unmodified reference (oracle/make_golden.py)."""
import os

import numpy as np
import torch

import oracle
from oracle import make_golden as mg
from oracle import loop

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, name)).items()}


def test_unet_matches_reference_golden():
    torch.set_num_threads(1)
    for name in ("unet_mnist", "unet_small3"):
        cfg, wkw, x, t = mg.golden_inputs(name)
        sd = oracle.init_state_dict(cfg, **wkw)
        with torch.no_grad():
            v = oracle.unet_forward(sd, cfg, x, t)
        ref = _load(name + ".npz")["v_ref"]
        # same ATen ops in the same order: bit-exact on one thread; tolerance only guards BLAS variation
        assert torch.allclose(v, ref, rtol=0, atol=1e-6), (name, (v - ref).abs().max())


def test_operators_match_reference_golden():
    ref = _load("operators_64.npz")
    x = mg.operator_input()
    for name, _, octor in mg.operator_cases():
        op = octor()
        y = op.H(x)
        z = op.H_adj(y)
        assert torch.equal(y.float(), ref[name + "_H_ref"]) or torch.allclose(y.float(), ref[name + "_H_ref"], atol=1e-6), name
        assert torch.allclose(z.float(), ref[name + "_Hadj_ref"], atol=1e-6), name


def test_box_mask_known_answer():
    # the reference's only numeric unit test (pnpflow/tests/test_unit.py:14-20)
    y = oracle.BoxInpainting(32).H(torch.ones(1, 3, 128, 128))
    assert torch.equal(y[:, :, 32:64, 32:64], torch.zeros(1, 3, 32, 32))
    assert y.sum() == 3 * (128 * 128 - 64 * 64)


def test_loop_matches_reference_golden():
    torch.set_num_threads(1)
    ref = _load("loop_mnist.npz")
    clean, kw, sigma, wkw = mg.loop_mnist_inputs()
    cfg = oracle.MNIST_28
    sd = oracle.init_state_dict(cfg, **wkw)
    deg = oracle.Denoising()
    y = loop.synthesize_measurement(clean, deg.H, sigma, 0)
    assert torch.equal(y, ref["y_ref"])
    x = oracle.pnp_flow_restore(lambda a, b: oracle.unet_forward(sd, cfg, a, b), y, deg, sigma, **kw)
    assert torch.allclose(x, ref["x_ref"], atol=2e-5), (x - ref["x_ref"]).abs().max()
    assert (oracle.psnr(x, clean) - oracle.psnr(ref["x_ref"], clean)).abs().max() < 1e-3


def test_adjointness():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 64, 64, generator=g)
    for name, _, octor in mg.operator_cases():
        op = octor()
        y = torch.randn(op.H(x).shape, generator=g)
        lhs = (op.H(x) * y).sum()
        rhs = (x * op.H_adj(y)).sum()
        assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs)), (name, lhs, rhs)


def test_layer_spec_counts():
    # SURVEY §3.2 probed counts: 54 ResBlocks; 14 / 1 attention blocks; param totals
    for cfg, nattn, nparam in ((oracle.CELEBA_128, 14, 34473667), (oracle.AFHQ_256, 1, 31045827), (oracle.MNIST_28, None, 917889)):
        spec = oracle.unet_layer_spec(cfg)
        if nattn is not None:
            assert sum(L.kind == 'res' for L in spec) == 54
            assert sum(L.kind == 'attn' for L in spec) == nattn
        sd = oracle.init_state_dict(cfg)
        assert sum(v.numel() for v in sd.values()) == nparam
