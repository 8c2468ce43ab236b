"""Differential tests of the oracle against the unmodified reference code (build container only)."""
import pytest
import torch

import oracle
from oracle import loop, ref_shim
from oracle import make_golden as mg

pytestmark = pytest.mark.reference


def test_unet_bit_exact_vs_reference_module():
    cfg = mg.SMALL3
    sd = oracle.init_state_dict(cfg, seed=9, perturb=0.2)
    net = ref_shim.build_reference_unet(cfg, sd)
    g = torch.Generator().manual_seed(1)
    x, t = torch.randn(3, 3, 32, 32, generator=g), torch.tensor([0.1, 0.5, 0.99])
    with torch.no_grad():
        assert torch.equal(net(x, t), oracle.unet_forward(sd, cfg, x, t))


def test_state_dict_keys_and_shapes_match_reference():
    _, _, _, N = ref_shim.load()
    for cfg in (oracle.MNIST_28, mg.SMALL3, oracle.CELEBA_128):
        ref = N.UNet(cfg.input_channels, cfg.input_height, cfg.ch, ch_mult=cfg.ch_mult,
                     num_res_blocks=cfg.num_res_blocks, attn_resolutions=cfg.attn_resolutions).state_dict()
        sd = oracle.init_state_dict(cfg)
        assert set(ref) == set(sd)
        for k in ref:
            assert tuple(ref[k].shape) == tuple(sd[k].shape), k


def test_operators_bit_exact_vs_reference():
    _, D, _, _ = ref_shim.load()
    x = mg.operator_input()
    for name, rctor, octor in mg.operator_cases():
        r, o = rctor(D), octor()
        yr, yo = r.H(x), o.H(x)
        assert torch.equal(yr, yo), name
        assert torch.equal(r.H_adj(yr), o.H_adj(yo)), name


@pytest.mark.parametrize("problem,alpha", [("box", 0.5), ("random", 0.01), ("sr2", 0.3), ("blur", 0.01), ("paintbrush", 0.5)])
def test_loop_bit_exact_vs_reference_solve_ip(problem, alpha):
    """Full solve_ip (measurement synthesis + 10 steps x 2 draws) on a tiny 64x64 net, global RNG shared."""
    _, D, _, _ = ref_shim.load()
    cfg = oracle.UNetConfig(3, 64, 32, (1, 2), 1, (16,))
    sd = oracle.init_state_dict(cfg, seed=2)
    net = ref_shim.build_reference_unet(cfg, sd)
    case = {n: (r, o) for n, r, o in mg.operator_cases()}[problem]
    g = torch.Generator().manual_seed(77)
    clean = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    args = ref_shim.RefArgs(steps_pnp=10, num_samples=2, alpha=alpha, dim_image=64)
    (y_ref, x_ref), = ref_shim.run_reference_solve_ip(net, [clean], case[0](D), 0.05, args)
    deg = case[1]()
    y = loop.synthesize_measurement(clean, deg.H, 0.05, 0)
    x = oracle.pnp_flow_restore(lambda a, b: oracle.unet_forward(sd, cfg, a, b), y, deg, 0.05,
                                steps_pnp=10, num_samples=2, alpha=alpha)
    assert torch.equal(y, y_ref)
    assert torch.equal(x, x_ref), (x - x_ref).abs().max()


@pytest.mark.parametrize("problem", ["box", "sr2", "blur"])
def test_laplace_loop_bit_exact_vs_reference_solve_ip(problem):
    """noise_type='laplace' (pnp_flow.py:42-43,64-66,81-85): measurement synthesis + 10 steps x 2 draws."""
    _, D, _, _ = ref_shim.load()
    cfg = oracle.UNetConfig(3, 64, 32, (1, 2), 1, (16,))
    sd = oracle.init_state_dict(cfg, seed=2)
    net = ref_shim.build_reference_unet(cfg, sd)
    case = {n: (r, o) for n, r, o in mg.operator_cases()}[problem]
    g = torch.Generator().manual_seed(78)
    clean = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    args = ref_shim.RefArgs(steps_pnp=10, num_samples=2, alpha=0.5, dim_image=64, noise_type='laplace', lr_pnp=0.05)
    torch.manual_seed(4321)
    (y_ref, x_ref), = ref_shim.run_reference_solve_ip(net, [clean], case[0](D), 0.05, args)
    deg = case[1]()
    torch.manual_seed(4321)
    y = loop.synthesize_measurement(clean, deg.H, 0.05, 0, 'laplace')
    x = oracle.pnp_flow_restore(lambda a, b: oracle.unet_forward(sd, cfg, a, b), y, deg, 0.05, steps_pnp=10,
                                num_samples=2, alpha=0.5, lr_pnp=0.05, noise_type='laplace')
    assert torch.equal(y, y_ref)
    assert torch.equal(x, x_ref), (x - x_ref).abs().max()
