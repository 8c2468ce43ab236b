"""GPU parity of Superresolution(mode='bicubic') (degradations.py:97-109,117-127) against the reference golden vectors and the
oracle.  The reference filters through the FFT; the engine evaluates the same separable circular 4*sf-tap filter directly
(PNPF_OP_SR_BICUBIC kernels), so the tolerances are those of an fp32 FFT vs an fp32 direct sum."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import make_golden as mg

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_bicubic_sr_matches_reference_golden():
    import pnpflow_b200 as P
    ref = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "operators_64.npz")).items()}
    x = mg.operator_input().cuda()
    op = P.Superresolution(2, 64, mode="bicubic")
    y = op.H(x)
    z = op.H_adj(y)
    assert y.shape == (2, 3, 32, 32) and z.shape == (2, 3, 64, 64)
    assert (y.cpu() - ref["sr2_bicubic_H_ref"]).abs().max() <= 4e-6
    assert (z.cpu() - ref["sr2_bicubic_Hadj_ref"]).abs().max() <= 4e-6


@pytest.mark.parametrize("sf,side", [(4, 256), (2, 128), (4, 64)])
def test_bicubic_sr_kernels_vs_oracle_fft(sf, side):
    """main.py:120-179 sizes (sf 2 at 128^2, sf 4 at 256^2): H, H_adj and both data terms against the oracle's FFT formulation."""
    import pnpflow_b200 as P
    eng_op, orc_op = P.Superresolution(sf, side, mode="bicubic"), oracle.Superresolution(sf, side, mode="bicubic", device="cuda")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, side, side, generator=g).cuda()
    u = torch.randn(2, 3, side // sf, side // sf, generator=g).cuda()
    assert (eng_op.H(x) - orc_op.H(x)).abs().max() <= 4e-6
    assert (eng_op.H_adj(u) - orc_op.H_adj(u)).abs().max() <= 4e-6
    z = eng_op.datafit_step(x, u, 0.7)
    assert (z - (x - 0.7 * orc_op.H_adj(orc_op.H(x) - u))).abs().max() <= 1e-5
    zl = eng_op.datafit_step(x, u, 0.7, noise_type='laplace')
    r = orc_op.H(x) - u
    ref = x - 0.7 * orc_op.H_adj(2 * torch.heaviside(r, torch.zeros_like(r)) - 1)
    bad = (zl - ref).abs() > 1e-5          # sign(Hx - y) may flip where |Hx - y| is at the fp32 noise of FFT vs direct sum
    assert (r.abs() < 1e-5).sum() <= 8 and bad.float().mean() < 0.02


def test_bicubic_sr_datafit_and_loop_vs_oracle():
    import pnpflow_b200 as P
    eng_op, orc_op = P.Superresolution(2, 64, mode="bicubic"), oracle.Superresolution(2, 64, mode="bicubic")
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 3, 64, 64, generator=g)
    sigma, lr_pnp, alpha, t = 0.05, 1.0, 0.3, 0.37
    y = orc_op.H(x) + 0.05 * torch.randn(orc_op.H(x).shape, generator=g)
    t1 = torch.ones(2) * t
    lr_t = oracle.learning_rate(sigma ** 2 * lr_pnp, t1, 'alpha_1_minus_t', alpha)
    z_ref = x - lr_t * oracle.loop.grad_datafit(x, y, orc_op.H, orc_op.H_adj, sigma)
    z = eng_op.datafit_step(x.cuda(), y.float().cuda(), P.gamma_schedule(lr_pnp, t, 'alpha_1_minus_t', alpha))
    assert (z.cpu() - z_ref).abs().max() < 1e-5
    # adjointness <Hx, y> = <x, H^T y> on the engine operator
    u = torch.randn(2, 3, 32, 32, generator=g).cuda()
    lhs, rhs = (eng_op.H(x.cuda()) * u).sum().item(), (x.cuda() * eng_op.H_adj(u)).sum().item()
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))
    # the reference's own operator object is recognised
    from pnpflow_b200.degradations import as_engine_operator

    class Superresolution:                         # duck-typed stand-in with the reference's attributes
        def __init__(self, o):
            self.sf, self.mode, self.filter = o.sf, o.mode, o.filter
        H = H_adj = None
    assert as_engine_operator(Superresolution(orc_op)).mode == "bicubic"
    # 10 steps x 2 draws on a small net
    cfg = oracle.UNetConfig(3, 64, 32, (1, 2), 1, (16,))
    sd = oracle.init_state_dict(cfg, seed=2)
    clean = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    y = oracle.loop.synthesize_measurement(clean, orc_op.H, 0.05, 0).float()
    T, S = 10, 2
    noise = [torch.randn(2, 3, 64, 64, generator=g) for _ in range(T * S)]
    x_ref = oracle.pnp_flow_restore(lambda a, b: oracle.unet_forward(sd, cfg, a, b), y, orc_op, 0.05,
                                    steps_pnp=T, num_samples=S, alpha=alpha, noise=noise)
    eng = P.UNetEngine(cfg, sd, max_batch=2 * S)
    xe = P.restore(eng, y.cuda(), eng_op, 0.05, steps_pnp=T, num_samples=S, alpha=alpha, noise=[n.cuda() for n in noise]).cpu()
    rel = ((xe - x_ref).norm() / x_ref.norm()).item()
    dpsnr = (oracle.psnr(xe, clean) - oracle.psnr(x_ref, clean)).abs().max().item()
    assert rel < 3e-2 and dpsnr < 0.01, (rel, dpsnr)
