"""BASELINE-size parity of the restoration loop: the full 128^2 / 256^2 nets, T = 100 steps, S = 5 draws, B = 2 images,
engine vs the fp32 oracle run on the same GPU (eager torch, TF32 off) on identical y / weights / noise.

The tolerance is the north star's: |dPSNR| < 0.01 dB PER IMAGE (BASELINE.json: "PSNR equal to 2 decimals").  This is what
licenses the engine's 16-bit operand / activation storage (fp16 since round 2).  Cases:
  * cfg2-like  CelebA 128^2 box inpainting (half 20, sigma .05, alpha .5) — SEEDED path: ``noise=None`` on both sides, the
    engine must consume the same Philox stream as the reference's ``torch.randn_like`` calls (pnp_flow.py:48)
  * cfg3-like  CelebA 128^2 Gaussian deblurring (sigma_b 1, k 61) — through the plugin: ``PNP_FLOW.solve_ip`` (measurement
    synthesis :77-80 + loop) against ``oracle.loop.synthesize_measurement`` + ``pnp_flow_restore``
  * cfg4       AFHQ 256^2 SR x4, S = 5, injected noise
  * cfg5-like  AFHQ 256^2 random inpainting p = .7, S = 5, T reduced 200 -> 100, injected noise
  * the loop at larger velocity magnitudes (end_conv gain 1e-2 / 3e-2 instead of the recipe's 1e-3) next to the reference's own
    TF32 spread: reported as measured (gpurun_out/parity_baseline.json).
Every case appends its numbers to gpurun_out/parity_baseline.json (copied to profiles/ by the builder).
"""
import json
import os

import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_DB = 0.01


def _record(entry):
    path = os.path.join(ROOT, "gpurun_out", "parity_baseline.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    data = []
    if os.path.exists(path):
        try:
            data = json.load(open(path))
        except Exception:
            data = []
    data.append(entry)
    json.dump(data, open(path, "w"), indent=1)


def _engine_op(problem, side):
    import pnpflow_b200 as P
    return {"inpainting": lambda: P.BoxInpainting(20 if side == 128 else 40),
            "random_inpainting": lambda: P.RandomInpainting(0.7),
            "superresolution": lambda: P.Superresolution(2 if side == 128 else 4, side),
            "gaussian_deblurring_FFT": lambda: P.GaussianDeblurring(1.0 if side == 128 else 3.0, 61, "fft", 3, side, "cuda")}[problem]()


def _setup(cfg, problem, B, seed_clean, end_gain=1e-3):
    from pnpflow_b200 import synth
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    side = cfg.input_height
    sd = oracle.init_state_dict(cfg, seed=0, end_gain=end_gain)
    sdg = {k: v.to(dev) for k, v in sd.items()}
    deg_o, sigma, alpha = oracle.make_degradation(problem, side, 3, dev)
    clean = synth.synthetic_clean(B, 3, side, seed_clean).to(dev)
    model = lambda a, b: oracle.unet_forward(sdg, cfg, a, b)           # noqa: E731
    return dev, side, sd, model, deg_o, sigma, alpha, clean


def _compare(tag, x, x_ref, clean, extra=None, tol=TOL_DB):
    p_ref, p_eng = oracle.psnr(x_ref, clean), oracle.psnr(x, clean)
    d = (p_ref - p_eng).abs()
    rel = ((x - x_ref).norm() / x_ref.norm()).item()
    e = dict(case=tag, rel_l2=rel, psnr_ref=p_ref.tolist(), psnr_engine=p_eng.tolist(), dpsnr_max=d.max().item(),
             max_abs=(x - x_ref).abs().max().item())
    e.update(extra or {})
    _record(e)
    print(e)
    assert torch.isfinite(x).all()
    assert d.max().item() < tol, e
    return e


@pytest.mark.parametrize("name,problem,seeded", [("celeba128", "inpainting", True), ("afhq256", "superresolution", False),
                                                 ("afhq256", "random_inpainting", False)])
def test_100_step_loop_vs_oracle(name, problem, seeded):
    import pnpflow_b200 as P
    cfg = oracle.CELEBA_128 if name == "celeba128" else oracle.AFHQ_256
    B, T, S = 2, 100, 5
    dev, side, sd, model, deg_o, sigma, alpha, clean = _setup(cfg, problem, B, 1234)
    y = oracle.loop.synthesize_measurement(clean, deg_o.H, sigma, 0).float()
    eng = P.UNetEngine(cfg, sd, max_batch=B * S)
    if seeded:
        # identical seeds: both sides draw torch.randn_like from the global CUDA generator, one call per draw
        torch.manual_seed(4321)
        x_ref = oracle.pnp_flow_restore(model, y, deg_o, sigma, steps_pnp=T, num_samples=S, alpha=alpha)
        torch.manual_seed(4321)
        x = P.restore(eng, y, _engine_op(problem, side), sigma, steps_pnp=T, num_samples=S, alpha=alpha)
    else:
        g = torch.Generator(device=dev).manual_seed(11)
        noise = [torch.randn(B, 3, side, side, generator=g, device=dev) for _ in range(T * S)]
        x_ref = oracle.pnp_flow_restore(model, y, deg_o, sigma, steps_pnp=T, num_samples=S, alpha=alpha, noise=noise)
        x = P.restore(eng, y, _engine_op(problem, side), sigma, steps_pnp=T, num_samples=S, alpha=alpha, noise=noise)
    _compare(f"{name}/{problem}/T{T}/S{S}/B{B}/{'seeded' if seeded else 'injected'}", x, x_ref, clean)


def test_solve_ip_plugin_vs_oracle_cfg3():
    """Row a11: the drop-in call.  PNP_FLOW(model, device, args).solve_ip(loader, degradation, sigma) — measurement synthesis
    with torch.manual_seed(batch) (pnp_flow.py:77-80), then the loop drawing from the same global generator — against the
    oracle's synthesize_measurement + pnp_flow_restore.  y must agree to the blur kernel's tolerance, x to 0.01 dB."""
    import pnpflow_b200 as P
    from oracle.ref_shim import RefArgs
    cfg = oracle.CELEBA_128
    B, T, S = 2, 100, 5
    problem = "gaussian_deblurring_FFT"
    dev, side, sd, model, deg_o, sigma, alpha, clean = _setup(cfg, problem, B, 1237)
    args = RefArgs(steps_pnp=T, num_samples=S, alpha=alpha, dim_image=side, max_batch=2, lr_pnp=1.0,
                   save_path='/tmp/pnpflow_b200_test')
    eng = P.UNetEngine(cfg, sd, max_batch=B * S)
    m = P.PNP_FLOW(eng, dev, args)
    loader = [(clean.cpu(), torch.zeros(B)), (clean.flip(0).cpu(), torch.zeros(B))]
    res = m.solve_ip(loader, _engine_op(problem, side), sigma)
    assert len(res) == 2
    for batch, (clean_b, y_eng, x_eng) in enumerate(res):
        y_ref = oracle.loop.synthesize_measurement(clean_b.to(dev), deg_o.H, sigma, batch).float()
        x_ref = oracle.pnp_flow_restore(model, y_ref, deg_o, sigma, steps_pnp=T, num_samples=S, alpha=alpha)
        dy = (y_eng.to(dev) - y_ref).abs().max().item()
        assert dy <= 4e-6, dy                        # separable fp32 blur vs the oracle's FFT; the noise term is bit-identical
        _compare(f"celeba128/solve_ip/{problem}/batch{batch}", x_eng.to(dev), x_ref, clean_b.to(dev), dict(dy_max=dy))


@pytest.mark.parametrize("end_gain", [1e-2, 3e-2])
def test_loop_at_larger_velocity_magnitude(end_gain):
    """The recipe's end_conv gain 1e-3 gives rms|v| ~ 0.05; a trained net has |v| ~ 1.  With random weights a larger gain makes
    the 100-step loop CHAOTIC: any perturbation of an evaluation is amplified step after step, so "engine
    vs fp32 oracle" has to be read next to the reference's OWN numeric spread.  Three runs on identical y / weights / noise:
    the fp32 oracle (TF32 off), the oracle with cuDNN TF32 allowed (= the reference's default GPU numerics), and
    the engine.  Reported as measured (gpurun_out/parity_baseline.json).  With fp16 operands (round 2) the engine stays inside
    the reference's own TF32 spread: the 0.01 dB bar is asserted at gain 1e-2 as well, the chaotic 3e-2 case for sanity only."""
    import pnpflow_b200 as P
    cfg = oracle.CELEBA_128
    B, T, S = 2, 100, 2
    problem = "inpainting"
    dev, side, sd, model, deg_o, sigma, alpha, clean = _setup(cfg, problem, B, 1234, end_gain=end_gain)
    y = oracle.loop.synthesize_measurement(clean, deg_o.H, sigma, 0).float()
    g = torch.Generator(device=dev).manual_seed(11)
    noise = [torch.randn(B, 3, side, side, generator=g, device=dev) for _ in range(T * S)]
    vr = []
    def model_rec(a, b):
        v = model(a, b)
        vr.append(v.pow(2).mean().sqrt().item())
        return v
    x_ref = oracle.pnp_flow_restore(model_rec, y, deg_o, sigma, steps_pnp=T, num_samples=S, alpha=alpha, noise=noise)
    torch.backends.cudnn.allow_tf32 = True               # the reference's own GPU path
    x_tf32 = oracle.pnp_flow_restore(model, y, deg_o, sigma, steps_pnp=T, num_samples=S, alpha=alpha, noise=noise)
    torch.backends.cudnn.allow_tf32 = False
    eng = P.UNetEngine(cfg, sd, max_batch=B * S)
    x = P.restore(eng, y, _engine_op(problem, side), sigma, steps_pnp=T, num_samples=S, alpha=alpha, noise=noise)
    p_ref = oracle.psnr(x_ref, clean)
    d_tf32 = (oracle.psnr(x_tf32, clean) - p_ref).abs().max().item()
    rel_tf32 = ((x_tf32 - x_ref).norm() / x_ref.norm()).item()
    # gain 1e-2 (rms|v| ~ 0.2): still well-posed -> the 0.01 dB bar holds (measured 7e-4 dB, the reference's own TF32 path 2.3e-3);
    # gain 3e-2 (rms|v| ~ 0.4): chaotic, the reference's TF32 path itself ends 0.3 dB from its fp32 run -> sanity bound only
    _compare(f"celeba128/{problem}/end_gain{end_gain}", x, x_ref, clean,
             dict(rms_v_first=vr[0], rms_v_last=vr[-1], reference_tf32_vs_fp32_dpsnr_max=d_tf32, reference_tf32_vs_fp32_rel_l2=rel_tf32,
                  note="compare dpsnr_max with reference_tf32_vs_fp32_dpsnr_max"), tol=TOL_DB if end_gain <= 1e-2 else 3.0)
