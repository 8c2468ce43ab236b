"""Parity numbers engine vs oracle (oracle run in fp32 torch on the same GPU, TF32 off) -> gpurun_out/parity_report.json
  * teacher-forced single evaluations of the full nets (bench weight recipe and scale-1 weights)
  * 100-step loops (cfg2/cfg3/cfg4-like operators at reduced batch) : rel-L2, PSNR both sides, |dPSNR|"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import oracle
import pnpflow_b200 as P

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
rep = {"teacher_forced": [], "loops": []}
quick = "--quick" in sys.argv

def rel(a, b):
    return ((a - b).norm() / b.norm()).item()

for name, cfg, B in (("celeba128", oracle.CELEBA_128, 2), ("afhq256", oracle.AFHQ_256, 1)):
    for end_gain in (1e-3, 1.0):
        sd = oracle.init_state_dict(cfg, seed=0, end_gain=end_gain)
        sdg = {k: v.to(dev) for k, v in sd.items()}
        eng = P.UNetEngine(cfg, sd, max_batch=B)
        g = torch.Generator().manual_seed(3)
        for t0 in (0.0, 0.5, 0.95):
            x = torch.randn(B, 3, cfg.input_height, cfg.input_height, generator=g).to(dev)
            t = torch.full((B,), t0, device=dev)
            with torch.no_grad():
                ref = oracle.unet_forward(sdg, cfg, x, t)
            v = eng(x, t)
            rep["teacher_forced"].append(dict(net=name, end_gain=end_gain, t=t0, rel_l2=rel(v, ref), ref_rms=ref.pow(2).mean().sqrt().item(),
                                              max_abs=(v - ref).abs().max().item()))
            print(rep["teacher_forced"][-1], flush=True)
        del eng

loops = [("celeba128", oracle.CELEBA_128, "inpainting", 2, 100, 5), ("celeba128", oracle.CELEBA_128, "gaussian_deblurring_FFT", 2, 100, 1),
         ("celeba128", oracle.CELEBA_128, "random_inpainting", 2, 100, 1), ("afhq256", oracle.AFHQ_256, "superresolution", 1, 100, 1)]
if quick:
    loops = loops[:1]
for name, cfg, problem, B, T, S in loops:
    side = cfg.input_height
    sd = oracle.init_state_dict(cfg, seed=0)
    sdg = {k: v.to(dev) for k, v in sd.items()}
    deg_o, sigma, alpha = oracle.make_degradation(problem, side, 3, dev)
    deg_e = {"inpainting": lambda: P.BoxInpainting(20 if side == 128 else 40), "random_inpainting": lambda: P.RandomInpainting(0.7),
             "superresolution": lambda: P.Superresolution(2 if side == 128 else 4, side),
             "gaussian_deblurring_FFT": lambda: P.GaussianDeblurring(1.0 if side == 128 else 3.0, 61)}[problem]()
    from pnpflow_b200 import synth
    clean = synth.synthetic_clean(B, 3, side, 1234).to(dev)
    y = oracle.loop.synthesize_measurement(clean, deg_o.H, sigma, 0).float()
    g = torch.Generator(device=dev).manual_seed(11)
    noise = [torch.randn(B, 3, side, side, generator=g, device=dev) for _ in range(T * S)]
    t0 = time.time()
    with torch.no_grad():
        x_ref = oracle.pnp_flow_restore(lambda a, b: oracle.unet_forward(sdg, cfg, a, b), y, deg_o, sigma, steps_pnp=T,
                                        num_samples=S, alpha=alpha, noise=noise)
    torch.cuda.synchronize(); t_ref = time.time() - t0
    eng = P.UNetEngine(cfg, sd, max_batch=B * S)
    t0 = time.time()
    x = P.restore(eng, y, deg_e, sigma, steps_pnp=T, num_samples=S, alpha=alpha, noise=noise)
    torch.cuda.synchronize(); t_eng = time.time() - t0
    p_ref, p_eng = oracle.psnr(x_ref, clean), oracle.psnr(x, clean)
    rep["loops"].append(dict(net=name, problem=problem, B=B, T=T, S=S, rel_l2=rel(x, x_ref), psnr_ref=p_ref.tolist(), psnr_engine=p_eng.tolist(),
                             dpsnr_max=(p_ref - p_eng).abs().max().item(), psnr_mean_ref=p_ref.mean().item(), psnr_mean_engine=p_eng.mean().item(),
                             eager_torch_oracle_s=t_ref, engine_s=t_eng))
    print(rep["loops"][-1], flush=True)
    del eng
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)
