import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
cfg = oracle.AFHQ_256
sd = oracle.init_state_dict(cfg)
x = torch.randn(1, 3, 256, 256); t = torch.tensor([0.5])
for n in (8, 16, 32, 64, 128):
    if n > (os.cpu_count() or 1): break
    torch.set_num_threads(n)
    with torch.no_grad():
        oracle.unet_forward(sd, cfg, x, t)
        t0 = time.time(); oracle.unet_forward(sd, cfg, x, t); dt = time.time() - t0
    print(f"threads {n}: {dt:.2f} s/eval (B=1, 256^2)", flush=True)
