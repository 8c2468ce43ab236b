"""Target of the ncu captures: ONE U-Net evaluation (AFHQ 256^2 net, batch 80 = cfg4's 16 images x 5 draws), no warm-up, so
that the n-th launch of a kernel family is the n-th layer of the plan that uses it (pick with ncu -k regex:... -s N -c M)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, pnpflow_b200 as P
from pnpflow_b200 import synth
net = synth.NETS[sys.argv[1] if len(sys.argv) > 1 else "afhq256"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 80
eng = P.UNetEngine(net, synth.random_state_dict(net), max_batch=B, use_cuda_graph=False)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, 3, net["input_height"], net["input_height"], device="cuda", generator=g)
t = torch.full((B,), 0.5, device="cuda")
v = eng.forward(x, t)
torch.cuda.synchronize()
print("ok", float(v.abs().mean()))
