"""A/B harness on ONE box: time one U-Net evaluation (AFHQ 256^2 net, batch 80 by default) under several environment settings
and/or library builds, each variant in its own process, the list run twice.  AB_NET=celeba128 AB_BATCH=320 times the 128^2 net.
usage: python tools/ab_env.py "NAME:ENV1=1,ENV2=1[,PNPF_LIB=ab/x.so]" ...      (a bare NAME: is the default build)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys
sys.path.insert(0, %r)
import torch, pnpflow_b200 as P
from pnpflow_b200 import synth
import os
net = synth.NETS[os.environ.get("AB_NET", "afhq256")]; B = int(sys.argv[2]); side = net["input_height"]
eng = P.UNetEngine(net, synth.random_state_dict(net), max_batch=B)
x = torch.randn(B, 3, side, side, device="cuda"); t = torch.full((B,), 0.5, device="cuda")
xb, tb, v, replay = eng.graphed(B)          # CUDA-graph replay = what PnPFlowSession.step runs
xb.copy_(x); tb.copy_(t)
for _ in range(3): replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for rep in range(5):
    e0.record()
    for _ in range(5): replay()
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 5)
print("%%-28s B=%%d min %%.3f  med %%.3f ms/eval  (%%.3f ms/img) launches %%d checksum %%.6e" %% (sys.argv[1], B, min(ts), sorted(ts)[2], min(ts) / B, eng.num_launches, v.double().abs().mean().item()), flush=True)
''' % ROOT
batch = os.environ.get("AB_BATCH", "80")
for rnd in range(int(os.environ.get("AB_ROUNDS", "2"))):
    for spec in sys.argv[1:]:
        name, _, envs = spec.partition(":")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            env[k] = os.path.join(ROOT, v) if k == "PNPF_LIB" else v
        subprocess.run([sys.executable, "-c", CHILD, name, batch], check=False, env=env)
