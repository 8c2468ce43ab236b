"""A/B harness: time one U-Net evaluation with several builds of the library on the SAME box.
usage: python tools/ab_libs.py ab/lib_a.so ab/lib_b.so ...   (each variant runs in its own process, interleaved twice)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, ctypes
sys.path.insert(0, %r)
import pnpflow_b200._lib as L
L.LIB_PATH = sys.argv[1]
probe = ctypes.CDLL(L.LIB_PATH)
for k in list(L.SYMBOLS):
    if not hasattr(probe, k): del L.SYMBOLS[k]      # older builds lack newer entry points
import torch, pnpflow_b200 as P
from pnpflow_b200 import synth
net = synth.NETS["afhq256"]; B = 80
eng = P.UNetEngine(net, synth.random_state_dict(net), max_batch=B)
x = torch.randn(B, 3, 256, 256, device="cuda"); t = torch.full((B,), 0.5, device="cuda")
for _ in range(3): eng.forward(x, t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for rep in range(5):
    e0.record()
    for _ in range(5): eng.forward(x, t)
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 5)
print("%%-28s min %%.2f  med %%.2f ms/eval" %% (sys.argv[1], min(ts), sorted(ts)[2]))
''' % ROOT
for rnd in range(2):
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, "-c", CHILD, os.path.join(ROOT, lib)], check=False)
