"""Experiment: one U-Net evaluation of B images as ONE engine call vs TWO half-batch engines on two CUDA streams
(HBM-bound GroupNorm / layout kernels of one half can overlap the tensor-core kernels of the other)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, pnpflow_b200 as P
from pnpflow_b200 import synth
net = synth.NETS["afhq256"]; B = int(sys.argv[1]) if len(sys.argv) > 1 else 80
sd = synth.random_state_dict(net)
x = torch.randn(B, 3, 256, 256, device="cuda"); t = torch.full((B,), 0.5, device="cuda")
def timeit(fn, n=5, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / n)
    return min(ts), sorted(ts)[len(ts) // 2]
eng = P.UNetEngine(net, sd, max_batch=B)
xb, tb, vb, replay = eng.graphed(B)
print("one engine  B=%d eager  min %.3f med %.3f ms" % ((B,) + timeit(lambda: eng.forward(x, t))), flush=True)
print("one engine  B=%d graph  min %.3f med %.3f ms" % ((B,) + timeit(replay)), flush=True)
del eng, xb, tb, vb, replay
torch.cuda.empty_cache()
for parts in (2, 4):
    h = B // parts
    engs = [P.UNetEngine(net, sd, max_batch=h) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    graphs = []
    for e, s in zip(engs, streams):
        with torch.cuda.stream(s):
            graphs.append(e.graphed(h))
    torch.cuda.synchronize()
    def run_parts():
        cur = torch.cuda.current_stream()
        for (xb_, tb_, vb_, rep), s in zip(graphs, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                rep()
        for s in streams:
            cur.wait_stream(s)
    print("%d engines x B=%d graphs on %d streams  min %.3f med %.3f ms" % ((parts, h, parts) + timeit(run_parts)), flush=True)
    del engs, graphs
    torch.cuda.empty_cache()
