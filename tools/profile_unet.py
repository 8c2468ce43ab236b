"""Per-op device-time table of one U-Net evaluation (CUDA events around every op) -> gpurun_out/unet_profile_<net>_b<B>.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pnpflow_b200 as P
from pnpflow_b200 import synth

net_name = sys.argv[1] if len(sys.argv) > 1 else "afhq256"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 80
net = synth.NETS[net_name]
eng = P.UNetEngine(net, synth.random_state_dict(net), max_batch=batch)
prof = eng.profile(batch)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(prof, open(os.path.join(ROOT, "gpurun_out", f"unet_profile_{net_name}_b{batch}.json"), "w"))
tot = sum(o["ms"] for o in prof)
tc = [o for o in prof if o["kind"] == "tc"]
print(f"{net_name} batch {batch}: total {tot:.2f} ms/eval, tc {sum(o['ms'] for o in tc):.2f} ms "
      f"({sum(o['flops'] for o in tc)/sum(o['ms'] for o in tc)/1e9:.1f} TFLOP/s), ops {len(prof)}")
# group by op class
import collections, re
groups = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
for o in prof:
    n = o["name"]
    key = ("tc:" if o["kind"] == "tc" else "simt:") + re.sub(r"^(down|up)_modules\.\d+\.(\d)a_\d", r"\1.L\2", n).split(".stats")[0].rsplit(".", 1)[-1] + (".stats" if n.endswith(".stats") else "")
    lvl = re.search(r"_modules\.\d+\.(\d)[ab]_", n)
    key = (f"L{lvl.group(1)} " if lvl else "   ") + key
    g = groups[key]; g[0] += o["ms"]; g[1] += o["flops"]; g[2] += o["bytes"]; g[3] += 1
for k, g in sorted(groups.items(), key=lambda kv: -kv[1][0])[:40]:
    print(f"{k:40s} n={g[3]:3d} {g[0]:8.3f} ms  {g[1]/max(g[0],1e-9)/1e9:8.1f} TF/s  {g[2]/max(g[0],1e-9)/1e6:8.1f} GB/s")
