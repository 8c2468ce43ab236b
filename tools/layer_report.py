"""Per-op relative error of the engine against oracle intermediates (fp32 torch on GPU) -> gpurun_out/layer_report_<net>.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
import oracle
from oracle import unet as U
import pnpflow_b200 as P

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else "small3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = {"small3": oracle.UNetConfig(3, 32, 32, (1, 2), 1, (16,)), "mid": oracle.UNetConfig(3, 64, 32, (1, 2, 4), 2, (16,)),
       "celeba128": oracle.CELEBA_128, "afhq256": oracle.AFHQ_256}[name]
sd = oracle.init_state_dict(cfg, seed=0)
sdg = {k: v.cuda() for k, v in sd.items()}
g = torch.Generator().manual_seed(3)
x = torch.randn(B, 3, cfg.input_height, cfg.input_height, generator=g).cuda()
t = torch.rand(B, generator=g).cuda()

ref = {}
with torch.no_grad():
    temb = U.time_embedding(sdg, t, cfg)
    hs, h = [], x
    for L in oracle.unet_layer_spec(cfg):
        p = L.prefix
        if L.kind == 'conv':
            h = U._conv(h, sdg, p)
        elif L.kind == 'res':
            inp = torch.cat([h, hs.pop()], 1) if L.skip_ch else h
            a1 = U.swish(U._gn(inp, sdg, p + '.norm1')); ref[p + '.norm1'] = a1
            h1 = U._conv(a1, sdg, p + '.conv1') + F.linear(U.swish(temb), sdg[p + '.temb_proj.weight'], sdg[p + '.temb_proj.bias'])[:, :, None, None]
            ref[p + '.conv1'] = h1
            a2 = U.swish(U._gn(h1, sdg, p + '.norm2')); ref[p + '.norm2'] = a2
            h = U.res_block(sdg, p, inp, temb)
        elif L.kind == 'attn':
            hn = U._gn(h, sdg, p + '.norm'); ref[p + '.norm'] = hn
            C = h.shape[1]
            q = U._conv(hn, sdg, p + '.attn_q', padding=0) * C ** -0.5
            k = U._conv(hn, sdg, p + '.attn_k', padding=0)
            ref[p + '.qk'] = torch.cat([q, k], 1)
            h = U.self_attention(sdg, p, h)
        elif L.kind == 'down':
            h = U._conv(h, sdg, p, stride=2)
        elif L.kind == 'up':
            u = F.interpolate(h, scale_factor=2, mode='nearest'); ref[p + '.nearest2x'] = u
            h = U._conv(u, sdg, p)
        elif L.kind == 'end':
            a = U.swish(U._gn(h, sdg, p + '.0')); ref[p + '.0'] = a
            h = U._conv(a, sdg, p + '.2')
        ref[p] = h
        if L.push:
            hs.append(h)
    v_ref = h

eng = P.UNetEngine(cfg, sd, max_batch=B)
names = eng.op_names()
rows = []
for i, n in enumerate(names[:-1]):
    if n in ref:
        a = eng.debug_activation(x, t, i)
        r = ref[n]
        rel = ((a - r).norm() / r.norm()).item()
        rows.append(dict(i=i, name=n, rel=rel, rms=r.pow(2).mean().sqrt().item()))
v = eng(x, t)
rows.append(dict(i=len(names) - 1, name="v", rel=((v - v_ref).norm() / v_ref.norm()).item(), rms=v_ref.pow(2).mean().sqrt().item()))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"layer_report_{name}.json"), "w"), indent=0)
prev = 0
for r in rows:
    flag = " <<<" if r["rel"] > 2.5 * max(prev, 2e-3) else ""
    print(f"{r['i']:4d} {r['name']:50s} rel {r['rel']:.4f} rms {r['rms']:.3f}{flag}")
    prev = r["rel"]
