"""Target of the ncu captures of the per-pixel PnP kernels: the data-fidelity step of the four BASELINE
operators, the interpolation and the push+average at cfg4's sizes (16 images of 256^2 x 3, S = 5), inputs far apart in memory."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, pnpflow_b200 as P
from pnpflow_b200 import _lib
lib = _lib.load()
B, S, side = 16, 5, 256
n = B * 3 * side * side
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, 3, side, side, device="cuda", generator=g)
z = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for op in (P.Superresolution(4, side), P.RandomInpainting(0.7), P.BoxInpainting(40), P.GaussianDeblurring(3.0, 61, "fft", 3, side),
           P.Superresolution(4, side, mode="bicubic")):
    y = op.H(x)
    flush.fill_(1)
    op.datafit_step(x, y, 0.5, out=z)
eps = torch.randn(S, B, 3, side, side, device="cuda", generator=g)
zt, v, out = torch.empty_like(eps), torch.randn(S, B, 3, side, side, device="cuda", generator=g), torch.empty_like(x)
flush.fill_(1)
_lib.check(lib.pnpf_interp(z.data_ptr(), eps.data_ptr(), 0.3, zt.data_ptr(), n, S, None))
flush.fill_(1)
_lib.check(lib.pnpf_push_accum(zt.data_ptr(), v.data_ptr(), 0.3, S, out.data_ptr(), n, None))
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
