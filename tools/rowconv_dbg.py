"""Cycle counters of CTA 0 of the row kernel.  Needs a library built with `make -C pnpflow_b200/csrc clean all EXTRA=-DPNPF_ROWCONV_CLOCKS`."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["PNPF_ROWCONV_DBG"] = "1"
from pnpflow_b200 import _lib
lib = _lib.load()
for (B, H, W, Cin, Cout) in ((80, 256, 256, 32, 32), (80, 256, 256, 64, 32), (80, 128, 128, 64, 64), (80, 256, 256, 32, 16)):
    x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
    w = torch.randn(Cout, Cin, 3, 3) * 0.05
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    print(f"--- B={B} {H}x{W} {Cin}->{Cout}", flush=True)
    for _ in range(2):
        _lib.check(lib.pnpf_conv2d_nhwc(x.data_ptr(), B, H, W, Cin, w.data_ptr(), None, Cout, 3, 1, None, 0, None, None, out.data_ptr(), 0, None))
print("=== fused GroupNorm variants", flush=True)
for (B, H, W, Ca, Cb, Cout) in ((80, 256, 256, 32, 0, 32), (80, 256, 256, 64, 32, 32), (80, 128, 128, 64, 0, 64)):
    xa = torch.randn(B, H, W, Ca, device="cuda").bfloat16()
    xb = torch.randn(B, H, W, Cb, device="cuda").bfloat16() if Cb else None
    C = Ca + Cb
    w = torch.randn(Cout, C, 3, 3) * 0.05
    gam, bet = torch.ones(C), torch.zeros(C)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    print(f"--- GN B={B} {H}x{W} {Ca}+{Cb}->{Cout}", flush=True)
    for _ in range(2):
        _lib.check(lib.pnpf_gn_conv2d_nhwc(xa.data_ptr(), Ca, xb.data_ptr() if Cb else None, Cb, B, H, W, gam.data_ptr(), bet.data_ptr(),
                                           w.data_ptr(), None, Cout, 1, out.data_ptr(), 0, None))
