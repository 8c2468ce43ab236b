"""Cycle counters of CTA 0 of both tensor-core kernels (producer / MMA issuer / transform / epilogue: total clocks and clocks
spent waiting on each barrier).  Needs a library built with `make -C pnpflow_b200/csrc clean all EXTRA=-DPNPF_ROWCONV_CLOCKS`
(set PNPF_LIB to its path)."""
import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["PNPF_ROWCONV_DBG"] = "1"
from pnpflow_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 80
ONLY_GN = len(sys.argv) > 2 and sys.argv[2] == "gn"


def timed(fn):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"    {e0.elapsed_time(e1):.3f} ms (incl. host-side weight packing of the layer API)", flush=True)


# DBG_SHAPES="H,W,Cin,Cout,C2,res;..." replaces the built-in list of plain convs (and skips the fused-GroupNorm variants)
CUSTOM = [tuple(int(v) for v in sh.split(",")) for sh in os.environ.get("DBG_SHAPES", "").split(";") if sh]
print("=== plain convs: B H W Cin Cout C2 res", flush=True)
for (H, W, Cin, Cout, C2, res) in () if ONLY_GN else CUSTOM if CUSTOM else ((256, 256, 32, 32, 0, 0), (256, 256, 32, 32, 0, 1), (256, 256, 32, 32, 32, 0), (256, 256, 64, 32, 0, 0),
                                   (128, 128, 64, 64, 0, 0), (128, 128, 64, 64, 0, 1), (128, 128, 64, 64, 64, 0), (128, 128, 128, 64, 0, 0),
                                   (128, 128, 64, 64, 128, 0),
                                   (64, 64, 128, 128, 0, 0), (64, 64, 128, 128, 0, 1), (64, 64, 256, 128, 0, 0), (64, 64, 128, 128, 256, 0),
                                   (32, 32, 256, 256, 0, 0), (32, 32, 512, 256, 0, 0), (32, 32, 256, 256, 512, 0)):
    x = torch.randn(B, H, W, Cin, device="cuda").half()
    w = torch.randn(Cout, Cin, 3, 3) * 0.05
    x2 = torch.randn(B, H, W, C2, device="cuda").half() if C2 else None
    w2 = torch.randn(Cout, C2, 1, 1) * 0.05 if C2 else None
    r = torch.randn(B, H, W, Cout, device="cuda").half() if res else None
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.float16)
    print(f"--- B={B} {H}x{W} {Cin}->{Cout} C2={C2} res={res}", flush=True)
    timed(lambda: _lib.check(lib.pnpf_conv2d_nhwc(x.data_ptr(), B, H, W, Cin, w.data_ptr(), None, Cout, 3, 1, x2.data_ptr() if C2 else None, C2,
                                                  w2.data_ptr() if C2 else None, r.data_ptr() if res else None, out.data_ptr(), 0, None)))
print("=== fused GroupNorm variants", flush=True)
for (H, W, Ca, Cb, Cout) in () if CUSTOM else ((256, 256, 32, 0, 32), (256, 256, 32, 32, 32), (256, 256, 64, 32, 32), (128, 128, 64, 0, 64), (128, 128, 64, 64, 64),
                             (128, 128, 64, 32, 64)):
    xa = torch.randn(B, H, W, Ca, device="cuda").half()
    xb = torch.randn(B, H, W, Cb, device="cuda").half() if Cb else None
    Cc = Ca + Cb
    w = torch.randn(Cout, Cc, 3, 3) * 0.05
    gam, bet = torch.ones(Cc), torch.zeros(Cc)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.float16)
    print(f"--- GN B={B} {H}x{W} {Ca}+{Cb}->{Cout}", flush=True)
    timed(lambda: _lib.check(lib.pnpf_gn_conv2d_nhwc(xa.data_ptr(), Ca, xb.data_ptr() if Cb else None, Cb, B, H, W, gam.data_ptr(), bet.data_ptr(),
                                                     w.data_ptr(), None, Cout, 1, out.data_ptr(), 0, None)))
