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
