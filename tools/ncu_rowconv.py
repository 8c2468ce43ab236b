import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pnpflow_b200 import _lib
lib = _lib.load()
B, H, W, Cin, Cout = 80, 256, 256, int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 32
x = torch.randn(B, H, W, Cin, device="cuda").half()
w = torch.randn(Cout, Cin, 3, 3) * 0.05
out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.float16)
for _ in range(3):
    _lib.check(lib.pnpf_conv2d_nhwc(x.data_ptr(), B, H, W, Cin, w.data_ptr(), None, Cout, 3, 1, None, 0, None, None, out.data_ptr(), 0, None))
torch.cuda.synchronize()
