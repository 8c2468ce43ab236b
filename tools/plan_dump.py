"""Print the U-Net launch plan (which kernel every conv of the net gets) WITHOUT a GPU: the size-query plan of
pnpf_workspace_bytes runs the same shape analysis as the real plan.
usage: [PNPF_NO_SUBPIX2=1 | PNPF_NO_SUBPIXEL=1 | ...] python tools/plan_dump.py [afhq256|celeba128] [batch]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["PNPF_PLAN_DUMP"] = "1"
from pnpflow_b200 import _lib, synth
net = synth.NETS[sys.argv[1] if len(sys.argv) > 1 else "afhq256"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 80
lib = _lib.load()
c = _lib.UNetConfigC()
c.input_channels, c.input_height, c.ch, c.num_levels = net["input_channels"], net["input_height"], net["ch"], len(net["ch_mult"])
for i, m in enumerate(net["ch_mult"]):
    c.ch_mult[i] = m
c.num_res_blocks, c.num_attn_resolutions = net["num_res_blocks"], len(net["attn_resolutions"])
for i, m in enumerate(net["attn_resolutions"]):
    c.attn_resolutions[i] = m
h = C.c_void_p()
_lib.check(lib.pnpf_create(C.byref(c), C.byref(h)))
n = lib.pnpf_workspace_bytes(h, B)
print(f"workspace for batch {B}: {n / 2**30:.2f} GiB", file=sys.stderr)
lib.pnpf_destroy(h)
