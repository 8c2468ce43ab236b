"""GPU parity of the sub-pixel form of Upsample (models.py:41-47: nearest x2 + 3x3 conv as four 2x2-tap phases of the
patch-streaming kernel on the low-resolution tensor) — the default plan since round 2 (PNPF_NO_SUBPIXEL=1 is the A/B switch
back to upsample2x + 3x3 conv)."""
import pytest
import torch
import torch.nn.functional as F

import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 32, 256, 256), (2, 64, 128, 128), (2, 128, 64, 64), (3, 32, 128, 64)])
def test_upconv2x_layer_vs_torch(B, H, Cin, Cout):
    from pnpflow_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, H, H, Cin, generator=g).cuda().half()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05).half().float()       # fp16-exact 3x3 weights
    b = torch.randn(Cout, generator=g)
    out = torch.empty(B, 2 * H, 2 * H, Cout, device="cuda", dtype=torch.float16)
    _lib.check(lib.pnpf_upconv2x_nhwc(x.data_ptr(), B, H, H, Cin, w.contiguous().data_ptr(), b.data_ptr(), Cout, out.data_ptr(), None))
    ref = F.conv2d(F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest"), w.cuda(), b.cuda(), padding=1)
    got = out.float().permute(0, 3, 1, 2)
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 2e-3, rel          # fp16 output rounding + fp16 rounding of the folded (summed) weights


def test_unet_with_subpixel_up_matches_oracle():
    from pnpflow_b200 import UNetEngine
    cfg = oracle.AFHQ_256
    sd = oracle.init_state_dict(cfg, seed=0)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 3, 256, 256, generator=g).cuda()
    t = torch.tensor([0.3]).cuda()
    with torch.no_grad():
        ref = oracle.unet_forward({k: v.cuda() for k, v in sd.items()}, cfg, x, t)
    eng = UNetEngine(cfg, sd, max_batch=1)
    impls = [eng.lib.pnpf_debug_op_impl(eng._h, i).decode() for i in range(len(eng.op_names()))]
    v = eng(x, t)
    assert torch.isfinite(v).all()
    rel = ((v - ref).norm() / ref.norm()).item()
    assert rel < 4e-2, rel
    assert sum("subpix" in s for s in impls) == 8, "the sub-pixel plan (4 single phases at C_out = 256, 2 x two-phase launches at 128 and 64) was not selected"
