"""GPU parity of the tcgen05 implicit-GEMM kernel (through the C ABI) against torch fp32 conv2d / bmm on the same
fp16-rounded operands.  Inputs and weights are exactly representable in fp16 (the engine's storage / MMA operand type), so the
only differences are the fp32 accumulation order (and fp16 rounding of the output when out is fp16; `bf16out` below is the
historical name of that switch)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _lib():
    from pnpflow_b200 import _lib
    return _lib.load(), _lib


def _conv_case(B, H, W, Cin, Cout, k, s, C2=0, res=False, bf16out=False, seed=0):
    lib, L = _lib()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(seed)
    dev = "cuda"
    x = torch.randn(B, Cin, H, W, generator=g).to(dev).half()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).half().float().contiguous()
    b = torch.randn(Cout, generator=g).contiguous()
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    xn = x.permute(0, 2, 3, 1).contiguous()
    x2n = w2 = rn = None
    ref = F.conv2d(x.float(), w.to(dev), b.to(dev), stride=s, padding=k // 2)
    if C2:
        x2 = torch.randn(B, C2, Ho, Wo, generator=g).to(dev).half()
        w2 = (torch.randn(Cout, C2, 1, 1, generator=g) / C2 ** 0.5).half().float().contiguous()
        x2n = x2.permute(0, 2, 3, 1).contiguous()
        ref = ref + F.conv2d(x2.float(), w2.to(dev))
    if res:
        r = torch.randn(B, Cout, Ho, Wo, generator=g).to(dev).half()
        rn = r.permute(0, 2, 3, 1).contiguous()
        ref = ref + r.float()
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.float16 if bf16out else torch.float32)
    L.check(lib.pnpf_conv2d_nhwc(xn.data_ptr(), B, H, W, Cin, w.data_ptr(), b.data_ptr(), Cout, k, s,
                                 x2n.data_ptr() if C2 else None, C2, w2.data_ptr() if C2 else None,
                                 rn.data_ptr() if res else None, out.data_ptr(), 0 if bf16out else 1, None))
    got = out.float().permute(0, 3, 1, 2)
    tol = max(4e-3, ref.abs().max().item() * 2.0 ** -11) if bf16out else 2e-3     # fp16 output: half an ulp of the largest value
    err = (got - ref).abs().max().item()
    assert err < tol, f"max abs err {err}"


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,s", [
    (1, 16, 16, 64, 64, 1, 1), (2, 16, 16, 64, 64, 3, 1), (1, 128, 128, 32, 32, 3, 1), (2, 64, 64, 96, 64, 3, 1),
    (2, 32, 32, 256, 256, 3, 1), (2, 16, 16, 512, 256, 3, 1), (1, 64, 64, 32, 16, 3, 1), (1, 256, 256, 32, 32, 3, 1),
    (2, 16, 16, 256, 512, 1, 1),
])
def test_conv_stride1(B, H, W, Cin, Cout, k, s):
    _conv_case(B, H, W, Cin, Cout, k, s)


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 32, 32, 64, 64), (1, 256, 256, 32, 32), (2, 64, 64, 128, 128)])
def test_conv_stride2(B, H, W, Cin, Cout):
    _conv_case(B, H, W, Cin, Cout, 3, 2)


@pytest.mark.parametrize("B,H,W,Cin,Cout,C2,res", [
    (2, 128, 128, 32, 32, 0, False), (1, 256, 256, 32, 32, 0, True), (2, 40, 128, 64, 32, 0, False), (1, 128, 256, 96, 32, 0, False),
    (2, 128, 128, 64, 64, 0, True), (1, 64, 128, 32, 64, 32, False), (2, 128, 256, 32, 32, 96, False), (3, 33, 128, 32, 32, 64, False),
    (1, 128, 128, 32, 16, 0, False),
    # output channels split over CTA pairs (weights of the full C_out do not fit next to the row ring)
    (2, 128, 128, 128, 64, 0, False), (3, 50, 128, 128, 64, 0, True), (1, 64, 128, 64, 64, 128, False), (2, 37, 256, 64, 64, 96, False),
    # many short ranges: more CTAs than 8-row pieces, pieces crossing image boundaries
    (5, 9, 128, 32, 32, 0, True), (7, 3, 128, 64, 64, 64, False),
])
@pytest.mark.parametrize("bf16out", [False, True])
def test_rowconv_wide_maps(B, H, W, Cin, Cout, C2, res, bf16out):
    """W % 128 == 0 and Cout <= 64 route to the row-streaming kernel (halo tile + row-shifted UMMA descriptors); bf16 NHWC
    outputs leave through the staging tiles + TMA store, fp32 outputs through per-thread stores."""
    _conv_case(B, H, W, Cin, Cout, 3, 1, C2=C2, res=res, bf16out=bf16out, seed=H + W + Cin)


@pytest.mark.parametrize("B,H,W,Ca,Cb,Cout,silu", [
    (2, 64, 128, 32, 0, 32, 1), (1, 256, 256, 32, 0, 32, 1), (2, 40, 128, 64, 0, 32, 1), (2, 64, 128, 64, 32, 32, 1),
    (1, 128, 256, 32, 32, 32, 1), (2, 64, 128, 64, 0, 64, 1), (2, 33, 128, 32, 0, 16, 1), (1, 64, 128, 32, 0, 64, 0), (2, 64, 256, 64, 32, 32, 1),
    (2, 64, 128, 64, 64, 64, 1), (3, 21, 128, 64, 32, 64, 1), (1, 128, 128, 128, 0, 64, 1),      # split output channels / 3 K chunks
])
@pytest.mark.parametrize("bf16out", [False, True])
def test_rowconv_fused_groupnorm_two_sources(B, H, W, Ca, Cb, Cout, silu, bf16out):
    """conv3x3(act(GroupNorm(cat[xa|xb]))) with the normalisation done in shared memory inside the conv kernel."""
    lib, L = _lib()
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(B + H + Ca + Cb)
    dev = "cuda"
    C = Ca + Cb
    xa = (torch.randn(B, Ca, H, W, generator=g) * 1.5 + 0.3).to(dev).half()
    xb = (torch.randn(B, Cb, H, W, generator=g) * 0.7 - 0.2).to(dev).half() if Cb else None
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).contiguous()
    beta = (0.1 * torch.randn(C, generator=g)).contiguous()
    w = (torch.randn(Cout, C, 3, 3, generator=g) / (C * 9) ** 0.5).half().float().contiguous()
    b = torch.randn(Cout, generator=g).contiguous()
    x = torch.cat([xa, xb], 1).float() if Cb else xa.float()
    a = F.group_norm(x, 32, gamma.to(dev), beta.to(dev), eps=1e-6)
    if silu:
        a = torch.sigmoid(a) * a
    ref = F.conv2d(a.half().float(), w.to(dev), b.to(dev), padding=1)
    xan = xa.permute(0, 2, 3, 1).contiguous()
    xbn = xb.permute(0, 2, 3, 1).contiguous() if Cb else None
    out = torch.full((B, H, W, Cout), float("nan"), device=dev, dtype=torch.float16 if bf16out else torch.float32)
    L.check(lib.pnpf_gn_conv2d_nhwc(xan.data_ptr(), Ca, xbn.data_ptr() if Cb else None, Cb, B, H, W, gamma.data_ptr(), beta.data_ptr(),
                                    w.data_ptr(), b.data_ptr(), Cout, silu, out.data_ptr(), 0 if bf16out else 1, None))
    got = out.float().permute(0, 3, 1, 2)
    # the engine rounds the normalised activation to fp16 like the reference above; tanh.approx / rounding-boundary
    # flips give rare 1-ulp(fp16) operand differences -> compare in relative L2 and with a loose max
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 3e-3, rel          # (tanh.approx.f32: 2^-11 relative; fp16 output rounding adds ~2^-12)
    assert (got - ref).abs().max().item() < 5e-2


@pytest.mark.parametrize("B,H,W,Cin,Cout,C2,res,bf16out", [
    (2, 64, 64, 128, 128, 0, False, False), (3, 64, 64, 256, 128, 0, True, True), (2, 128, 128, 128, 128, 0, False, True),
    (1, 8, 8, 256, 256, 0, False, False), (4, 32, 32, 512, 256, 0, False, True), (2, 64, 64, 128, 128, 256, False, True),
    (5, 16, 16, 256, 256, 384, False, False), (2, 24, 40, 64, 128, 0, True, False), (6, 32, 32, 256, 256, 0, True, True),
    (2, 128, 128, 192, 64, 0, False, True), (3, 64, 64, 64, 64, 192, False, False), (2, 32, 32, 64, 64, 0, True, True),
    # channel counts that are odd multiples of 32: the trailing half chunk is zero-filled by the TMA unit (128^2 net, level 1)
    (2, 64, 64, 32, 64, 0, False, True), (3, 64, 64, 96, 64, 0, True, False), (2, 64, 64, 64, 64, 96, False, True), (1, 32, 32, 160, 128, 32, False, False),
])
def test_patchconv_narrow_maps(B, H, W, Cin, Cout, C2, res, bf16out):
    """3x3 stride-1 convs with W <= 128 and C_out 128/256 route to the patch-streaming kernel (padded-linear tiles, one
    halo patch per 64-channel chunk, nine row-shifted UMMA descriptors); even batches run as CTA pairs (cta_group::2)."""
    _conv_case(B, H, W, Cin, Cout, 3, 1, C2=C2, res=res, bf16out=bf16out, seed=B + H + Cin)


def test_conv_ragged_edges():
    _conv_case(2, 28, 28, 32, 32, 3, 1)
    _conv_case(2, 14, 14, 64, 64, 3, 1)


def test_conv_fused_shortcut_residual_bf16():
    _conv_case(2, 32, 32, 128, 128, 3, 1, C2=192, res=True, bf16out=True)
    _conv_case(2, 64, 64, 32, 32, 3, 1, C2=96)
    _conv_case(2, 32, 32, 64, 64, 3, 1, res=True)


@pytest.mark.parametrize("batch,M,N,K", [(1, 128, 64, 64), (2, 256, 256, 256), (2, 256, 512, 128), (3, 1024, 1024, 256), (2, 64, 256, 64)])
def test_gemm_nt(batch, M, N, K):
    lib, L = _lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(1)
    A = torch.randn(batch, M, K, generator=g).cuda().half()
    Bm = torch.randn(batch, N, K, generator=g).cuda().half()
    out = torch.full((batch, M, N), float("nan"), device="cuda")
    L.check(lib.pnpf_gemm_nt(A.data_ptr(), Bm.data_ptr(), out.data_ptr(), batch, M, N, K, 1, None))
    ref = torch.bmm(A.float(), Bm.float().transpose(1, 2))
    assert (out - ref).abs().max().item() < 2e-3 * K ** 0.5


def test_unsupported_shape_is_an_error_not_a_fallback():
    lib, L = _lib()
    x = torch.zeros(1, 8, 8, 24, device="cuda", dtype=torch.float16)
    w = torch.zeros(16, 24, 3, 3)
    out = torch.zeros(1, 8, 8, 16, device="cuda")
    rc = lib.pnpf_conv2d_nhwc(x.data_ptr(), 1, 8, 8, 24, w.data_ptr(), None, 16, 3, 1, None, 0, None, None, out.data_ptr(), 1, None)
    assert rc != 0 and b"multiples of 32" in lib.pnpf_last_error()


@pytest.mark.parametrize("B", [1, 5, 150])
def test_fused_attention_core_vs_torch(B):
    """pnpf_attn.cuh: out = x + softmax(q k^T) v Wo^T + b for the 16x16 attention blocks (models.py:145-162), logits in TMEM,
    softmax in registers, probabilities and O through shared memory.  B = 150 gives every CTA more than one (image, query tile)
    unit, which exercises the barrier phases and the shared-memory / TMEM hand-over between units."""
    from pnpflow_b200 import _lib
    lib = _lib.load()
    L = C = 256
    g = torch.Generator().manual_seed(11 + B)
    q = (torch.randn(B, L, C, generator=g) * 0.125).half()           # logits ~ N(0, 4): a peaked but not one-hot softmax
    k = torch.randn(B, L, C, generator=g).half()
    v = torch.randn(B, L, C, generator=g).half()
    wo = (torch.randn(C, C, generator=g) * 0.05).half().float()
    bias = torch.randn(C, generator=g)
    res = torch.randn(B, L, C, generator=g).half()
    qk = torch.cat([q, k], dim=-1).contiguous().cuda()
    vT = v.transpose(1, 2).contiguous().cuda()
    out = torch.empty(B, L, C, device="cuda", dtype=torch.float16)
    _lib.check(lib.pnpf_attn_core_nhwc(qk.data_ptr(), vT.data_ptr(), wo.contiguous().data_ptr(), bias.data_ptr(), res.cuda().data_ptr(),
                                       out.data_ptr(), B, L, C, None))
    qf, kf, vf = q.float().cuda(), k.float().cuda(), v.float().cuda()
    P = torch.softmax(qf @ kf.transpose(1, 2), dim=-1)
    ref = res.float().cuda() + (P @ vf) @ wo.cuda().t() + bias.cuda()
    got = out.float()
    assert torch.isfinite(got).all()
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 2e-3, rel          # fp16 P, fp16 O and the fp16 output rounding
    attn_only = ((got - res.float().cuda() - bias.cuda()) - (ref - res.float().cuda() - bias.cuda())).norm() / (ref - res.float().cuda() - bias.cuda()).norm()
    assert attn_only.item() < 5e-3, attn_only.item()
