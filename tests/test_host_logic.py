"""CPU tests of the host side: C-ABI export list, plugin surface, host mask generation, schedules, sharding."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pnpflow_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pnpflow_b200.h")).read()
    declared = set(re.findall(r"\b(pnpf_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()                                   # raises if the .so is missing or lacks a symbol
    assert lib.pnpf_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (pnpf_[A-Za-z0-9_]+)", out))
    assert declared <= exported


def test_product_package_does_not_import_oracle():
    code = "import sys; import pnpflow_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_engine_plan_expects_the_reference_key_scheme():
    """pnpf_create needs no GPU: the weight registry must equal the reference's state_dict keys/shapes."""
    import ctypes as C
    from pnpflow_b200 import _lib
    lib = _lib.load()
    for cfg in (oracle.CELEBA_128, oracle.AFHQ_256, oracle.MNIST_28):
        c = _lib.UNetConfigC()
        c.input_channels, c.input_height, c.ch, c.num_levels = cfg.input_channels, cfg.input_height, cfg.ch, len(cfg.ch_mult)
        for i, m in enumerate(cfg.ch_mult):
            c.ch_mult[i] = m
        c.num_res_blocks, c.num_attn_resolutions = cfg.num_res_blocks, len(cfg.attn_resolutions)
        for i, m in enumerate(cfg.attn_resolutions):
            c.attn_resolutions[i] = m
        h = C.c_void_p()
        _lib.check(lib.pnpf_create(C.byref(c), C.byref(h)))
        names = [lib.pnpf_weight_name(h, i).decode() for i in range(lib.pnpf_num_weights(h))]
        sd = oracle.init_state_dict(cfg)
        assert set(names) == set(sd)
        # wrong shape / unknown key are errors with a message
        w = torch.zeros(3, 3)
        shape = (C.c_int64 * 2)(3, 3)
        assert lib.pnpf_load_weight(h, b"begin_conv.weight", w.data_ptr(), shape, 2) != 0
        assert b"begin_conv.weight" in lib.pnpf_last_error()
        assert lib.pnpf_load_weight(h, b"bogus", w.data_ptr(), shape, 2) != 0
        assert lib.pnpf_finalize_weights(h) != 0 and b"missing state_dict key" in lib.pnpf_last_error()
        if cfg is oracle.MNIST_28:        # 14x14 = 196 attention tokens: not a tensor-core tile multiple -> loud error
            assert lib.pnpf_workspace_bytes(h, 4) == 0 and b"attention over 196 tokens" in lib.pnpf_last_error()
        else:
            assert lib.pnpf_workspace_bytes(h, 4) > 0
        lib.pnpf_destroy(h)


def test_host_masks_equal_reference_restatement():
    import pnpflow_b200 as P
    x = torch.zeros(5, 3, 64, 64)
    assert (P.RandomInpainting(0.7)._host_mask(5, 64, 64) == oracle.RandomInpainting(0.7).mask(x)[:, 0].numpy()).all()
    assert (P.PaintbrushInpainting()._host_mask(5, 64, 64) == oracle.PaintbrushInpainting().mask(x)[:, 0].numpy()).all()
    # image i's mask does not depend on the batch size -> shards may slice the full-batch mask
    assert (P.RandomInpainting(0.7)._host_mask(5, 64, 64)[:2] == P.RandomInpainting(0.7)._host_mask(2, 64, 64)).all()
    with pytest.raises(Exception, match="at least 64"):
        P.PaintbrushInpainting()._host_mask(1, 32, 32)


def test_blur_taps_are_the_separable_factor_of_the_reference_kernel():
    import pnpflow_b200 as P
    for sigma in (1.0, 3.0):
        g = P.GaussianDeblurring(sigma, 61).taps_host
        k = oracle.operators.gaussian_kernel_2d(sigma, 61)
        assert (torch.outer(g, g) - k).abs().max() < 1e-8


def test_gamma_schedule_matches_reference_arithmetic():
    import pnpflow_b200 as P
    sigma = 0.05
    for style in ('alpha_1_minus_t', '1_minus_t', 'sqrt_1_minus_t', 'constant', 'unknown'):
        for it in (0, 1, 37, 99):
            t = torch.ones(1) * (1 / 100) * it
            lr_t = oracle.learning_rate(sigma ** 2 * 1.0, t, style, 0.3)
            ref = float(torch.as_tensor(lr_t).flatten()[0]) / sigma ** 2
            got = P.gamma_schedule(1.0, float(t[0]), style, 0.3)
            assert abs(got - ref) <= 2e-6 * max(1.0, abs(ref)), (style, it)


def test_step_time_is_bit_identical_to_the_reference_tensor_arithmetic():
    """pnp_flow.py:107-108: t1 = torch.ones(B) * delta * iteration (fp32 tensor x python scalars).  r01 computed
    fp32(double(delta * it)), which differs by 1 ulp on ~30 % of the steps."""
    import pnpflow_b200 as P
    for T in (20, 100, 200, 7):
        delta = 1 / T
        for it in range(T):
            ref = float((torch.ones(1) * delta * it)[0])
            assert P.step_time(delta, it) == ref, (T, it)


def test_as_engine_operator_accepts_foreign_plugins():
    import pnpflow_b200 as P
    from pnpflow_b200.degradations import _PythonOperator
    for o in (oracle.Denoising(), oracle.BoxInpainting(20), oracle.RandomInpainting(0.7), oracle.PaintbrushInpainting(),
              oracle.GaussianDeblurring(1.0, 61, "fft", 3, 128, "cpu"), oracle.Superresolution(2, 128)):
        e = P.as_engine_operator(o)
        assert type(e).__name__ == type(o).__name__ and isinstance(e, P.Degradation)

    class Custom:
        def H(self, x): return 2 * x
        def H_adj(self, x): return 2 * x
    w = P.as_engine_operator(Custom())
    assert isinstance(w, _PythonOperator)
    x, y = torch.ones(1, 1, 2, 2), torch.zeros(1, 1, 2, 2)
    assert torch.allclose(w.datafit_step(x, y, 0.1), x - 0.1 * 4 * x)
    with pytest.raises(TypeError):
        P.as_engine_operator(object())


def test_subpixel_weight_fold_equals_nearest_upsample_plus_conv3x3():
    """models.py:41-47 (Upsample = F.interpolate(scale 2, nearest) -> 3x3 conv) restated as four 2x2 convolutions on the
    low-resolution tensor with the weights folded by the C ABI's host-side pnpf_fold_subpixel_weights (the opt-in
    sub-pixel plan of the up convs): output pixel (2h+a, 2w+b) = sum_ij W_ab[i,j] * x[h-1+a+i, w-1+b+j]."""
    import torch.nn.functional as F
    from pnpflow_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    O, I, H, W = 6, 5, 7, 9
    w = torch.randn(O, I, 3, 3, generator=g, dtype=torch.float32)
    x = torch.randn(2, I, H, W, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w.double(), padding=1)
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for a in (0, 1):
        for b in (0, 1):
            f = torch.empty(O, I, 2, 2, dtype=torch.float32)
            _lib.check(lib.pnpf_fold_subpixel_weights(w.contiguous().data_ptr(), O, I, a, b, f.data_ptr()))
            y = F.conv2d(xp, f.double())                      # y[h', w'] = sum_ij f[i,j] xp[h'+i, w'+j], xp index = x index + 1
            out[:, :, a::2, b::2] = y[:, :, a:a + H, b:b + W]
    assert (out - ref).abs().max() < 1e-5


def test_cpu_tensors_are_rejected_not_silently_computed():
    import pnpflow_b200 as P
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.BoxInpainting(20).H(torch.ones(1, 3, 128, 128))


def test_shard_bounds_cover_batch():
    from pnpflow_b200.sharding import shard_bounds
    for B in (1, 7, 64, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _dist_worker(rank, world, port, B, q):
    import torch.distributed as dist
    from pnpflow_b200 import sharding
    import pnpflow_b200 as P
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = torch.randn(B, 3, 8, 8, generator=g)
        lo, hi = sharding.shard_bounds(B, world, rank)
        mine = sharding.scatter_batch(full if rank == 0 else None, full.shape, full.dtype, "cpu")
        ok = torch.equal(mine, full[lo:hi])
        back = sharding.gather_batch(mine * 2, B)
        ok &= torch.equal(back, full * 2)
        # full-batch noise: every rank draws the same stream and keeps its slice
        gen = torch.Generator().manual_seed(123)
        it = sharding.FullBatchNoise((B, 3, 8, 8), lo, hi, "cpu", gen)
        n0, n1 = next(it), next(it)
        gen2 = torch.Generator().manual_seed(123)
        r0, r1 = torch.randn(B, 3, 8, 8, generator=gen2), torch.randn(B, 3, 8, 8, generator=gen2)
        ok &= torch.equal(n0, r0[lo:hi]) and torch.equal(n1, r1[lo:hi])
        # sharded random mask == slice of the full-batch mask
        op = sharding.shard_operator(P.RandomInpainting(0.7), lo, hi, B)
        ok &= bool((op._host_mask(hi - lo, 64, 64) == P.RandomInpainting(0.7)._host_mask(B, 64, 64)[lo:hi]).all())
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 5])
def test_scatter_gather_noise_and_masks_world2_gloo(B):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + B) % 2000
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_subpixel_pair_weight_packing_equals_nearest_upsample_plus_conv3x3():
    """The operand of the two-column-phase launch (SUBPIX = 2 of pnpf_patchconv.cuh: both column phases b of an output-row parity a as
    a 2 x 3-tap convolution with N = 2*Cout): evaluated as a plain sum on the CPU it must reproduce models.py:41-47."""
    import torch.nn.functional as F
    from pnpflow_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(6)
    O, I, H, W = 4, 3, 5, 6
    w = (torch.randn(O, I, 3, 3, generator=g) * 0.25).half().float()            # fp16-exact weights; the folded sums round once more
    x = torch.randn(2, I, H, W, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w.double(), padding=1)
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for a in (0, 1):
        packed = torch.empty(2 * O, 6 * I, dtype=torch.float32)
        _lib.check(lib.pnpf_pack_subpixel_pair_weights(w.contiguous().data_ptr(), O, I, a, packed.data_ptr()))
        wp = packed.double().view(2, O, 2, 3, I)                                  # [b][o][i][c][ch]
        for b in (0, 1):
            acc = torch.zeros(2, O, H, W, dtype=torch.float64)
            for i in (0, 1):
                for c in (0, 1, 2):
                    patch = xp[:, :, a + i:a + i + H, c:c + W]                    # x[h-1+a+i, w-1+c]
                    acc += torch.einsum("oc,bchw->bohw", wp[b, :, i, c, :], patch)
            out[:, :, a::2, b::2] = acc
    assert (out - ref).abs().max() < 2e-3          # fp16 rounding of the folded (summed) weights
