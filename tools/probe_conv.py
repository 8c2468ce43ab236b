"""GPU probe of the tcgen05 implicit-GEMM kernel: each case runs in its own subprocess with a timeout so a
deadlocked variant cannot hang the whole call.  Results -> gpurun_out/probe_conv.jsonl"""
import ctypes as C, json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: kind, params
    "gemm_1tile":   ("gemm", dict(batch=1, M=128, N=64, K=64)),
    "gemm_k256":    ("gemm", dict(batch=1, M=128, N=64, K=256)),
    "gemm_n256":    ("gemm", dict(batch=2, M=256, N=256, K=256)),
    "gemm_n512":    ("gemm", dict(batch=2, M=256, N=512, K=128)),
    "gemm_big":     ("gemm", dict(batch=4, M=1024, N=1024, K=256)),
    "c1x1_64":      ("conv", dict(B=1, H=16, W=16, Cin=64, Cout=64, k=1, s=1)),
    "c3x3_64_16":   ("conv", dict(B=2, H=16, W=16, Cin=64, Cout=64, k=3, s=1)),
    "c3x3_32_128":  ("conv", dict(B=1, H=128, W=128, Cin=32, Cout=32, k=3, s=1)),
    "c3x3_96_64":   ("conv", dict(B=2, H=64, W=64, Cin=96, Cout=64, k=3, s=1)),
    "c3x3_256_32":  ("conv", dict(B=2, H=32, W=32, Cin=256, Cout=256, k=3, s=1)),
    "c3x3_512_16":  ("conv", dict(B=2, H=16, W=16, Cin=512, Cout=256, k=3, s=1)),
    "c3x3_s2":      ("conv", dict(B=2, H=32, W=32, Cin=64, Cout=64, k=3, s=2)),
    "c3x3_s2_256":  ("conv", dict(B=1, H=256, W=256, Cin=32, Cout=32, k=3, s=2)),
    "c3x3_cout16":  ("conv", dict(B=1, H=64, W=64, Cin=32, Cout=16, k=3, s=1)),
    "c3x3_w28":     ("conv", dict(B=2, H=28, W=28, Cin=32, Cout=32, k=3, s=1)),
    "c3x3_w14":     ("conv", dict(B=2, H=14, W=14, Cin=64, Cout=64, k=3, s=1)),
    "c3x3_fused":   ("conv", dict(B=2, H=32, W=32, Cin=128, Cout=128, k=3, s=1, C2=192, res=True, bf16out=True)),
    "c3x3_fused32": ("conv", dict(B=2, H=64, W=64, Cin=32, Cout=32, k=3, s=1, C2=96, res=False)),
    "perf_32_256":  ("conv", dict(B=16, H=256, W=256, Cin=32, Cout=32, k=3, s=1, perf=True)),
    "perf_64_128":  ("conv", dict(B=16, H=128, W=128, Cin=64, Cout=64, k=3, s=1, perf=True)),
    "perf_128_64":  ("conv", dict(B=16, H=64, W=64, Cin=128, Cout=128, k=3, s=1, perf=True)),
    "perf_256_32":  ("conv", dict(B=16, H=32, W=32, Cin=256, Cout=256, k=3, s=1, perf=True)),
    "perf_512_32":  ("conv", dict(B=16, H=32, W=32, Cin=512, Cout=256, k=3, s=1, perf=True)),
}


def run_case(name):
    import torch
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = C.CDLL(os.path.join(ROOT, "pnpflow_b200", "libpnpflow_sm100a.so"))
    lib.pnpf_last_error.restype = C.c_char_p
    VP, I = C.c_void_p, C.c_int
    lib.pnpf_conv2d_nhwc.argtypes = [VP, I, I, I, I, VP, VP, I, I, I, VP, I, VP, VP, VP, I, VP]
    lib.pnpf_gemm_nt.argtypes = [VP, VP, VP, I, I, I, I, I, VP]
    kind, p = CASES[name]
    g = torch.Generator(device="cpu").manual_seed(hash(name) % 1000)
    dev = "cuda"
    res = dict(case=name, **{k: v for k, v in p.items()})
    if kind == "gemm":
        A = torch.randn(p["batch"], p["M"], p["K"], generator=g).to(dev).half()
        Bm = torch.randn(p["batch"], p["N"], p["K"], generator=g).to(dev).half()
        out = torch.full((p["batch"], p["M"], p["N"]), float("nan"), device=dev, dtype=torch.float32)
        rc = lib.pnpf_gemm_nt(A.data_ptr(), Bm.data_ptr(), out.data_ptr(), p["batch"], p["M"], p["N"], p["K"], 1, None)
        if rc:
            res["error"] = lib.pnpf_last_error().decode()
            return res
        ref = torch.bmm(A.float(), Bm.float().transpose(1, 2))
        err = (out - ref).abs()
        res.update(max_err=float(err.max()), ref_rms=float(ref.pow(2).mean().sqrt()), nan=int(torch.isnan(out).sum()))
        if not (err.max() < 1e-2):
            bad = (err > 1e-2) | torch.isnan(out)
            res["bad_rows"] = bad.any(dim=2)[0].nonzero().flatten()[:16].tolist()
            res["bad_cols"] = bad.any(dim=1)[0].nonzero().flatten()[:16].tolist()
            res["bad_frac"] = float(bad.float().mean())
        return res
    B, H, W, Cin, Cout, k, s = p["B"], p["H"], p["W"], p["Cin"], p["Cout"], p["k"], p["s"]
    C2 = p.get("C2", 0)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev).half()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).half().float()
    b = torch.randn(Cout, generator=g)
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    x2 = w2 = None
    if C2:
        x2 = torch.randn(B, C2, Ho, Wo, generator=g).to(dev).half()
        w2 = (torch.randn(Cout, C2, 1, 1, generator=g) / C2 ** 0.5).half().float()
        x2_nhwc = x2.permute(0, 2, 3, 1).contiguous()
    resid = None
    if p.get("res"):
        resid = torch.randn(B, Cout, Ho, Wo, generator=g).to(dev).half()
        resid_nhwc = resid.permute(0, 2, 3, 1).contiguous()
    f32 = 0 if p.get("bf16out") else 1
    out = torch.full((B, Ho, Wo, Cout), float("nan"), device=dev, dtype=torch.float32 if f32 else torch.float16)
    wc, bc = w.contiguous(), b.contiguous()
    w2c = w2.contiguous() if C2 else None
    args = (x_nhwc.data_ptr(), B, H, W, Cin, wc.data_ptr(), bc.data_ptr(), Cout, k, s,
            x2_nhwc.data_ptr() if C2 else None, C2, w2c.data_ptr() if C2 else None,
            resid_nhwc.data_ptr() if resid is not None else None, out.data_ptr(), f32, None)
    rc = lib.pnpf_conv2d_nhwc(*args)
    if rc:
        res["error"] = lib.pnpf_last_error().decode()
        return res
    ref = F.conv2d(x.float(), w.to(dev), b.to(dev), stride=s, padding=k // 2)
    if C2:
        ref = ref + F.conv2d(x2.float(), w2.to(dev))
    if resid is not None:
        ref = ref + resid.float()
    got = out.float().permute(0, 3, 1, 2)
    err = (got - ref).abs()
    tol = 1e-2 if f32 else 5e-2
    res.update(max_err=float(err.max()), ref_rms=float(ref.pow(2).mean().sqrt()), nan=int(torch.isnan(got).sum()))
    if not (err.max() < tol):
        bad = (err > tol) | torch.isnan(got)
        res["bad_frac"] = float(bad.float().mean())
        res["bad_ch"] = bad.any(dim=3).any(dim=2).any(dim=0).nonzero().flatten()[:16].tolist()
        res["bad_h"] = bad.any(dim=3).any(dim=1).any(dim=0).nonzero().flatten()[:16].tolist()
        res["bad_w"] = bad.any(dim=2).any(dim=1).any(dim=0).nonzero().flatten()[:16].tolist()
    if p.get("perf"):
        for _ in range(3):
            lib.pnpf_conv2d_nhwc(*args)
        # the layer-level call includes weight packing + malloc; time the kernel through CUDA events around calls
        # is meaningless -> use the profiler-free approach: many launches inside one timing is not exposed at this
        # level, so report the end-to-end call time only as an upper bound.
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            lib.pnpf_conv2d_nhwc(*args)
        torch.cuda.synchronize()
        res["call_ms_upper_bound"] = (time.perf_counter() - t0) / 5 * 1e3
        res["gflop"] = 2.0 * B * Ho * Wo * Cout * Cin * k * k / 1e9
    return res


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        print("RESULT " + json.dumps(run_case(sys.argv[2])), flush=True)
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    outp = os.path.join(ROOT, "gpurun_out", "probe_conv.jsonl")
    names = sys.argv[1:] or list(CASES)
    with open(outp, "a") as f:
        for n in names:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, __file__, "--case", n], capture_output=True, text=True, timeout=150)
                line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                rec = json.loads(line[-1][7:]) if line else dict(case=n, crashed=True, rc=r.returncode, stderr=r.stderr[-1500:])
            except subprocess.TimeoutExpired:
                rec = dict(case=n, timeout=True)
            rec["wall_s"] = round(time.time() - t0, 1)
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec), flush=True)
