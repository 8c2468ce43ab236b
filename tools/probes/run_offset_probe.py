import ctypes as C, os, json, torch
here = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(here, "umma_offset_probe.so"))
VP, I = C.c_void_p, C.c_int
lib.run_probe.argtypes = [VP, VP, VP, I, I, I]
res = []
for rowb in (128, 64):
    K = rowb // 2
    g = torch.Generator().manual_seed(0)
    A = torch.randn(256, K, generator=g).cuda().bfloat16()
    B = torch.randn(32, K, generator=g).cuda().bfloat16()
    for off in (0, 1, 2, 3, 4, 7, 8, 9, 13):
        for mode in ("zero", "addr"):
            base_off = 0 if mode == "zero" else ((off * rowb) >> 7) & 7
            out = torch.zeros(128, 32, device="cuda")
            rc = lib.run_probe(A.data_ptr(), B.data_ptr(), out.data_ptr(), rowb, off, base_off)
            ref = A[off:off + 128].float() @ B.float().t()
            err = (out - ref).abs().max().item()
            res.append(dict(rowb=rowb, off=off, mode=mode, base_off=base_off, rc=rc, err=err))
            print(res[-1], flush=True)
json.dump(res, open(os.path.join(here, "..", "..", "gpurun_out", "offset_probe.json"), "w"))
