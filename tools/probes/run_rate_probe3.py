import ctypes as C, os, json
here = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(here, "umma_rate_probe3.so"))
res = []
for N in (16, 32, 48, 64, 96, 128, 192, 256):
    for a_shift in (0, 1):
        for nacc in (1, 2):
            if nacc * N > 512: continue
            buf = (C.c_longlong * 148)()
            rc = lib.run_rate3(buf, 148, N, 4096, a_shift, nacc)
            r = dict(N=N, a_shift=a_shift, nacc=nacc, rc=rc, cyc_per_mma=round(sorted(buf)[74] / 4096, 1), floor=N / 2)
            res.append(r); print(r, flush=True)
json.dump(res, open(os.path.join(here, "..", "..", "gpurun_out", "rate_probe3.json"), "w"))
