import ctypes as C, os, json
here = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(here, "umma_rate_probe.so"))
res = []
iters = 4096
for rowb in (64, 128):
    for N in (16, 32, 64, 128, 256):
        for nrot in (1, 2, 3, 4, 8):
            if nrot * N > 512: continue
            for nblocks in (1, 148):
                buf = (C.c_longlong * nblocks)()
                rc = lib.run_rate(buf, nblocks, N, nrot, iters, rowb)
                cyc = sorted(buf)[len(buf) // 2]
                r = dict(rowb=rowb, N=N, nrot=nrot, nblocks=nblocks, rc=rc, cyc_per_mma=cyc / iters, floor=128 * N / 256)
                res.append(r); print(r, flush=True)
json.dump(res, open(os.path.join(here, "..", "..", "gpurun_out", "rate_probe.json"), "w"))
