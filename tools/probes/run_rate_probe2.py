import ctypes as C, os, json
here = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(here, "umma_rate_probe2.so"))
res = []
iters = 3600
def run(N, a_off, nbt, spin, lane0only, commit_every, nblocks=148):
    buf = (C.c_longlong * nblocks)()
    rc = lib.run_rate2(buf, nblocks, N, iters, a_off, nbt, spin, lane0only, commit_every)
    cyc = sorted(buf)[len(buf) // 2]
    r = dict(N=N, a_off=a_off, nbt=nbt, spin=spin, lane0only=lane0only, commit_every=commit_every, rc=rc, cyc_per_mma=round(cyc / iters, 1))
    res.append(r); print(r, flush=True)
run(32, 0, 1, 0, 0, 0)
for a_off in (1, 2): run(32, a_off, 1, 0, 0, 0)
run(32, 0, 9, 0, 0, 0)
run(32, 1, 9, 0, 0, 0)
run(32, 0, 1, 0, 0, 6)
run(32, 0, 1, 0, 0, 18)
for spin in (1, 4, 9): run(32, 0, 1, spin, 0, 0)
for spin in (1, 4, 9): run(32, 0, 1, spin, 1, 0)
run(32, 1, 9, 9, 0, 18)
run(32, 1, 9, 9, 1, 18)
run(96, 1, 3, 9, 1, 18)
run(96, 0, 1, 0, 0, 0)
run(192, 0, 1, 0, 0, 0)
json.dump(res, open(os.path.join(here, "..", "..", "gpurun_out", "rate_probe2.json"), "w"))
