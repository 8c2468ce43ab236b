#!/usr/bin/env python
"""bench.py — restored images/sec of the PnP-Flow hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--config cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE outer PnP-Flow iteration (pnp_flow.py:107-121: data-fidelity step + S x (noise draw, interpolate,
U-Net velocity, Euler push) + MC average) over the per-GPU batch of synthetic measurements.  Restoring an image
costs T such steps, so   images/sec = (N * B_per_gpu) / (T * seconds_per_step).
Workload at every N: cfg4 of BASELINE.json (AFHQ-Cat 256x256x3, 4x super-resolution, T=100, S=5 MC draws, the
reference default of config/method_config/pnp_flow.yaml), 16 images per GPU (128 over 8) -> weak scaling.

engine arm:      pnpflow_b200 (hand-written sm_100a CUDA behind the C ABI), fp16 tensor-core operands (11-bit significand) / fp32 accumulate.
torch_gpu_baseline: the reference's own GPU path like for like (oracle port, eager PyTorch + cuDNN TF32 on the same B200).
reference arm:   the reference algorithm on the host CPU cores (oracle port; the reference itself is pure PyTorch and
                 its tree does not exist on the GPU box), rank 0 only, a bounded sample extrapolated linearly.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIGS = {
    # name: net, problem, images per GPU, T, S, sigma, alpha   (BASELINE.json configs 2-5; main.py:120-179 constants)
    "cfg2": dict(net="celeba128", problem="inpainting", b_per_gpu=64, T=100, S=5, sigma=0.05, alpha=0.5),
    "cfg3": dict(net="celeba128", problem="gaussian_deblurring_FFT", b_per_gpu=32, T=100, S=5, sigma=0.05, alpha=0.01),
    "cfg4": dict(net="afhq256", problem="superresolution", b_per_gpu=16, T=100, S=5, sigma=0.05, alpha=0.3),
    "cfg5": dict(net="afhq256", problem="random_inpainting", b_per_gpu=32, T=200, S=5, sigma=0.01, alpha=0.01),
    # S is not stated for cfg2-4 -> the reference default S=5 is the headline, S=1 (demo notebook) reported next to it
    "cfg4s1": dict(net="afhq256", problem="superresolution", b_per_gpu=16, T=100, S=1, sigma=0.05, alpha=0.3),
}
METRIC = "restored images/sec at 100 PnP steps, 256x256x3"
UNIT = "images/s"


def make_operator(P, problem, side):
    if problem == "inpainting":
        return P.BoxInpainting(20 if side == 128 else 40)
    if problem == "random_inpainting":
        return P.RandomInpainting(0.7)
    if problem == "superresolution":
        return P.Superresolution(2 if side == 128 else 4, side)
    if problem == "gaussian_deblurring_FFT":
        return P.GaussianDeblurring(1.0 if side == 128 else 3.0, 61, "fft", 3, side)
    if problem == "denoising":
        return P.Denoising()
    raise KeyError(problem)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [int(float(r[1])) for r in self.rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(self.rows), reasons=sorted(reasons))


def peaks():
    """Roofline denominators: burst bf16 for a kernel timed alone (per-op CUDA events), sustained bf16 for the whole step."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        burst = d.get("bf16_tflops")
        return dict(burst=burst, sustained=d.get("bf16_tflops_sustained", burst), hbm=d.get("hbm_gbs"),
                    src="measured (MEASURED_PEAKS.json: bf16_tflops burst for per-kernel numbers, bf16_tflops_sustained for the whole step, hbm_gbs)")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


def kernel_family(impl):
    """'rowconv<32,32,1> nsplit=..' -> ('rowconv', 'rowconv<32,32,1>'): kernel FUNCTION and template instantiation."""
    inst = impl.split(">")[0] + ">" if "<" in impl else impl.split(" ")[0]
    return inst.split("<")[0], inst


def ncu_traffic(kernel_prefix):
    """DRAM bytes per launch of a kernel from the newest committed `ncu --set full` summary (tools/summarize_profiles.py)."""
    for name in ("r02_kernel_dram_bytes.json", "r01_kernel_dram_bytes.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(tp):
            continue
        try:
            ks = json.load(open(tp))["kernels"]
        except Exception:
            continue
        hits = [v for k, v in ks.items() if k.startswith(kernel_prefix) and v.get("dram_bytes_per_launch_mean")]
        if hits:
            n = sum(h.get("launches", 1) for h in hits)
            return sum(h["dram_bytes_per_launch_mean"] * h.get("launches", 1) for h in hits) / n, name
    return None, None


def ncu_durations(prefix):
    """gpu__time_duration of the committed ncu capture of the per-pixel kernels (tools/ncu_pnp.py: cfg4 sizes, 16 images, S = 5)."""
    tp = os.path.join(ROOT, "profiles", "r02_pnp_pixel_ncu.json")
    if not os.path.exists(tp):
        return []
    try:
        return [float(l[k]) for l in json.load(open(tp))["launches"] if l.get("kernel", "").startswith(prefix)
                for k in l if k.startswith("gpu__time_duration")]
    except Exception:
        return []


def time_hbm_kernels(sess, x, y, pk, K=10):
    """The per-pixel kernels of one PnP step timed ALONE with CUDA events, L2 flushed (a 256 MB write)
    before every launch; achieved = algorithmic bytes / time against the measured copy bandwidth."""
    from pnpflow_b200 import _lib
    lib = _lib.load()
    S, n = sess.S, sess.n
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=x.device)
    sp = _lib.stream_ptr

    def timed(fn):
        tot = 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(K):
            flush.fill_(1)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / K
    opname = type(sess.op).__name__
    sf = getattr(sess.op, "sf", 1)
    by_datafit = {"Superresolution": (8 + 4.0 / (sf * sf)) * n, "GaussianDeblurring": 2 * 12.0 * n}.get(opname, 12.0 * n)
    sess.eps.normal_()
    out = []
    for name, fn, by in (
            ("datafit_step[" + opname + "]", lambda: sess.op.datafit_step(x, y, 0.5, out=sess.z), by_datafit),
            ("interp", lambda: _lib.check(lib.pnpf_interp(sess.z.data_ptr(), sess.eps.data_ptr(), 0.3, sess.zt.data_ptr(), n, S, sp())), (4.0 + 8.0 * S) * n),
            ("push_accum", lambda: _lib.check(lib.pnpf_push_accum(sess.zt.data_ptr(), sess.v.data_ptr(), 0.3, S, sess.xbuf[0].data_ptr(), n, sp())), (8.0 * S + 4.0) * n)):
        ms = timed(fn)
        tr, src = ncu_traffic({"datafit_step": "blur" if opname == "GaussianDeblurring" else "datafit_diag"}.get(name.split("[")[0], name))
        row = {"kernel": name, "ms": ms, "algorithmic_bytes": by, "achieved": by / (ms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
               "frac": by / (ms * 1e-3) / 1e9 / pk["hbm"], "traffic": tr, "traffic_source": src,
               "note": "one launch timed alone with CUDA events (includes ~2 us of launch latency on a 10-30 us kernel)"}
        # the same kernel under ncu (no launch latency in the number; only comparable at the captured size = cfg4's 16 images, S = 5)
        if sess.n == 16 * 3 * 256 * 256 and S == 5 and (opname == "Superresolution" or not name.startswith("datafit")):
            d = ncu_durations({"datafit_step": "datafit_diag", "interp": "interp", "push_accum": "push_accum"}[name.split("[")[0]])
            if d:
                row["ncu_duration_us"] = d[0]
                row["frac_at_ncu_duration"] = by / (d[0] * 1e-6) / 1e9 / pk["hbm"]
        if opname == "GaussianDeblurring" and name.startswith("datafit"):
            row["note"] = ("two launches of the separable circular 61-tap filter: 244 FMA per pixel out of shared memory — bound by the "
                           "shared-memory pipe, not by HBM; the HBM fraction is reported for completeness")
        out.append(row)
    return out


# ------------------------------------------------------------------------------------------------------------
# Baseline legs (the oracle port of the reference algorithm): cpu_reference / torch_gpu_reference below are the ONLY places
# bench.py executes oracle/, and only to time the reference next to the engine — the engine arm never touches it
# ------------------------------------------------------------------------------------------------------------
def cpu_reference(cfgname, steps, warmup):
    import oracle
    c = CONFIGS[cfgname]
    ocfg = oracle.AFHQ_256 if c["net"] == "afhq256" else oracle.CELEBA_128
    side = ocfg.input_height
    # oneDNN convolutions stop scaling (and collapse when oversubscribed) beyond ~16 threads at these batch sizes:
    # measured on the GPU box's 128-thread host: 0.22 s/eval at 16 threads vs 0.97 s at 64 and 61 s at 128.
    torch.set_num_threads(min(os.cpu_count() or 1, 16))
    cores = torch.get_num_threads()
    sd = oracle.init_state_dict(ocfg, seed=0)
    deg, sigma, alpha = oracle.make_degradation(c["problem"], side, 3, "cpu")
    Bs = 2 if side == 256 else 4                      # bounded sample
    g = torch.Generator().manual_seed(1234)
    clean = torch.rand(Bs, 3, side, side, generator=g) * 2 - 1
    y = oracle.loop.synthesize_measurement(clean, deg.H, sigma, 0)
    times = []

    def trace(it, x):
        times.append(time.perf_counter())

    t0 = time.perf_counter()
    times.append(t0)
    # steps_pnp=T keeps t = it/T small like the first iterations of the real run; only warmup+steps iterations execute
    class _Stop(Exception):
        pass

    def trace2(it, x):
        trace(it, x)
        if it + 1 >= warmup + steps:
            raise _Stop()
    try:
        oracle.pnp_flow_restore(lambda a, b: oracle.unet_forward(sd, ocfg, a, b), y, deg, sigma, steps_pnp=c["T"],
                                num_samples=c["S"], alpha=alpha, trace=trace2)
    except _Stop:
        pass
    per_step = (times[-1] - times[warmup]) / steps
    value = Bs / (c["T"] * per_step)
    sample = (f"{steps} PnP steps (after {warmup} warm-up) of {cfgname} at batch {Bs} with S={c['S']} on the host CPU, "
              f"extrapolated linearly to T={c['T']} steps")
    return dict(value=value, per_step_s=per_step, cores=cores, sample=sample, batch=Bs)


def torch_gpu_reference(cfgname, steps, warmup, dev):
    """The reference's own GPU path, like for like: the oracle port of PNP_FLOW.solve_ip + UNet.forward run
    EAGERLY in PyTorch on the same B200 (cuDNN convolutions with TF32 allowed = the reference's defaults, S sequential U-Net
    passes per step at the config's per-GPU batch).  A reported baseline next to cpu_baseline, never on the product path."""
    import oracle
    c = CONFIGS[cfgname]
    ocfg = oracle.AFHQ_256 if c["net"] == "afhq256" else oracle.CELEBA_128
    side, B = ocfg.input_height, c["b_per_gpu"]
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, False      # torch defaults = reference
    sd = {k: v.to(dev) for k, v in oracle.init_state_dict(ocfg, seed=0).items()}
    deg, sigma, alpha = oracle.make_degradation(c["problem"], side, 3, dev)
    g = torch.Generator().manual_seed(1234)
    clean = (torch.rand(B, 3, side, side, generator=g) * 2 - 1).to(dev)
    y = oracle.loop.synthesize_measurement(clean, deg.H, sigma, 0)
    ev = []

    class _Stop(Exception):
        pass

    def trace(it, x):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.append(e)
        if it + 1 >= warmup + steps:
            raise _Stop()
    try:
        oracle.pnp_flow_restore(lambda a, b: oracle.unet_forward(sd, ocfg, a, b), y, deg, sigma, steps_pnp=c["T"],
                                num_samples=c["S"], alpha=alpha, trace=trace)
    except _Stop:
        pass
    torch.cuda.synchronize()
    per_step_ms = ev[warmup - 1].elapsed_time(ev[-1]) / steps
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    del sd
    torch.cuda.empty_cache()
    return {"value": B / (c["T"] * per_step_ms / 1e3), "unit": UNIT, "ms_per_step": per_step_ms, "kind": "port",
            "what": "oracle port of the reference loop + U-Net, eager PyTorch fp32 on this GPU (cuDNN, TF32 allowed = reference defaults), "
                    f"batch {B}, S={c['S']} sequential passes per step",
            "sample": f"{steps} PnP steps after {warmup} warm-up, extrapolated linearly to T={c['T']}"}


# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--weights", default="recipe", choices=["recipe", "untouched"],
                    help="recipe: SURVEY §8d re-drawn init_scale=0 layers; untouched: the literal random init (gain 1e-10 there)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-shard-check", action="store_true")
    a = ap.parse_args()
    K, W = a.steps, max(a.warmup, 3 if a.impl == "engine" else 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    c = CONFIGS[a.config]
    cfg_desc = {"workload": f"{a.config}: {c['net']} {c['problem']} T={c['T']} S={c['S']} {c['b_per_gpu']} images/GPU",
                "global_batch": c["b_per_gpu"] * max(world, a.gpus if world == 1 else world), "parallelism": f"batch-sharded dp{world}",
                "weights": a.weights,
                "l2": "per-step working set (activations of the S*B U-Net batch, GBs) >> 126 MB L2; no explicit flush"}

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference(a.config, max(1, min(K, a.cpu_steps)), 1)   # bounded sample: <= cpu_steps CPU steps
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": max(1, min(K, a.cpu_steps)), "warmup": 1,
                "ms_per_step": r["per_step_s"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "impl": "reference", "config": cfg_desc,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    import torch.distributed as dist
    import pnpflow_b200 as P
    from pnpflow_b200 import sharding, synth
    assert torch.cuda.is_available(), "engine arm needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    net = synth.NETS[c["net"]]
    side, B, S, T = net["input_height"], c["b_per_gpu"], c["S"], c["T"]
    Btot = B * world
    if a.weights == "untouched":
        sd = synth.random_state_dict(net, seed=0, inner_gain=1e-10, end_gain=1e-10)     # models.py:84,137,212-216,432
    else:
        sd = synth.random_state_dict(net, seed=0)
    eng = P.UNetEngine(net, sd, device=dev, max_batch=S * B)
    op_full = make_operator(P, c["problem"], side)
    lo, hi = sharding.shard_bounds(Btot, world, rank)
    op = sharding.shard_operator(op_full, lo, hi, Btot)
    # measurements: rank 0 synthesises the full batch on its GPU and scatters the shards over NCCL (outside the step loop)
    if rank == 0:
        clean = synth.synthetic_clean(Btot, 3, side, 1234 + 4).to(dev)
        y_full = op_full.H(clean)
        torch.manual_seed(0)
        y_full = y_full + torch.randn_like(y_full) * c["sigma"]
        y_shape = tuple(y_full.shape)
    else:
        y_full = None
        y_shape = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world > 1:
            tv = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(tv, op=dist.ReduceOp.MAX)
            return float(tv.item())
        return v

    scatter_ms = scatter_first_ms = 0.0
    if world > 1:
        obj = [y_shape]
        dist.broadcast_object_list(obj, src=0)
        y_shape = obj[0]
        # first scatter = lazy NCCL communicator / channel setup (hundreds of ms, once per process); the steady-state
        # scatter a long-running job pays per batch is timed on the second call, on the device, max over ranks
        t_sc0 = time.perf_counter()
        y = sharding.scatter_batch(y_full, y_shape, torch.float32, dev)
        torch.cuda.synchronize()
        scatter_first_ms = (time.perf_counter() - t_sc0) * 1e3
        sharding.gather_batch(torch.zeros((hi - lo, 3, side, side), device=dev), Btot)       # warms the all-gather path too
        barrier()
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es0.record()
        y = sharding.scatter_batch(y_full, y_shape, torch.float32, dev)
        es1.record()
        torch.cuda.synchronize()
        scatter_ms = max_over_ranks(es0.elapsed_time(es1))
    else:
        y = y_full
    sess = P.PnPFlowSession(eng, op, tuple(y.shape), steps_pnp=T, lr_pnp=1.0, alpha=c["alpha"], num_samples=S, device=dev)
    full_shape = (Btot, 3, side, side)

    def noise_stream(seed):
        """The reference's noise: torch.randn_like of the FULL batch per draw (pnp_flow.py:48).  One rank: the session draws
        it itself (noise_it=None); sharded: every rank draws the full-batch tensor from the same seed and keeps its slice."""
        torch.manual_seed(seed)
        return iter(sharding.FullBatchNoise(full_shape, lo, hi, dev)) if world > 1 else None

    # ---------------- device-resident timing (value) ----------------
    nz = noise_stream(1)
    x = sess.initial_state(y)
    for i in range(W):
        x = sess.step(x, y, i, nz)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        x = sess.step(x, y, W + i, nz)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / K
    value = Btot / (T * ms_per_step / 1e3)
    finite = bool(torch.isfinite(x).all())

    # ---------------- end-to-end through the public API with HOST buffers (e2e) ----------------
    y_host = y.cpu().pin_memory()
    x_host = torch.empty(sess.shape, dtype=torch.float32).pin_memory()
    y_dev = torch.empty_like(y)
    xe = sess.initial_state(y)
    for i in range(3):
        y_dev.copy_(y_host, non_blocking=True)
        xe = sess.step(xe, y_dev, i, nz)
        x_host.copy_(xe, non_blocking=True)
    barrier()
    e0.record()
    for i in range(K):
        y_dev.copy_(y_host, non_blocking=True)          # H2D of the step's input (pinned)
        xe = sess.step(xe, y_dev, 3 + i, nz)
        x_host.copy_(xe, non_blocking=True)             # D2H of the step's result
    e1.record()
    barrier()
    ms_e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = Btot / (T * (ms_e / K) / 1e3)
    gather_ms, g_ms = 0.0, None
    if world > 1:
        g_ms = []
        for _ in range(3):                                   # steady state = the fastest of three (a cold proxy thread can cost tens of ms)
            barrier()
            e0.record()
            full = sharding.gather_batch(x, Btot)
            e1.record()
            torch.cuda.synchronize()
            g_ms.append(max_over_ranks(e0.elapsed_time(e1)))
        gather_ms = min(g_ms)
        assert full.shape[0] == Btot
    # a whole batch through the job: scatter + T steps (host buffers) + gather
    e2e_batch_value = Btot / ((T * (ms_e / K) + scatter_ms + gather_ms) / 1e3)

    # ---------------- sharded == unsharded (N > 1): 2 steps from the same seed, all ranks vs rank 0 alone ----------------
    shard_check = None
    if world > 1 and not a.no_shard_check:
        nz = noise_stream(77)
        xs = sess.initial_state(y)
        for i in range(2):
            xs = sess.step(xs, y, 50 + i, nz)
        full = sharding.gather_batch(xs, Btot)
        barrier()
        if rank == 0:
            try:
                sess1 = P.PnPFlowSession(eng, op_full, y_shape, steps_pnp=T, lr_pnp=1.0, alpha=c["alpha"], num_samples=S, device=dev)
                torch.manual_seed(77)
                x1 = sess1.initial_state(y_full)
                for i in range(2):
                    x1 = sess1.step(x1, y_full, 50 + i)              # world-1 semantics: randn_like of the full batch
                d = (x1 - full).abs().max().item()
                # noise floor of the comparison: the SAME unsharded run with the images in reversed order.  A different batch
                # composition changes how rows are split over CTAs, hence the grouping of the fp32 partial sums behind the
                # GroupNorm statistics, hence an occasional fp16 rounding flip — the only way two runs of the engine differ.
                idx = torch.arange(Btot - 1, -1, -1, device=dev)
                op_rev = sharding.shard_operator(op_full, 0, Btot, Btot)
                if hasattr(op_rev, "_host_mask"):
                    fwd_mask = op_rev._host_mask
                    op_rev._host_mask = lambda B_, H_, W_: fwd_mask(B_, H_, W_)[::-1].copy()
                y_rev = y_full.index_select(0, idx).contiguous()
                sess2 = P.PnPFlowSession(eng, op_rev, y_shape, steps_pnp=T, lr_pnp=1.0, alpha=c["alpha"], num_samples=S, device=dev)
                torch.manual_seed(77)
                g77 = [torch.randn(full_shape, device=dev).index_select(0, idx).contiguous() for _ in range(2 * S)]
                x2 = sess2.initial_state(y_rev)
                it77 = iter(g77)
                for i in range(2):
                    x2 = sess2.step(x2, y_rev, 50 + i, it77)
                floor = (x2.index_select(0, idx) - x1).abs().max().item()
                rel = ((x1 - full).norm() / x1.norm()).item()
                shard_check = {"what": "2 PnP steps, seed 77: gathered sharded result vs the same full batch run unsharded on rank 0; floor = the "
                                       "unsharded run repeated with the images in reversed order (same arithmetic, different CTA work split)",
                               "max_abs_diff": d, "rel_l2": rel, "reorder_floor_max_abs_diff": floor,
                               "consistent": bool(d <= max(3.0 * floor, 1e-3)),
                               "checksum_sharded": full.double().sum().item(), "checksum_unsharded": x1.double().sum().item()}
                del sess2
                del sess1
            except Exception as ex:
                shard_check = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}
        barrier()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- the U-Net evaluation alone (CUDA-graph replay), to split the step time ----------------
    _, _, _, replay = eng.graphed(S * B)
    replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        replay()
    e1.record()
    torch.cuda.synchronize()
    unet_replay_ms = e0.elapsed_time(e1) / 5

    # ---------------- rooflines ----------------
    # Per-op CUDA events of one U-Net evaluation (pnpf_profile_forward).  Every op has algorithmic FLOPs and algorithmic HBM
    # bytes (each operand read once, each output written once); its roof is the slower of FLOPs / BURST bf16 peak (a kernel timed
    # alone) and bytes / measured copy bandwidth.  Kernels are grouped by kernel FUNCTION (rowconv / patchconv / conv_gemm /
    # gn_apply ...); the per-instantiation list is kept under `instantiations`.  The headline `frac` is the WHOLE-STEP conv-GEMM
    # fraction (the north star's metric): images/s x T x S x F_conv / sustained bf16 peak.
    prof = eng.profile(S * B)
    pk = peaks()
    pk_fl, pk_by = pk["burst"] * 1e12, pk["hbm"] * 1e9

    def group(keyfn):
        fams = {}
        for o in prof:
            f = fams.setdefault(keyfn(o), dict(launches=0, ms=0.0, flops=0.0, bytes=0.0, roof_ms=0.0, tensor_roof_ms=0.0, hbm_roof_ms=0.0))
            f["launches"] += 1; f["ms"] += o["ms"]; f["flops"] += o["flops"]; f["bytes"] += o["bytes"]
            t_fl, t_by = o["flops"] / pk_fl * 1e3, o["bytes"] / pk_by * 1e3
            f["roof_ms"] += max(t_fl, t_by); f["tensor_roof_ms"] += t_fl; f["hbm_roof_ms"] += t_by
        tot = sum(f["ms"] for f in fams.values())
        rows = []
        for name, f in sorted(fams.items(), key=lambda kv: -kv[1]["ms"]):
            if f["ms"] <= 0:
                continue
            bound = "tensor" if f["tensor_roof_ms"] >= f["hbm_roof_ms"] else "hbm"
            rows.append({"kernel": name, "launches": f["launches"], "ms": f["ms"], "share": f["ms"] / tot, "bound": bound,
                         "tflops": f["flops"] / (f["ms"] * 1e-3) / 1e12, "gbs": f["bytes"] / (f["ms"] * 1e-3) / 1e9,
                         "frac_of_tensor_burst": f["tensor_roof_ms"] / f["ms"], "frac_of_hbm": f["hbm_roof_ms"] / f["ms"],
                         "frac_of_roof": f["roof_ms"] / f["ms"], "flops": f["flops"], "bytes": f["bytes"]})
        return rows, tot
    kernels, total_ms = group(lambda o: kernel_family(o.get("impl", "?"))[0])
    insts, _ = group(lambda o: kernel_family(o.get("impl", "?"))[1])
    dom = kernels[0]
    traffic, traffic_src = ncu_traffic(dom["kernel"])
    tc = [o for o in prof if o["kind"] == "tc"]
    simt = [o for o in prof if o["kind"] == "simt"]
    tc_ms, tc_fl = sum(o["ms"] for o in tc), sum(o["flops"] for o in tc)
    simt_ms, simt_by = sum(o["ms"] for o in simt), sum(o["bytes"] for o in simt)
    # whole step: executed tensor-core FLOPs of the engine plan (sub-pixel up convs execute 2.25x fewer MACs than the
    # reference's upsample + 3x3 form) and the reference's algorithmic conv FLOPs (F_conv)
    F_CONV_REF = {"afhq256": 188.37e9, "celeba128": 48.84e9}[c["net"]]
    evals_per_s = value * T * S / world
    ws_exec = evals_per_s * eng.flops_per_image / 1e12
    ws_ref = evals_per_s * F_CONV_REF / 1e12
    roofline = {"bound": "tensor", "kernel": "whole PnP step (all launches), conv-GEMM FLOPs of the reference U-Net",
                "achieved": ws_ref, "peak": pk["sustained"], "unit": "TFLOP/s", "frac": ws_ref / pk["sustained"],
                "frac_executed_flops": ws_exec / pk["sustained"], "achieved_executed": ws_exec,
                "flops_per_eval_reference": F_CONV_REF, "flops_per_eval_executed": eng.flops_per_image,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk["src"],
                "dominant_kernel": {"kernel": dom["kernel"], "launches": dom["launches"], "share_of_unet_eval": dom["share"], "bound": dom["bound"],
                                    "achieved": dom["tflops"] if dom["bound"] == "tensor" else dom["gbs"],
                                    "peak": pk["burst"] if dom["bound"] == "tensor" else pk["hbm"],
                                    "unit": "TFLOP/s" if dom["bound"] == "tensor" else "GB/s",
                                    "frac": dom["frac_of_tensor_burst"] if dom["bound"] == "tensor" else dom["frac_of_hbm"],
                                    "frac_of_tensor_burst": dom["frac_of_tensor_burst"], "frac_of_hbm": dom["frac_of_hbm"],
                                    "avg_launch_ms": dom["ms"] / dom["launches"],
                                    "algorithmic_flops_per_launch": dom["flops"] / dom["launches"],
                                    "algorithmic_bytes_per_launch": dom["bytes"] / dom["launches"],
                                    "traffic_dram_bytes_per_launch_ncu": traffic},
                "kernels": kernels, "instantiations": insts,
                "tensor_aggregate": {"kernel": "all tensor-core launches of one U-Net evaluation, per-op events, vs BURST peak",
                                     "achieved": tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else None, "peak": pk["burst"], "unit": "TFLOP/s",
                                     "frac": tc_fl / (tc_ms * 1e-3) / 1e12 / pk["burst"] if tc_ms > 0 else None,
                                     "launches": len(tc), "tc_ms_per_eval": tc_ms,
                                     "tc_share_of_eval": tc_ms / (tc_ms + simt_ms) if tc_ms + simt_ms > 0 else None},
                "simt": {"achieved_gbs": simt_by / (simt_ms * 1e-3) / 1e9 if simt_ms > 0 else None, "peak_gbs": pk["hbm"],
                         "ms_per_eval": simt_ms, "algorithmic_bytes": simt_by},
                "hbm_kernels": time_hbm_kernels(sess, x, y, pk),
                "whole_step_conv_gemm_frac": ws_ref / pk["sustained"]}

    cpu = None
    if not a.no_cpu_baseline and world == 1:          # reported baseline: rank 0 at N=1 only
        r = cpu_reference(a.config, a.cpu_steps, 1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    tgpu = None
    if not a.no_gpu_baseline and world == 1:
        try:
            tgpu = torch_gpu_reference(a.config, 3, 2, dev)
        except Exception as ex:                        # a reported extra: never fail the bench line over it
            tgpu = {"unavailable": f"{type(ex).__name__}: {ex}"[:200]}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "impl": "engine", "config": cfg_desc, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": y_host.numel() * 4, "d2h_bytes_per_step": x_host.numel() * 4,
                    "ms_per_step": ms_e / K},
            "e2e_batch": {"value": e2e_batch_value, "unit": UNIT,
                          "what": "one whole batch through the job: NCCL scatter of y + T steps with host buffers + NCCL gather of x"},
            "gpu_launches": K * sess.launches_per_step, "roofline": roofline, "cpu_baseline": cpu, "torch_gpu_baseline": tgpu,
            "finite": finite, "unet_flops_per_image": eng.flops_per_image, "unet_launches_per_eval": eng.num_launches,
            "unet_graph_replay_ms": unet_replay_ms, "scatter_ms": scatter_ms, "gather_ms": gather_ms,
            "scatter_first_call_ms": scatter_first_ms, "gather_ms_samples": g_ms if world > 1 else None, "noise": "torch Philox, full-batch randn per draw sliced per rank" if world > 1 else
            "torch Philox randn_like per draw", "shard_check": shard_check, "workspace_gb": eng.workspace_bytes / 2 ** 30}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
