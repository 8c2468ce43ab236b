"""Turn gpurun_out/ captures into the tracked summaries under profiles/ (run in the build container; needs `ncu` for .ncu-rep)."""
import csv, json, os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(P, exist_ok=True)

def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("pnpf::", "").replace("void ", "").strip()

# ---- launch list (ncu --metrics gpu__time_duration.sum): share of time per kernel
lp = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if l.startswith('"'))]
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        k = short(r[ki]); agg[k][0] += 1; agg[k][1] += float(r[vi].replace(",", "")) / 1e3
    tot = sum(v[1] for v in agg.values())
    out = dict(command="ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline",
               note="cold-cache, serialised per-launch times: compare SHARES, not absolutes", total_us=tot, launches=len(rows) - 1,
               kernels=[dict(kernel=k, launches=v[0], us=round(v[1], 1), share=round(v[1] / tot, 4)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])])
    json.dump(out, open(os.path.join(P, f"{tag}_launch_summary.json"), "w"), indent=1)
    import gzip, shutil
    with open(lp, "rb") as fi, gzip.open(os.path.join(P, f"{tag}_launches.csv.gz"), "wb") as fo:
        shutil.copyfileobj(fi, fo)
    print("launch summary:", [(d["kernel"][:40], d["share"]) for d in out["kernels"][:8]])

# ---- full captures
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
def family(kname):
    """ncu kernel name -> the family name bench.py uses (rowconv<BK,BN,KCH>, patchconv<BN>, conv_gemm<BK,BN>, gn_apply, ...)"""
    m = re.match(r"(\w+?)(_kernel)?<([^>]*)>", kname)
    if not m:
        return kname.replace("_kernel", "")
    base, args = m.group(1), [a.strip().replace("(int)", "").replace("(bool)", "") for a in m.group(3).split(",")]
    keep = {"rowconv": 3, "patchconv": 1, "conv_gemm": 2}.get(base, 0)
    return f"{base}<{','.join(args[:keep])}>" if keep else base

dram = collections.defaultdict(list)
for f in sorted(os.listdir(G)):
    if not f.startswith(tag): continue
    if f.endswith(".raw.csv"):                       # exported on the GPU box by tools/capture_profiles.sh
        raw = open(os.path.join(G, f)).read()
        f = f.replace(".raw.csv", ".ncu-rep")
    elif f.endswith(".ncu-rep") and not os.path.exists(os.path.join(G, f.replace(".ncu-rep", ".raw.csv"))):
        raw = subprocess.run(["ncu", "-i", os.path.join(G, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        continue
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3: continue
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": short(vals[hdr.index("Kernel Name")]) if "Kernel Name" in hdr else "?"}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS or any(h.endswith(k) for k in KEYS):
                d[h.split(".TriageCompute.")[-1] + (f" [{u}]" if u else "")] = v
        res.append(d)
        try:
            rd = [float(v.replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[k.split("[")[1].rstrip("]")]
                  for k, v in d.items() if k.startswith("dram__bytes_read.sum") or k.startswith("dram__bytes_write.sum")]
            if len(rd) == 2:
                dram[family(d["kernel"])].append(sum(rd))
        except Exception:
            pass
    name = f if f.startswith(tag) else f"{tag}_{f}"
    json.dump(dict(source=f, command="ncu --set full --clock-control none --import-source on (see tools/, DESIGN.md §5)", launches=res),
              open(os.path.join(P, name.replace(".ncu-rep", "_ncu.json")), "w"), indent=1)
    print(f, [(d["kernel"][:30], d.get("gpu__time_duration.sum [us]")) for d in res])
if dram:
    json.dump(dict(source=f"{tag}_*_ncu.json (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                   kernels={k: dict(dram_bytes_per_launch=v, dram_bytes_per_launch_mean=sum(v) / len(v)) for k, v in dram.items()}),
              open(os.path.join(P, f"{tag}_kernel_dram_bytes.json"), "w"), indent=1)
# parity numbers written by tests/test_gpu_parity_baseline.py during the same gpurun call (fresh file only: gpurun_out is scratch)
for src, dst in (("parity_baseline.json", f"{tag}_parity_baseline.json"),):
    sp = os.path.join(G, src)
    if os.path.exists(sp) and "--with-parity" in sys.argv:
        json.dump(json.load(open(sp)), open(os.path.join(P, dst), "w"), indent=1)
