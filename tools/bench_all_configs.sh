#!/bin/bash
# Every BASELINE.json GPU configuration (plus the S = 1 and untouched-init variants SURVEY §8d asks for) on ONE B200, one complete
# JSON line each -> gpurun_out/<tag>_bench_lines.jsonl     (gpurun --timeout 1500 -- 'bash tools/bench_all_configs.sh r02')
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_bench_lines.jsonl
mkdir -p gpurun_out; : > $OUT
run() { echo "=== bench.py $*" >&2; timeout 600 python bench.py "$@" 2>gpurun_out/${TAG}_bench_err.log | tail -1 >> $OUT; }
run --config cfg4
run --config cfg4s1 --no-cpu-baseline
run --config cfg4 --weights untouched --no-cpu-baseline --no-gpu-baseline
run --config cfg2 --no-cpu-baseline
run --config cfg3 --no-cpu-baseline
run --config cfg5 --steps 20 --no-cpu-baseline
python - <<PY
import json
for l in open("$OUT"):
    try:
        b = json.loads(l)
        r = b["roofline"]
        print(b["config"]["workload"], "| weights", b["config"]["weights"], "| %.3f img/s, %.2f ms/step, e2e %.3f, conv-GEMM frac %.3f (executed %.3f), clocks %s" % (
            b["value"], b["ms_per_step"], b["e2e"]["value"], r["frac"], r["frac_executed_flops"], b["clocks"]))
        print("    dominant:", r["dominant_kernel"]["kernel"], "share %.2f frac %.2f" % (r["dominant_kernel"]["share_of_unet_eval"], r["dominant_kernel"]["frac"]),
              "| hbm kernels:", [(k["kernel"], round(k["frac"], 2)) for k in r["hbm_kernels"]],
              "| torch gpu baseline:", (b.get("torch_gpu_baseline") or {}).get("value"))
    except Exception as e:
        print("unparsable line:", e, l[:200])
PY
