#!/bin/bash
# Round evidence (run under gpurun on one B200): launch list of the bench command + `ncu --set full` captures of the
# tensor-core kernels and gn_apply.  The raw metric pages are exported to CSV ON THE BOX and the .ncu-rep files deleted unless
# KEEP_REP=1 (gpurun copies back at most 64 MiB; a capture with --import-source is ~17 MB).  Outputs land in gpurun_out/;
# tools/summarize_profiles.py turns them into profiles/<tag>_*.json.
TAG=${1:-r02}
SRC=${KEEP_REP:+--import-source on}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
cap() {   # name, kernel regex, first launch, count
    ncu --set full --clock-control none $SRC -k regex:$2 -s $3 -c $4 -o gpurun_out/${TAG}_$1 -f python tools/ncu_unet.py > gpurun_out/${TAG}_ncu_$1.log 2>&1
    ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1.raw.csv 2>/dev/null
    [ -z "$KEEP_REP" ] && rm -f gpurun_out/${TAG}_$1.ncu-rep
}
# row kernel: launches 1.. = level-0 down blocks (fused GroupNorm 32->32, identity shortcut), 26.. = level-1 up blocks (split C_out)
cap rowconv_l0 rowconv 1 2
cap rowconv_l1up rowconv 26 2
# patch kernel: launches 1.. = level-2 down blocks (128->128), 13.. = level 3 (256->256, CTA pairs)
cap patchconv_l2 patchconv 1 2
cap patchconv_l3 patchconv 13 2
cap gn_apply gn_apply 2 2
# per-pixel PnP kernels (data-fidelity step of the BASELINE operators, interpolate, push + average): tools/ncu_pnp.py
ncu --set full --clock-control none -k regex:"datafit|interp|push_accum|blur|sr_bicubic" -o gpurun_out/${TAG}_pnp_pixel -f python tools/ncu_pnp.py > gpurun_out/${TAG}_ncu_pnp_pixel.log 2>&1
ncu -i gpurun_out/${TAG}_pnp_pixel.ncu-rep --page raw --csv > gpurun_out/${TAG}_pnp_pixel.raw.csv 2>/dev/null
[ -z "$KEEP_REP" ] && rm -f gpurun_out/${TAG}_pnp_pixel.ncu-rep
ls -la gpurun_out/ | head -30
