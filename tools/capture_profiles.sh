#!/bin/bash
# Round evidence (run under gpurun on one B200): launch list of the bench command + `ncu --set full` captures of the three
# tensor-core kernels.  Outputs land in gpurun_out/; tools/summarize_profiles.py turns them into profiles/<tag>_*.json.
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
# row kernel: launches 1.. = level-0 down blocks (fused GroupNorm 32->32, identity shortcut), 26.. = level-1 up blocks (split C_out)
ncu --set full --clock-control none --import-source on -k regex:rowconv -s 1 -c 2 -o gpurun_out/${TAG}_rowconv_l0 -f python tools/ncu_unet.py > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rowconv -s 26 -c 2 -o gpurun_out/${TAG}_rowconv_l1up -f python tools/ncu_unet.py > gpurun_out/${TAG}_ncu2.log 2>&1
# patch kernel: launches 0.. = level-2 down blocks (128->128), 12.. = level 3 (256->256, CTA pairs)
ncu --set full --clock-control none --import-source on -k regex:patchconv -s 1 -c 2 -o gpurun_out/${TAG}_patchconv_l2 -f python tools/ncu_unet.py > gpurun_out/${TAG}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:patchconv -s 13 -c 2 -o gpurun_out/${TAG}_patchconv_l3 -f python tools/ncu_unet.py > gpurun_out/${TAG}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_apply -s 2 -c 2 -o gpurun_out/${TAG}_gn_apply -f python tools/ncu_unet.py > gpurun_out/${TAG}_ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep
