#!/bin/bash
# Validate and A/B the opt-in kernel paths written at the end of round 1 on one B200:
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/validate_optin.sh > gpurun_out/optin.log 2>&1; tail -40 gpurun_out/optin.log'
# Every step runs under its own `timeout` (a hanging kernel must not hold the box) and in its own process (the library reads
# the switches once).
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 240 "$@" 2>&1 | tail -6; }
# sub-pixel up convs: layer level first (no plan change), then the whole net
PNPF_TEST_SUBPIXEL=1 run python -m pytest tests/test_gpu_zz_subpixel_up.py -m gpu -q -x -k layer
PNPF_SUBPIXEL_UP=1 PNPF_TEST_SUBPIXEL=1 run python -m pytest tests/test_gpu_zz_subpixel_up.py -m gpu -q -x -k unet
# three taps per weight slot: the existing patch-kernel and U-Net tests with the switch on
PNPF_PATCH_TG=3 run python -m pytest tests/test_gpu_layers.py tests/test_gpu_unet.py -m gpu -q -x
# fused GroupNorm / concat inputs in the patch kernel
PNPF_PATCH_GN=1 PNPF_TEST_PATCH_GN=1 run python -m pytest tests/test_gpu_zz_patchgn.py -m gpu -q -x
PNPF_PATCH_GN=1 PNPF_PATCH_TG=3 PNPF_TEST_PATCH_GN=1 run python -m pytest tests/test_gpu_zz_patchgn.py -m gpu -q -x
# three row slots for the unsplit row-kernel layout (host-side choice only: same kernel)
PNPF_ROW_MINSLOT=3 run python -m pytest tests/test_gpu_unet.py -m gpu -q -x
# same-box A/B (only meaningful for the paths that passed above)
echo "=== A/B"
AB_ROUNDS=2 timeout 400 python tools/ab_env.py "base:" "minslot3:PNPF_ROW_MINSLOT=3" "subpix:PNPF_SUBPIXEL_UP=1" "tg3:PNPF_PATCH_TG=3" "gn256:PNPF_PATCH_GN=256" "gn:PNPF_PATCH_GN=1" \
    "gn+tg3:PNPF_PATCH_GN=1,PNPF_PATCH_TG=3" "tg3+subpix:PNPF_PATCH_TG=3,PNPF_SUBPIXEL_UP=1" \
    "all256:PNPF_PATCH_TG=3,PNPF_SUBPIXEL_UP=1,PNPF_PATCH_GN=256" "all:PNPF_PATCH_TG=3,PNPF_SUBPIXEL_UP=1,PNPF_PATCH_GN=1" 2>&1 | grep -v Warning | tail -24
