#!/bin/bash
# GPU-box dev run: A/B of the round-2 switches on ONE box (separate processes: the switches are read once)
mkdir -p gpurun_out
{
echo "== UNet3D call: normalisation fusion on / off, PDL on, single-buffered epilogue staging"
AP_FUSE_NORMS=1 timeout 400 python scripts/dev_time_unet.py 2>&1 | tail -2
AP_FUSE_NORMS=0 timeout 400 python scripts/dev_time_unet.py 2>&1 | tail -2
AP_PDL=1 timeout 400 python scripts/dev_time_unet.py 2>&1 | tail -2
AP_GEMM_EPI_DOUBLE=0 timeout 400 python scripts/dev_time_unet.py 2>&1 | tail -2
echo "== reference attention generations"
bash scripts/dev_attn_ab.sh
} > gpurun_out/r02_ab.log 2>&1
tail -40 gpurun_out/r02_ab.log
