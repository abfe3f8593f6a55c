#!/bin/bash
# GPU-box: ncu launch lists (per-kernel durations) of ONE UNet3D call with / without the normalisation fusion
mkdir -p gpurun_out
for f in ${FUSE_LIST:-1 0}; do
  AP_FUSE_NORMS=$f AP_SHAPE_LOG=gpurun_out/r02_shapes_fuse$f.json timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/r02_unet3d_call_launches_fuse$f.csv python scripts/profile_unet.py > gpurun_out/r02_prof_fuse$f.log 2>&1
  python scripts/summarize_launches.py gpurun_out/r02_unet3d_call_launches_fuse$f.csv > gpurun_out/r02_unet3d_call_launches_fuse$f.summary.txt 2>&1
  echo "== fuse=$f"; head -30 gpurun_out/r02_unet3d_call_launches_fuse$f.summary.txt
done
