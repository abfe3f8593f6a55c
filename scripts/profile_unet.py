"""GPU-box helper for ncu: builds the full-size denoising UNet, warms up, then runs ONE call between
cudaProfilerStart/Stop (use `ncu --profile-from-start off ...`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from aniportrait_b200 import ops  # noqa: E402
from aniportrait_b200.models import ReferenceAttentionControl  # noqa: E402

dev = torch.device("cuda:0")
F = int(os.environ.get("F", 16))
H = int(os.environ.get("H", 64))
pipe = bench.build_product_pipeline(dev)
unet3d, unet2d = pipe.denoising_unet, pipe.reference_unet
writer = ReferenceAttentionControl(unet2d, do_classifier_free_guidance=True, mode="write", fusion_blocks="full")
reader = ReferenceAttentionControl(unet3d, do_classifier_free_guidance=True, mode="read", fusion_blocks="full")
clip = torch.randn(1, 768, device=dev, dtype=torch.float16)
ehs = torch.cat([torch.zeros_like(clip), clip]).unsqueeze(1)
with torch.no_grad():
    unet2d(torch.randn(2, 4, H, H, device=dev, dtype=torch.float16), torch.zeros((), device=dev),
           encoder_hidden_states=ehs)
    reader.update(writer)
    x = ops.ncfhw_to_nhwc(torch.randn(2, 4, F, H, H, device=dev, dtype=torch.float16), 64)
    pose = [torch.randn(F, H // s, H // s, c, device=dev, dtype=torch.float16) * 0.1
            for c, s in [(320, 1), (320, 2), (640, 4), (1280, 8), (1280, 8)]]
    tt = torch.tensor([500.0], device=dev)
    unet3d.prepare_reference(2, F, ehs, ehs_key="profile")     # per-video constants, as the pipeline does once per video
    for _ in range(2):
        unet3d.forward_nhwc(x, 2, F, tt, ehs, pose, ehs_key="profile")
    torch.cuda.synchronize()
    if os.environ.get("AP_SHAPE_LOG"):
        import json
        ops.SHAPE_LOG = []
    torch.cuda.profiler.start()
    unet3d.forward_nhwc(x, 2, F, tt, ehs, pose, ehs_key="profile")
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    if ops.SHAPE_LOG is not None:
        json.dump(ops.SHAPE_LOG, open(os.environ["AP_SHAPE_LOG"], "w"))
print("done")
