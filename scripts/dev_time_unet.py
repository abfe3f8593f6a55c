"""Dev helper (GPU box): time one full-size denoising-UNet call (512x512, F=16, CFG) with random weights."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from aniportrait_b200 import ops  # noqa: E402
from aniportrait_b200.models import ReferenceAttentionControl  # noqa: E402
from helpers import MOTION_KWARGS  # noqa: E402
from aniportrait_b200.models import UNet2DConditionModel, UNet3DConditionModel  # noqa: E402

dev = torch.device("cuda:0")
F = int(os.environ.get("F", 16))
H = int(os.environ.get("H", 64))
torch.manual_seed(0)


def rand_init(m):
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() > 1:
                fan_in = p[0].numel()
                p.copy_(torch.randn_like(p) * min(0.05, fan_in ** -0.5))
            elif "norm" in n and n.endswith("weight"):
                p.fill_(1.0)
            else:
                p.zero_()


t0 = time.time()
with torch.device(dev):
    unet3d = UNet3DConditionModel(sample_size=64, cross_attention_dim=768, attention_head_dim=8,
                                  use_inflated_groupnorm=True, unet_use_cross_frame_attention=False,
                                  unet_use_temporal_attention=False, use_motion_module=True,
                                  motion_module_resolutions=[1, 2, 4, 8], motion_module_mid_block=True,
                                  motion_module_type="Vanilla", motion_module_kwargs=dict(MOTION_KWARGS)).half()
    unet2d = UNet2DConditionModel(sample_size=64, cross_attention_dim=768, attention_head_dim=8).half()
rand_init(unet3d)
rand_init(unet2d)
print(f"models built in {time.time() - t0:.1f}s")
writer = ReferenceAttentionControl(unet2d, do_classifier_free_guidance=True, mode="write", fusion_blocks="full")
reader = ReferenceAttentionControl(unet3d, do_classifier_free_guidance=True, mode="read", fusion_blocks="full")
clip = torch.randn(1, 768, device=dev, dtype=torch.float16)
ehs = torch.cat([torch.zeros_like(clip), clip]).unsqueeze(1)
ref_lat = torch.randn(1, 4, H, H, device=dev, dtype=torch.float16)
with torch.no_grad():
    unet2d(ref_lat.repeat(2, 1, 1, 1), torch.zeros((), device=dev), encoder_hidden_states=ehs)
    reader.update(writer)
    x = ops.ncfhw_to_nhwc(torch.randn(2, 4, F, H, H, device=dev, dtype=torch.float16), 64)
    pose = [torch.randn(F, H // s, H // s, c, device=dev, dtype=torch.float16) * 0.1
            for c, s in [(320, 1), (320, 2), (640, 4), (1280, 8), (1280, 8)]]
    unet3d.prepare_reference(2, F, ehs, ehs_key="dev")
    for it in range(3):
        out = unet3d.forward_nhwc(x, 2, F, 500.0, ehs, pose, ehs_key="dev")
    torch.cuda.synchronize()
    n0 = ops.KERNEL_LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    tw = time.time()
    e0.record()
    for it in range(iters):
        out = unet3d.forward_nhwc(x, 2, F, 500.0, ehs, pose, ehs_key="dev")
    e1.record()
    torch.cuda.synchronize()
    tw = time.time() - tw
ms = e0.elapsed_time(e1) / iters
if os.environ.get("GRAPH", "1") == "1":
    tt = torch.tensor([500.0], device=dev)
    with torch.no_grad():
        sstream = torch.cuda.Stream()
        with torch.cuda.stream(sstream):
            out = unet3d.forward_nhwc(x, 2, F, tt, ehs, pose, ehs_key="dev")
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = unet3d.forward_nhwc(x, 2, F, tt, ehs, pose, ehs_key="dev")
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for it in range(iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    print(f"CUDA-graph replay: {e0.elapsed_time(e1) / iters:.2f} ms per UNet call")
flops = 36.43e12 * (F / 16) * (H / 64) ** 2
print(f"UNet3D call F={F} {H}x{H}: {ms:.2f} ms (wall {tw / iters * 1e3:.2f} ms), launches/call={(ops.KERNEL_LAUNCHES - n0) // iters}, "
      f"{flops / ms / 1e9:.1f} TFLOP/s algorithmic, finite={torch.isfinite(out.float()).all().item()}, "
      f"mem={torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
