"""One-process GPU check of the 8-bit frame output (tests + host-copy timing at 16 x 512 x 512); dev aid, not a test."""
import sys
import time
import types

import pytest
import torch

rc = pytest.main(["tests/test_ops_gpu.py", "tests/test_pipeline_gpu.py", "-q", "-x", "-k",
                  "pack_frames or no_cfg_single"])
print("pytest rc", int(rc), flush=True)
from aniportrait_b200.pipelines.pipeline_pose2vid_long import Pose2VideoPipeline  # noqa: E402

holder = types.SimpleNamespace()
video = torch.rand(16, 3, 512, 512, device="cuda").half().view(1, 16, 3, 512, 512).permute(0, 2, 1, 3, 4)
for name in ("_to_host_f32", "_to_host_u8"):
    fn = getattr(Pose2VideoPipeline, name)
    fn(holder, video)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out = fn(holder, video)
    ms = (time.perf_counter() - t0) * 100
    print(f"{name}: {ms:.2f} ms per 16-frame clip, {out.numel() * out.element_size() / 1e6:.1f} MB to the host", flush=True)
sys.exit(int(rc))
