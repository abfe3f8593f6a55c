"""Dev helper: synchronising wall-clock split of the per-video (non-denoise) phase of run_device at the bench config."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from aniportrait_b200 import _lib

dev = torch.device("cuda:0")
_lib.init(0)
pipe = bench.build_product_pipeline(dev)
ref_image, poses, ref_pose = bench.synthetic_inputs(1000, bench.L)
clip_pixels = pipe.clip_image_processor.preprocess(ref_image.resize((224, 224)), return_tensors="pt").pixel_values.to(dev, torch.float16)
ref_t = pipe.ref_image_processor.preprocess(ref_image, height=bench.H, width=bench.W).to(dev, torch.float16)
pose_t = torch.cat([pipe.cond_image_processor.preprocess(p, height=bench.H, width=bench.W) for p in poses], 0).to(dev, torch.float16)
lat0 = torch.randn((1, 4, bench.L, bench.H // 8, bench.W // 8), dtype=torch.float16).to(dev)
for i in range(3):
    pipe.profile_phases = i == 2
    pipe.run_device(clip_pixels, ref_t, pose_t, lat0, 2, bench.GUIDANCE)
    torch.cuda.synchronize()
print(json.dumps({k: round(v, 2) for k, v in pipe.phase_detail.items()}), pipe.collect_timings())
