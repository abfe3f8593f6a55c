"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference/src, imported
through oracle/diffusers_shim) on CPU in fp32 with seeded synthetic weights/inputs.

    python oracle/make_golden.py [case ...]

Only the small OUTPUT tensors (plus the seeds / shapes needed to regenerate weights and inputs deterministically) are
committed; weights are re-created on the test machine by aniportrait_b200.synthetic.randomize_state_dict with the same
seed (torch's CPU generator is deterministic for a given torch version; the fixture records torch.__version__).
"""
from __future__ import annotations

import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from aniportrait_b200.synthetic import randomize_state_dict  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def seeded_inputs_unet3d(B, Fr, h, w, chans, seed):
    """Inputs of one denoising-UNet call; shared verbatim by the GPU tests."""
    g = torch.Generator().manual_seed(seed)
    sample = torch.randn(1, 4, Fr, h, w, generator=g).repeat(B, 1, 1, 1, 1)
    clip = torch.randn(1, 768, generator=g)
    ehs = torch.cat([torch.zeros_like(clip), clip], 0).unsqueeze(1) if B == 2 else clip.unsqueeze(1)
    ref_lat = torch.randn(1, 4, h, w, generator=g)
    sizes = [(chans[0], h), (chans[0], h // 2), (chans[1], h // 4), (chans[2], h // 8), (chans[3], h // 8)]
    pose = [0.5 * torch.randn(1, c, Fr, s, s * w // h, generator=g).repeat(B, 1, 1, 1, 1) for c, s in sizes]
    return sample, ehs, ref_lat, pose


def _load(model, seed):
    sd = randomize_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd)
    return sd


def case_unet3d(name, chans, Fr, h, w, timestep, seeds=(101, 102, 103)):
    """ReferenceNet write pass + denoising UNet read pass under CFG, as pipeline_pose2vid_long.py:475-544 wires them."""
    ref_import.activate()
    from src.models.mutual_self_attention import ReferenceAttentionControl
    t0 = time.time()
    unet3d = ref_import.build_unet3d(chans)
    unet2d = ref_import.build_unet2d(chans)
    _load(unet3d, seeds[0])
    _load(unet2d, seeds[1])
    sample, ehs, ref_lat, pose = seeded_inputs_unet3d(2, Fr, h, w, chans, seeds[2])
    writer = ReferenceAttentionControl(unet2d, do_classifier_free_guidance=True, mode="write", batch_size=1,
                                       fusion_blocks="full")
    reader = ReferenceAttentionControl(unet3d, do_classifier_free_guidance=True, mode="read", batch_size=1,
                                       fusion_blocks="full")
    with torch.no_grad():
        unet2d(ref_lat.repeat(2, 1, 1, 1), torch.zeros((), dtype=torch.long), encoder_hidden_states=ehs,
               return_dict=False)
        reader.update(writer, dtype=torch.float32)
        out = unet3d(sample, torch.tensor(timestep), encoder_hidden_states=ehs, pose_cond_fea=pose,
                     return_dict=False)[0]
    torch.save(dict(case=name, chans=tuple(chans), frames=Fr, h=h, w=w, timestep=timestep, seeds=tuple(seeds),
                    out=out.float().contiguous(), torch_version=torch.__version__,
                    generator="reference src/models via oracle/diffusers_shim, fp32 CPU"),
               os.path.join(GOLDEN, name + ".pt"))
    print(f"{name}: out {tuple(out.shape)} |out|={out.norm():.4f} in {time.time() - t0:.1f}s")


CASES = {
    # full SD1.5 width (the real model size), 256x256-pixel equivalent latents, 4-frame window
    "unet3d_full_f4_32x32": lambda: case_unet3d("unet3d_full_f4_32x32", (320, 640, 1280, 1280), 4, 32, 32, 479),
    # reduced width, 16-frame window (temporal attention at the production window length), non-square latent
    "unet3d_small_f16_16x24": lambda: case_unet3d("unet3d_small_f16_16x24", (64, 128, 256, 256), 16, 16, 24, 959,
                                                   seeds=(111, 112, 113)),
}

if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 8)
    names = sys.argv[1:] or list(CASES)
    for n in names:
        CASES[n]()
