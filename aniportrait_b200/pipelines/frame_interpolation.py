"""`-acc` frame interpolation around the FILM TorchScript model (reference src/utils/frame_interpolation.py:12-69,
called by scripts/pose2vid.py:178-179 on `pipe(...).videos`).

The reference walks the video pair by pair; for every pair it inserts `inter_frames` frames one model call at a time
(batch 1), moving both inputs to the GPU and the prediction back to the host as fp32 for every call:
(F-1)*inter_frames launches of the whole network, 2 H2D + 1 D2H copies each.

Which frame is synthesised from which two neighbours, in which order and at which `dt`, depends on `inter_frames` only —
never on the data and never on the pair. So the order is computed ONCE (`insertion_schedule`) and every step of it runs for
ALL pairs of the video as one batch: `inter_frames` model calls per video instead of (F-1)*inter_frames, frames stay on the
device between the steps (the reference's prediction -> fp32 host -> fp16 device round trip changes no value), one host copy
at the end. The network itself is the user's TorchScript file (`./pretrained_model/film_net_fp16.pt`, not part of
the reference repository): opaque library code, called with the batch layout its exporter documents (x0, x1 `[n, 3, H, W]`,
dt `[n, 1]`).

Parity: `tests/test_host_cpu.py::test_frame_interpolation_matches_reference_order_and_values` runs the unmodified
reference function and this one on a stand-in network (the real weights are absent: parity with the real FILM file is
unpinned) and requires identical frames, for inter_frames 1..5.
"""
from __future__ import annotations

import bisect
import os
from typing import List, Tuple

import torch


def init_frame_interpolation_model(checkpoint_name: str = None, device: str = "cuda"):
    """Reference :12-20: TorchScript FILM in fp16 on the GPU."""
    checkpoint_name = checkpoint_name or os.path.join("./pretrained_model/film_net_fp16.pt")
    model = torch.jit.load(checkpoint_name, map_location="cpu")
    model.eval()
    return model.half().to(device=device)


def insertion_schedule(inter_frames: int) -> List[Tuple[int, int, int, torch.Tensor]]:
    """The order in which the reference fills the `inter_frames` slots between two frames (reference :31-60).

    Slots are numbered 0 (first frame) .. inter_frames + 1 (second frame), at times linspace(0, 1). At every step the
    reference looks at all (gap between two already known neighbours, still empty slot) combinations and takes the one whose
    slot lies closest to the middle of its gap (first one on ties, in gap-major order). Returns, per step,
    (left slot, right slot, new slot, dt) with dt the fp16 `[1, 1]` tensor the network is given: the reference builds it as
    fp16(t_new - t_left) / (t_right - t_left) with the divisor rounded to fp16 by type promotion; the same two torch
    operations are evaluated here so that the value is the same bit pattern."""
    n = int(inter_frames)
    if n < 1:
        return []
    times = torch.linspace(0, 1, n + 2)                    # float32, like the reference
    known = [0, n + 1]                                      # sorted slots that hold a frame
    empty = list(range(1, n + 1))
    steps = []
    while empty:
        left_t = times[known[:-1]][:, None]                 # [gaps, 1]
        right_t = times[known[1:]][:, None]
        off_centre = ((times[empty][None, :] - left_t) / (right_t - left_t) - 0.5).abs()     # [gaps, empty]
        flat = int(torch.argmin(off_centre))                # first minimum in gap-major order
        gap, which = divmod(flat, len(empty))
        left, right, new = known[gap], known[gap + 1], empty[which]
        dt = torch.full((1, 1), float(times[new] - times[left]), dtype=torch.float16) / (times[right] - times[left])
        steps.append((left, right, new, dt))
        known.insert(bisect.bisect_left(known, new), new)
        del empty[which]
    return steps


def _model_device(model, default="cuda"):
    try:
        return next(model.parameters()).device
    except (StopIteration, AttributeError, RuntimeError):
        return torch.device(default)


@torch.no_grad()
def batch_images_interpolation_tool(input_tensor: torch.Tensor, model, inter_frames: int = 1,
                                    max_pairs_per_call: int = 32, device=None) -> torch.Tensor:
    """input_tensor [bs, c, F, H, W] in [0, 1] -> fp32 CPU tensor [bs, c, (F-1)*(inter_frames+1)+1, H, W]
    (reference signature :23; same frames in the same order). `max_pairs_per_call` bounds the batch of one network call;
    `device` defaults to where the network's parameters live ("cuda" for a parameter-less module, like the reference)."""
    if input_tensor.dim() != 5:
        raise ValueError(f"expected a video tensor [bs, c, frames, h, w], got {tuple(input_tensor.shape)}")
    n = int(inter_frames)
    bs, c, frames, h, w = input_tensor.shape
    if frames < 2 or n < 1:
        return input_tensor.detach().cpu().float()
    device = torch.device(device) if device is not None else _model_device(model)
    src = input_tensor.detach().to(device)                  # the given frames are handed through unrounded (reference :62-65)
    pairs = frames - 1
    # slots[s]: [pairs * bs, c, h, w] — the frame at slot s of every pair (pair-major, then the video batch)
    slots = {0: src[:, :, :-1].permute(2, 0, 1, 3, 4).reshape(pairs * bs, c, h, w),
             n + 1: src[:, :, 1:].permute(2, 0, 1, 3, 4).reshape(pairs * bs, c, h, w)}
    chunk = max(1, int(max_pairs_per_call)) * bs
    for left, right, new, dt in insertion_schedule(n):
        x0, x1 = slots[left], slots[right]
        dt = dt.to(device)
        outs = []
        for i in range(0, x0.shape[0], chunk):
            a = x0[i:i + chunk].to(torch.float16).contiguous()          # the network sees fp16 inputs (reference :47-50)
            b = x1[i:i + chunk].to(torch.float16).contiguous()
            outs.append(model(a, b, dt.expand(a.shape[0], 1).contiguous()).clamp(0, 1))
        slots[new] = outs[0] if len(outs) == 1 else torch.cat(outs)
    # interleave: pair p contributes slots 0..n, the last frame of the video closes the sequence
    per_pair = torch.stack([slots[s].float() for s in range(n + 1)], dim=1)            # [pairs*bs, n+1, c, h, w]
    per_pair = per_pair.view(pairs, bs, n + 1, c, h, w).permute(1, 3, 0, 2, 4, 5).reshape(bs, c, pairs * (n + 1), h, w)
    return torch.cat([per_pair, src[:, :, -1:].float()], dim=2).cpu()
