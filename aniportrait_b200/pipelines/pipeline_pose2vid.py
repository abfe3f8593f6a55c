"""Pose2VideoPipeline, single-window variant — mirror of the reference's src/pipelines/pipeline_pose2vid.py (used by the
stage-2 training validator, train_stage_2.py:164-223; clips of up to 24 frames).

Differences to the long pipeline, as in the reference: ALL `video_length` frames go through the denoising UNet as one
window (temporal attention spans the whole clip, no overlap averaging; pipeline_pose2vid.py:396-430), PoseGuider runs once
on the whole clip (:396-399), and the CLIP processor receives the portrait un-squashed (:320-322). Everything runs on the
same device core (cached CUDA-graph session, kernels) as pipeline_pose2vid_long.Pose2VideoPipeline.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Union

import torch

from .pipeline_pose2vid_long import Pose2VideoPipeline as _LongPipeline
from .pipeline_pose2vid_long import Pose2VideoPipelineOutput  # noqa: F401


class Pose2VideoPipeline(_LongPipeline):
    @torch.no_grad()
    def __call__(self, ref_image, pose_images, ref_pose_image, width, height, video_length, num_inference_steps,
                 guidance_scale, num_images_per_prompt=1, eta: float = 0.0,
                 generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
                 output_type: Optional[str] = "tensor", return_dict: bool = True,
                 callback: Optional[Callable[[int, int, torch.Tensor], None]] = None,
                 callback_steps: Optional[int] = 1, **kwargs):
        """Reference signature (pipeline_pose2vid.py:287-305)."""
        if video_length > 32:
            raise ValueError("the single-window pipeline is limited by the temporal positional encoding (max_len 32); "
                             "use pipeline_pose2vid_long for longer clips")
        return super().__call__(ref_image, pose_images, ref_pose_image, width, height, video_length,
                                num_inference_steps, guidance_scale, num_images_per_prompt, eta, generator, output_type,
                                return_dict, callback, callback_steps, context_schedule="uniform",
                                context_frames=max(int(video_length), 1), context_stride=1, context_overlap=0,
                                clip_resize=False, **kwargs)
