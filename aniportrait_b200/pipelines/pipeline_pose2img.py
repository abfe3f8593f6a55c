"""Pose2ImagePipeline — mirror of the reference's src/pipelines/pipeline_pose2img.py (stage-1 training validator,
train_stage_1.py; one pose map -> one image). A one-frame clip through the same device core as the video pipelines: the
latents are `[1, 4, 1, h, w]` (reference :268), the result object carries `.images` `[1, 3, 1, H, W]` (:361-372)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Union

import numpy as np
import torch

from .pipeline_pose2vid_long import Pose2VideoPipeline as _LongPipeline


@dataclass
class Pose2ImagePipelineOutput:
    images: Union[torch.Tensor, np.ndarray]


class Pose2ImagePipeline(_LongPipeline):
    @torch.no_grad()
    def __call__(self, ref_image, pose_image, ref_pose_image, width, height, num_inference_steps, guidance_scale,
                 num_images_per_prompt=1, eta: float = 0.0,
                 generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
                 output_type: Optional[str] = "tensor", return_dict: bool = True,
                 callback: Optional[Callable[[int, int, torch.Tensor], None]] = None,
                 callback_steps: Optional[int] = 1, **kwargs):
        """Reference signature (pipeline_pose2img.py:196-213)."""
        images = super().__call__(ref_image, [pose_image], ref_pose_image, width, height, 1, num_inference_steps,
                                  guidance_scale, num_images_per_prompt, eta, generator, output_type, False, callback,
                                  callback_steps, context_frames=1, context_overlap=0, **kwargs)
        if not return_dict:
            return images
        return Pose2ImagePipelineOutput(images=images)
