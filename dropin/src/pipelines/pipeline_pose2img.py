"""Import-path shim (see dropin/src/models/unet_3d.py)."""
from aniportrait_b200.pipelines.pipeline_pose2img import Pose2ImagePipeline, Pose2ImagePipelineOutput  # noqa: F401
