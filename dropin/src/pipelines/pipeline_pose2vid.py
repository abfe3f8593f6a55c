"""Import-path shim (see dropin/src/models/unet_3d.py)."""
from aniportrait_b200.pipelines.pipeline_pose2vid import Pose2VideoPipeline, Pose2VideoPipelineOutput  # noqa: F401
