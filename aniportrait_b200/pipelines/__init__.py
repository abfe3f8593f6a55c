from .pipeline_pose2vid_long import Pose2VideoPipeline, Pose2VideoPipelineOutput  # noqa: F401
from .pipeline_pose2img import Pose2ImagePipeline, Pose2ImagePipelineOutput  # noqa: F401
from .scheduler import DDIMScheduler  # noqa: F401
