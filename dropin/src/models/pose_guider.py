"""Import-path shim (see dropin/src/models/unet_3d.py)."""
from aniportrait_b200.models.pose_guider import PoseGuider  # noqa: F401
