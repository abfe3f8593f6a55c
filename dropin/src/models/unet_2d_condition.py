"""Import-path shim (see dropin/src/models/unet_3d.py)."""
from aniportrait_b200.models.unet_2d_condition import UNet2DConditionModel, UNet2DConditionOutput  # noqa: F401
