"""Import-path shim: lets the reference's scripts (`from src.models.unet_3d import UNet3DConditionModel`) pick up the
B200-native implementation unchanged. Put `dropin/` before the reference checkout on PYTHONPATH."""
from aniportrait_b200.models.unet_3d import UNet3DConditionModel, UNet3DConditionOutput  # noqa: F401
