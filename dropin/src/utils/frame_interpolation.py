"""Import-path shim (see dropin/src/models/unet_3d.py): the `-acc` FILM wrapper, batched over the frame pairs of a video."""
from aniportrait_b200.pipelines.frame_interpolation import (  # noqa: F401
    batch_images_interpolation_tool, init_frame_interpolation_model)
