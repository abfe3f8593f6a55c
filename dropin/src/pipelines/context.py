"""Import-path shim (see dropin/src/models/unet_3d.py)."""
from aniportrait_b200.pipelines.context import get_context_scheduler, get_total_steps, ordered_halving, uniform  # noqa: F401
