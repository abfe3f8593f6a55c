"""Import-path shim (see dropin/src/models/unet_3d.py)."""
from aniportrait_b200.models.mutual_self_attention import ReferenceAttentionControl, torch_dfs  # noqa: F401
