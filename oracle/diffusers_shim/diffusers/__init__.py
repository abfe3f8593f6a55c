"""TEST INFRASTRUCTURE ONLY — a minimal restatement of the diffusers==0.24.0 names that /root/reference/src imports.

The reference (Zejun-Yang/AniPortrait) pins diffusers==0.24.0 (requirements.txt:5), which is not installed in the
authoring container and has no network to fetch. This shim restates, from the published 0.24.0 algorithms, exactly the
leaf primitives the reference's hot path uses (Attention + AttnProcessor2_0, FeedForward/GEGLU, Timesteps,
TimestepEmbedding, ResnetBlock2D, Down/Upsample2D, DDIMScheduler, AutoencoderKL, ModelMixin/ConfigMixin,
DiffusionPipeline, VaeImageProcessor, randn_tensor) so that the reference's own model/pipeline wiring can be imported
UNMODIFIED from /root/reference and run on CPU to generate golden vectors (oracle/make_golden.py).

It is never imported by the product (aniportrait_b200/) — only by oracle/ and tests/. Parity is therefore pinned to
"reference wiring + restated diffusers leaves"; see DESIGN.md §oracle ("parity unpinned" w.r.t. real diffusers).
"""
from .pipeline_utils import DiffusionPipeline  # noqa: F401
from .models.autoencoder_kl import AutoencoderKL  # noqa: F401
from .schedulers import DDIMScheduler  # noqa: F401

__version__ = "0.24.0+shim"
