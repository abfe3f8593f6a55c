"""Audio front-end pieces of SURVEY.md 8f (N3). Only the autoregressive head-pose decoder loop is here."""
from .pose_infer import enable_kv_cache, kv_cached_infer  # noqa: F401
