"""aniportrait_b200 — B200-native (sm_100a) implementation of AniPortrait's denoising hot path.

Layout: csrc/ (CUDA kernels + C ABI), _lib.py/ops.py (ctypes binding), models/ and pipelines/ (host-side mirror of the
reference's src/models and src/pipelines call surface).
"""
__version__ = "0.1.0"
