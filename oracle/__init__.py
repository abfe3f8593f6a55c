"""TEST INFRASTRUCTURE — CPU oracle for the AniPortrait denoising hot path. See oracle/functional.py.
Never imported by the product package aniportrait_b200/."""
