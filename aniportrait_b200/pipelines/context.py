"""Sliding-window ("context") scheduler with the semantics of the reference's src/pipelines/context.py:7-76 (AnimateDiff's
`uniform` policy): same window lists for the same arguments (tests/test_host_cpu.py compares them with the oracle and, when
the reference tree is present, with the reference itself). Host-side integer logic only."""
import math
from typing import Callable, Iterator, List, Optional


def ordered_halving(val: int) -> float:
    """Van der Corput-style fraction: the 64-bit binary expansion of `val`, reversed, read as a number in [0, 1)."""
    rev = 0
    for _ in range(64):
        rev = (rev << 1) | (val & 1)
        val >>= 1
    return rev / float(1 << 64)


def uniform(step: int = 0, num_steps: Optional[int] = None, num_frames: int = 0, context_size: Optional[int] = None,
            context_stride: int = 3, context_overlap: int = 4, closed_loop: bool = True) -> Iterator[List[int]]:
    """Windows of `context_size` frame indices. A clip that fits one window is returned whole; otherwise, for every
    dilation level l < min(context_stride, ceil(log2(L / size)) + 1), windows of frames spaced 2^l apart start every
    (size * 2^l - overlap) frames, shifted by a step-dependent offset and wrapped around the clip (closed loop)."""
    if num_frames <= context_size:
        yield list(range(num_frames))
        return
    levels = min(context_stride, int(math.ceil(math.log2(num_frames / context_size))) + 1)
    frac = ordered_halving(step)
    shift = int(round(num_frames * frac))
    end = num_frames + shift - (0 if closed_loop else context_overlap)
    for level in range(levels):
        hop = 1 << level                                   # frame spacing inside a window at this level
        if context_size * hop - context_overlap == 0:
            raise ValueError("range() arg 3 must not be zero")     # what the reference's range(...) raises (context.py:36-40)
        if context_size * hop - context_overlap < 0:
            continue                                       # negative range step: the reference yields nothing at this level
        origin = int(frac * hop) + shift
        while origin < end:
            yield [(origin + k * hop) % num_frames for k in range(context_size)]
            origin += context_size * hop - context_overlap


def get_context_scheduler(name: str) -> Callable:
    if name != "uniform":
        raise ValueError(f"Unknown context_overlap policy {name}")
    return uniform


def get_total_steps(scheduler, timesteps, num_steps=None, num_frames=0, context_size=None, context_stride=3,
                    context_overlap=4, closed_loop=True) -> int:
    """Number of UNet calls of a whole run (windows per step, summed over the steps)."""
    total = 0
    for i in range(len(timesteps)):
        total += sum(1 for _ in scheduler(i, num_steps, num_frames, context_size, context_stride, context_overlap))
    return total
