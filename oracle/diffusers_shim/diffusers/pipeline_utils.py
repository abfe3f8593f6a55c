"""DiffusionPipeline: only register_modules / device / to / progress_bar (what Pose2VideoPipeline uses)."""
import torch
from tqdm.auto import tqdm


class DiffusionPipeline:
    def __init__(self):
        self._modules_registered = []

    def register_modules(self, **kwargs):
        if not hasattr(self, "_modules_registered"):
            self._modules_registered = []
        for name, module in kwargs.items():
            setattr(self, name, module)
            self._modules_registered.append(name)

    @property
    def device(self):
        for name in self._modules_registered:
            m = getattr(self, name)
            if isinstance(m, torch.nn.Module):
                for p in m.parameters():
                    return p.device
        return torch.device("cpu")

    def to(self, *args, **kwargs):
        for name in self._modules_registered:
            m = getattr(self, name)
            if isinstance(m, torch.nn.Module):
                m.to(*args, **kwargs)
        return self

    def progress_bar(self, iterable=None, total=None):
        cfg = getattr(self, "_progress_bar_config", {"disable": True})
        if iterable is not None:
            return tqdm(iterable, **cfg)
        return tqdm(total=total, **cfg)

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs
