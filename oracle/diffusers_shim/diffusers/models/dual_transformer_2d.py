import torch.nn as nn


class DualTransformer2DModel(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("DualTransformer2DModel is not on the AniPortrait hot path")
