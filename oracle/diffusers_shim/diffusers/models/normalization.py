import torch.nn as nn


class AdaLayerNormSingle(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("AdaLayerNormSingle is not on the AniPortrait hot path")
