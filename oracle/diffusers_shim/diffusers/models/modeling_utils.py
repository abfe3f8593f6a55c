import itertools

import torch


class ModelMixin(torch.nn.Module):
    """diffusers.models.modeling_utils.ModelMixin — only what the reference touches: dtype/device, config fallthrough."""
    _supports_gradient_checkpointing = False

    def __init__(self):
        super().__init__()

    def __getattr__(self, name):
        is_in_config = "_internal_dict" in self.__dict__ and name in self.__dict__["_internal_dict"]
        if is_in_config and name not in self.__dict__:
            return self.__dict__["_internal_dict"][name]
        return super().__getattr__(name)

    @property
    def device(self):
        for t in itertools.chain(self.parameters(), self.buffers()):
            return t.device
        return torch.device("cpu")

    @property
    def dtype(self):
        for t in itertools.chain(self.parameters(), self.buffers()):
            if t.is_floating_point():
                return t.dtype
        return torch.float32
