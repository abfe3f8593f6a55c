"""ReferenceAttentionControl (mirror of the reference's src/models/mutual_self_attention.py).

The reference monkey-patches `forward` of every (Temporal)BasicTransformerBlock. Here the blocks already implement both
behaviours natively (aniportrait_b200/models/blocks.py BasicTransformerBlock.run); this object only selects the blocks
(same `fusion_blocks` rule and the same stable sort by descending width) and manages their `bank` lists, so
`writer = ReferenceAttentionControl(reference_unet, mode="write", ...)`, `reader.update(writer)` and `.clear()` keep the
reference's semantics (pipeline_pose2vid_long.py:393-406,485,569-570)."""
from __future__ import annotations

import torch

from .blocks import BasicTransformerBlock, TemporalBasicTransformerBlock


def torch_dfs(model: torch.nn.Module):
    result = [model]
    for child in model.children():
        result += torch_dfs(child)
    return result


class ReferenceAttentionControl:
    def __init__(self, unet, mode="write", do_classifier_free_guidance=False,
                 attention_auto_machine_weight=float("inf"), gn_auto_machine_weight=1.0, style_fidelity=1.0,
                 reference_attn=True, reference_adain=False, fusion_blocks="midup", batch_size=1) -> None:
        assert mode in ["read", "write"]
        assert fusion_blocks in ["midup", "full"]
        if reference_adain:
            raise NotImplementedError("reference_adain is unused by AniPortrait")
        self.unet = unet
        self.mode = mode
        self.reference_attn = reference_attn
        self.fusion_blocks = fusion_blocks
        self.do_classifier_free_guidance = do_classifier_free_guidance
        if self.reference_attn:
            for i, m in enumerate(self._modules(self.unet)):
                m.bank = []
                m._ref_mode = mode
                m._ref_cfg = bool(do_classifier_free_guidance)
                m._bank_kv = None
                m._attn2_const = None

    def _modules(self, unet, types=(BasicTransformerBlock, TemporalBasicTransformerBlock)):
        if self.fusion_blocks == "midup":
            mods = torch_dfs(unet.mid_block) + torch_dfs(unet.up_blocks)
        else:
            mods = torch_dfs(unet)
        mods = [m for m in mods if isinstance(m, types)]
        return sorted(mods, key=lambda x: -x.norm1.normalized_shape[0])

    def update(self, writer, dtype=torch.float16):
        if not self.reference_attn:
            return
        readers = self._modules(self.unet, (TemporalBasicTransformerBlock,))
        writers = [m for m in writer._modules(writer.unet) if type(m) is BasicTransformerBlock]
        for r, w in zip(readers, writers):
            r.bank = [v.clone().to(dtype) for v in w.bank]
            r._bank_kv = None
            r._attn2_const = None

    def clear(self):
        if not self.reference_attn:
            return
        for m in self._modules(self.unet):
            m.bank.clear()
            m._bank_kv = None
            m._attn2_const = None
