"""Dev helper: run-to-run reproducibility of the kernels and of one UNet3D call (same inputs twice -> max abs diff)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from aniportrait_b200 import _lib, ops
from helpers import build_unet3d, build_unet2d, seeded_inputs_unet3d

dev = torch.device("cuda:0")
_lib.init(0)
torch.manual_seed(0)


def rep(name, fn, n=4):
    outs = [fn().float().clone() for _ in range(n)]
    d = max((o - outs[0]).abs().max().item() for o in outs[1:])
    print(f"{name:40s} max|run_i - run_0| = {d:.3e}   (|x|max {outs[0].abs().max().item():.2e})", flush=True)


for tokens, heads, d, frames, bank_from in [(256, 8, 40, 6, 3), (1024, 8, 40, 4, 2), (4096, 8, 40, 2, 1), (256, 8, 80, 4, 2),
                                            (64, 8, 160, 6, 3), (100, 2, 40, 3, 1)]:
    dpad = ops.head_pad(d)
    hp = heads * dpad
    qkv = torch.zeros(frames * tokens, 3, heads, dpad, device=dev, dtype=torch.float16)
    qkv[..., :d] = torch.randn(frames * tokens, 3, heads, d, device=dev) * 1.5
    qkv = qkv.reshape(frames * tokens, 3 * hp)
    bank = torch.zeros(tokens, 2, heads, dpad, device=dev, dtype=torch.float16)
    bank[..., :d] = torch.randn(tokens, 2, heads, d, device=dev) * 1.5
    bank = bank.reshape(tokens, 2 * hp)
    rep(f"attention N={tokens} d={d}", lambda: ops.attention(qkv[:, :hp], qkv[:, hp:2 * hp], qkv[:, 2 * hp:], frames, tokens,
        heads, d, dpad, bank_k=bank[:, :hp], bank_v=bank[:, hp:], bank_tokens=tokens, n_banks=1,
        first_bank_frame=bank_from, frames_per_bank=frames))

x = torch.randn(8, 16, 16, 320, device=dev, dtype=torch.float16)
wt = ops.pack_conv3x3_weight(torch.randn(320, 320, 3, 3, device=dev, dtype=torch.float16) * 0.02)
rep("conv3x3 320", lambda: ops.conv3x3(x, wt, 320))
a = torch.randn(2048, 320, device=dev, dtype=torch.float16)
w = torch.randn(2560, 320, device=dev, dtype=torch.float16) * 0.05
rep("gemm linear", lambda: ops.gemm(a, w))
g = torch.ones(320, device=dev); b = torch.zeros(320, device=dev)
rep("group_norm", lambda: ops.group_norm(x, g, b, 32, 1e-5, silu=True))
rep("layer_norm", lambda: ops.layer_norm(a, g, b))
qkv_t = torch.randn(2 * 16 * 64, 3 * 320, device=dev, dtype=torch.float16)
rep("temporal_attention", lambda: ops.temporal_attention(qkv_t, 2, 16, 64, 320, 8))

chans = (64, 128, 256, 256)
from aniportrait_b200.models import ReferenceAttentionControl
unet3d, _ = build_unet3d(chans, 201, dev)
unet2d, _ = build_unet2d(chans, 202, dev)
sample, ehs, ref_lat, pose = seeded_inputs_unet3d(2, 16, 16, 16, chans, 203)
h16 = lambda t: t.to(dev, torch.float16)
writer = ReferenceAttentionControl(unet2d, do_classifier_free_guidance=True, mode="write", fusion_blocks="full")
reader = ReferenceAttentionControl(unet3d, do_classifier_free_guidance=True, mode="read", fusion_blocks="full")
with torch.no_grad():
    def ref_pass():
        writer.clear()
        unet2d(h16(ref_lat).repeat(2, 1, 1, 1), torch.zeros((), device=dev), encoder_hidden_states=h16(ehs))
        return torch.cat([m.bank[0].reshape(-1) for m in writer._modules(unet2d)])
    rep("ReferenceNet banks", ref_pass)
    reader.update(writer)
    rep("UNet3D call (read mode, CFG)", lambda: unet3d(h16(sample), 500, encoder_hidden_states=h16(ehs),
                                                        pose_cond_fea=[h16(p) for p in pose]).sample)
    for v in ("AP_ATTENTION_V5",):
        pass
