"""Pose2VideoPipeline — mirror of the reference's src/pipelines/pipeline_pose2vid_long.py (the pipeline every script
uses: scripts/pose2vid.py:166-176, audio2vid.py:230-240, vid2vid.py:213-223).

Same constructor and `__call__` signature and the same result object (`.videos`: fp32 CPU tensor [1, 3, F, H, W] in
[0, 1]). The denoising loop is re-designed for B200 while producing the reference's numbers:
  * latents live in one channels-last fp16 buffer [L, h, w, 4]; per window a gather kernel builds the CFG-duplicated UNet
    input, a scatter kernel accumulates the prediction into an fp32 [2, L, h, w, 4] buffer, and ONE kernel per step does
    overlap averaging + classifier-free guidance + the DDIM v-prediction update (reference :521-559);
  * PoseGuider does not depend on the timestep: it is evaluated once per window (the reference re-runs it on a
    CFG-duplicated batch every step, :531-536) — identical values, 25x less work;
  * ReferenceNet runs once; each reader block projects its bank to K/V once per video;
  * VAE decode is batched over frames (reference: one frame per call, :118-121).
With torch.distributed initialised and `dist_mode` set, frame windows are sharded across ranks: rank 0 runs the
ReferenceNet once and broadcasts the 16 banks over NCCL; "windows" mode additionally all-reduces the fp32 prediction
accumulator once per step. There is no collective inside any kernel's critical path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Union

import numpy as np
import torch

from .. import ops
from ..models.mutual_self_attention import ReferenceAttentionControl
from .sharding import plan_windows, windows_of_rank
from .image_processor import VaeImageProcessor


@dataclass
class Pose2VideoPipelineOutput:
    videos: Union[torch.Tensor, np.ndarray]


class Pose2VideoPipeline:
    _optional_components = []

    def __init__(self, vae, image_encoder, reference_unet, denoising_unet, pose_guider, scheduler,
                 image_proj_model=None, tokenizer=None, text_encoder=None):
        self.vae = vae
        self.image_encoder = image_encoder
        self.reference_unet = reference_unet
        self.denoising_unet = denoising_unet
        self.pose_guider = pose_guider
        self.scheduler = scheduler
        self.image_proj_model = image_proj_model
        self.tokenizer = tokenizer
        self.text_encoder = text_encoder
        self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1)
        from transformers import CLIPImageProcessor
        self.clip_image_processor = CLIPImageProcessor()
        self.ref_image_processor = VaeImageProcessor(vae_scale_factor=self.vae_scale_factor, do_convert_rgb=True)
        self.cond_image_processor = VaeImageProcessor(vae_scale_factor=self.vae_scale_factor, do_convert_rgb=True,
                                                      do_normalize=True)
        self.timings = {}
        self.use_cuda_graph = True   # capture the per-window UNet step once, replay it every DDIM step
        self._warmed = set()
        self._side_stream = None

    # -------------------------------------------------------------------------------------------- plumbing
    def _nn_modules(self):
        return [m for m in (self.vae, self.image_encoder, self.reference_unet, self.denoising_unet, self.pose_guider)
                if isinstance(m, torch.nn.Module)]

    def to(self, *args, **kwargs):
        for m in self._nn_modules():
            m.to(*args, **kwargs)
        return self

    @property
    def device(self):
        for m in self._nn_modules():
            for p in m.parameters():
                return p.device
        return torch.device("cpu")

    _execution_device = device

    def progress_bar(self, iterable=None, total=None):
        from tqdm.auto import tqdm
        cfg = getattr(self, "_progress_bar_config", {"disable": True})
        return tqdm(iterable, **cfg) if iterable is not None else tqdm(total=total, **cfg)

    def set_progress_bar_config(self, **kwargs):
        self._progress_bar_config = kwargs

    def enable_vae_slicing(self):
        pass

    def disable_vae_slicing(self):
        pass

    # -------------------------------------------------------------------------------------------- stages
    def prepare_latents(self, batch_size, num_channels_latents, width, height, video_length, dtype, device, generator,
                        latents=None):
        shape = (batch_size, num_channels_latents, video_length, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        if latents is None:
            rand_device = device
            if generator is not None and generator.device.type == "cpu" and torch.device(device).type != "cpu":
                rand_device = "cpu"   # diffusers randn_tensor: sample where the generator lives, then move
            latents = torch.randn(shape, generator=generator, device=rand_device, dtype=dtype).to(device)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    @torch.no_grad()
    def decode_latents_device(self, latents: torch.Tensor, frame_batch: int = 8):
        """latents [1, 4, F, h, w] -> device tensor [1, 3, F, H, W] in [0, 1] (reference :113-123, batched)."""
        video_length = latents.shape[2]
        z = (1 / 0.18215 * latents).permute(0, 2, 1, 3, 4).reshape(-1, *latents.shape[1:2], *latents.shape[3:])
        frames = []
        for i in range(0, z.shape[0], frame_batch):
            frames.append(self.vae.decode(z[i:i + frame_batch].to(self.vae.dtype)).sample)
        video = torch.cat(frames)
        video = video.view(-1, video_length, *video.shape[1:]).permute(0, 2, 1, 3, 4)
        return (video / 2 + 0.5).clamp(0, 1)

    def decode_latents(self, latents: torch.Tensor):
        """Reference-compatible: numpy fp32 on the host (reference :113-126)."""
        return self.decode_latents_device(latents).cpu().float().numpy()

    def _pose_maps_to_tensor(self, pose_images, height, width, device):
        """cond_image_processor.preprocess for the pose maps. uint8 HxWx3 arrays of the target size (what the scripts
        pass, pose2vid.py:153-158) take a fast path: the bytes go to the GPU and the reference's `2*x - 1` (no /255, see
        image_processor.py) is evaluated there in fp32 — identical values, 4x fewer bytes over PCIe, no host float pass."""
        frames = list(pose_images)
        if all(isinstance(p, np.ndarray) and p.dtype == np.uint8 and p.ndim == 3 and p.shape[:2] == (height, width)
               for p in frames):
            u8 = torch.from_numpy(np.ascontiguousarray(np.stack(frames, 0))).to(device, non_blocking=True)
            return u8.permute(0, 3, 1, 2).to(torch.float32) * 2.0 - 1.0
        return torch.cat([self.cond_image_processor.preprocess(p, height=height, width=width) for p in frames], dim=0)

    def _broadcast_banks(self, writer, device):
        """NCCL broadcast (rank 0 -> all) of the ReferenceNet banks: 16 tensors [dup, N, C] fp16, ~46 MB at 512x512."""
        rank = torch.distributed.get_rank()
        for m in writer._modules(writer.unet):
            if rank == 0:
                t = m.bank[0].contiguous()
                shape = torch.tensor(list(t.shape), device=device)
            else:
                shape = torch.zeros(3, dtype=torch.long, device=device)
            torch.distributed.broadcast(shape, 0)
            if rank != 0:
                t = torch.empty(*[int(v) for v in shape], dtype=torch.float16, device=device)
            torch.distributed.broadcast(t, 0)
            m.bank = [t]

    def _alpha_pair(self, t: int):
        if hasattr(self.scheduler, "alpha_pair"):
            return self.scheduler.alpha_pair(t)
        sch = self.scheduler   # a diffusers DDIMScheduler
        prev = t - sch.config.num_train_timesteps // sch.num_inference_steps
        a_t = float(sch.alphas_cumprod[t])
        a_p = float(sch.alphas_cumprod[prev]) if prev >= 0 else float(sch.final_alpha_cumprod)
        return a_t, a_p

    # -------------------------------------------------------------------------------------------- device core
    @torch.no_grad()
    def run_device(self, clip_pixels, ref_image_tensor, pose_cond, latents, num_inference_steps, guidance_scale,
                   context_schedule="uniform", context_frames=16, context_stride=1, context_overlap=4, callback=None,
                   callback_steps=1, clip_image_embeds=None, dist_mode=None, decode=True):
        """The hot path on device-resident inputs.
        clip_pixels [1,3,224,224] (or clip_image_embeds [1,768]); ref_image_tensor [1,3,H,W] in [-1,1];
        pose_cond [L,3,H,W] (pose maps as the reference's cond_image_processor emits them); latents [1,4,L,h,w].
        Returns the decoded video on the device, fp16 [1,3,L,H,W] in [0,1] (None if decode=False)."""
        device = self.device
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        cfg = guidance_scale > 1.0
        dup = 2 if cfg else 1
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        timesteps = [int(t) for t in self.scheduler.timesteps]
        rank, world = 0, 1
        if dist_mode is not None and torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()

        # CLIP image embedding (once per video; library module) --------------------------------------------
        if clip_image_embeds is None:
            clip_image_embeds = self.image_encoder(clip_pixels.to(device, dtype=self.image_encoder.dtype)).image_embeds
        ehs = clip_image_embeds.to(device).unsqueeze(1)
        if cfg:
            ehs = torch.cat([torch.zeros_like(ehs), ehs], dim=0)
        ehs = ehs.to(torch.float16).contiguous()

        writer = ReferenceAttentionControl(self.reference_unet, do_classifier_free_guidance=cfg, mode="write",
                                           batch_size=1, fusion_blocks="full")
        reader = ReferenceAttentionControl(self.denoising_unet, do_classifier_free_guidance=cfg, mode="read",
                                           batch_size=1, fusion_blocks="full")
        if latents.shape[0] != 1:
            raise NotImplementedError("one video per call (the reference fixes batch_size = 1)")
        L, h, w = latents.shape[2], latents.shape[3], latents.shape[4]
        lat = latents[0].permute(1, 2, 3, 0).to(torch.float16).contiguous()            # [L, h, w, 4]

        # reference image -> VAE latent (once) -> ReferenceNet write pass (once) -------------------------
        if world == 1 or rank == 0:
            ref_image_latents = self.vae.encode(ref_image_tensor.to(dtype=self.vae.dtype, device=device)) \
                .latent_dist.mean * 0.18215
            self.reference_unet(ref_image_latents.repeat(dup, 1, 1, 1), torch.zeros((), device=device),
                                encoder_hidden_states=ehs, return_dict=False)
        if world > 1:
            self._broadcast_banks(writer, device)
        reader.update(writer)

        # pose maps -> PoseGuider, once per window ------------------------------------------------------------
        pose_cond = pose_cond.to(device=device, dtype=self.pose_guider.dtype)
        windows, inv_count = plan_windows(L, num_inference_steps, context_schedule, context_frames, context_stride,
                                          context_overlap)
        shard = world > 1 and dist_mode == "windows"
        my_windows = windows_of_rank(windows, rank, world, shard)
        win_idx = [torch.tensor(wd, dtype=torch.int32, device=device) for wd in my_windows]
        win_pose = []
        for wd in my_windows:
            fea = self.pose_guider.forward_nhwc(pose_cond[wd])
            win_pose.append([f.to(torch.float16).contiguous() for f in fea])
        inv_count = inv_count.to(device=device, dtype=torch.float32)
        acc = torch.zeros(dup, L, h, w, 4, dtype=torch.float32, device=device)

        # denoising loop -----------------------------------------------------------------------------------------
        # Shapes are static, so the per-window work (gather -> ~800 kernel launches of the UNet -> scatter-accumulate)
        # is captured once in a CUDA graph per window and replayed every step; only the timestep buffer changes.
        t_dev = torch.zeros(1, dtype=torch.float32, device=device)

        def window_step(idx, pose_fea):
            x = ops.gather_window(lat, idx, dup, 64)
            pred = self.denoising_unet.forward_nhwc(x, dup, idx.numel(), t_dev, ehs, pose_fea)
            ops.scatter_accumulate(pred, idx, acc)

        graphs = []
        self.denoising_unet.prepare_reference(dup, len(my_windows[0]) if my_windows else context_frames, ehs)
        if self.use_cuda_graph and len(win_idx) > 0:
            t_dev.fill_(float(timesteps[0]))
            side = self._side_stream if getattr(self, "_side_stream", None) is not None else torch.cuda.Stream(device=device)
            self._side_stream = side
            side.wait_stream(torch.cuda.current_stream(device))
            warm_key = (L, h, w, dup, tuple(len(wd) for wd in my_windows))
            if warm_key not in self._warmed:
                with torch.cuda.stream(side):
                    # first use of these shapes: one eager pass sets kernel attributes / workspaces / weight packing
                    window_step(win_idx[0], win_pose[0])
                    acc.zero_()
                self._warmed.add(warm_key)
            torch.cuda.current_stream(device).wait_stream(side)
            pool = None
            for idx, pose_fea in zip(win_idx, win_pose):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool, stream=side):
                    window_step(idx, pose_fea)
                pool = g.pool()
                graphs.append(g)
            acc.zero_()
        ev[1].record()
        with self.progress_bar(total=num_inference_steps) as progress_bar:
            for i, t in enumerate(timesteps):
                t_dev.fill_(float(t))
                if graphs:
                    for g in graphs:
                        g.replay()
                else:
                    for idx, pose_fea in zip(win_idx, win_pose):
                        window_step(idx, pose_fea)
                if shard:
                    torch.distributed.all_reduce(acc)
                a_t, a_p = self._alpha_pair(t)
                ops.cfg_ddim_step(acc, inv_count, float(guidance_scale), a_t, a_p, lat)
                progress_bar.update()
                if callback is not None and i % callback_steps == 0:
                    callback(i, t, lat.permute(3, 0, 1, 2).unsqueeze(0))
        ev[2].record()
        reader.clear()
        writer.clear()

        # decode ---------------------------------------------------------------------------------------------
        latents_out = lat.permute(3, 0, 1, 2).unsqueeze(0)                                # [1, 4, L, h, w]
        self.last_latents = latents_out
        video = None
        if decode:
            if shard and L % world == 0:   # frames sharded over ranks, gathered on every rank
                part = self.decode_latents_device(latents_out[:, :, rank::world]).contiguous()
                parts = [torch.empty_like(part) for _ in range(world)]
                torch.distributed.all_gather(parts, part)
                video = torch.empty(1, 3, L, part.shape[-2], part.shape[-1], device=device, dtype=part.dtype)
                for r in range(world):
                    video[:, :, r::world] = parts[r]
            else:
                video = self.decode_latents_device(latents_out)
        ev[3].record()
        self._events = ev
        self._meta = dict(windows=len(windows), steps=len(timesteps))
        return video

    def collect_timings(self):
        """Call after a synchronize: per-phase device times of the last run_device()."""
        ev = self._events
        self.timings = dict(reference_ms=ev[0].elapsed_time(ev[1]), denoise_ms=ev[1].elapsed_time(ev[2]),
                            decode_ms=ev[2].elapsed_time(ev[3]), **self._meta)
        return self.timings

    # -------------------------------------------------------------------------------------------- __call__
    @torch.no_grad()
    def __call__(self, ref_image, pose_images, ref_pose_image, width, height, video_length, num_inference_steps,
                 guidance_scale, num_images_per_prompt=1, eta: float = 0.0,
                 generator: Optional[Union[torch.Generator, List[torch.Generator]]] = None,
                 output_type: Optional[str] = "tensor", return_dict: bool = True,
                 callback: Optional[Callable[[int, int, torch.Tensor], None]] = None,
                 callback_steps: Optional[int] = 1, context_schedule="uniform", context_frames=16, context_stride=1,
                 context_overlap=4, context_batch_size=1, interpolation_factor=1, clip_image_embeds=None,
                 latents=None, dist_mode=None, **kwargs):
        """Reference signature (pipeline_pose2vid_long.py:338-363). Host-side preprocessing, then run_device(), then the
        fp32 host copy of the video.
        dist_mode (only with torch.distributed initialised):
             None       every rank computes the whole video redundantly (reference behaviour)
             "windows"  the windows of ONE long video are sharded over ranks; fp32 prediction accumulator all-reduced
                        (NCCL) once per step; every rank ends with the full latents and video
             "clips"    every rank denoises its OWN clip (its own pose_images / latents) of the same reference portrait
           In both distributed modes rank 0 alone runs the ReferenceNet and broadcasts the 16 banks (NCCL)."""
        if eta != 0.0:
            raise NotImplementedError("eta > 0 is unused by AniPortrait")
        if context_batch_size != 1:
            raise NotImplementedError("context_batch_size > 1 cannot work in the reference either (bank batch mismatch)")
        if interpolation_factor not in (0, 1):
            raise NotImplementedError("latent interpolation is disabled in the reference (interpolation_factor=1)")
        device = self.device
        if device.type != "cuda":
            raise RuntimeError("aniportrait_b200.Pose2VideoPipeline runs on CUDA (sm_100a) only: no CPU fallback")
        clip_pixels = None
        if clip_image_embeds is None:
            clip_pixels = self.clip_image_processor.preprocess(ref_image.resize((224, 224)),
                                                               return_tensors="pt").pixel_values
        embed_dtype = self.image_encoder.dtype if isinstance(self.image_encoder, torch.nn.Module) else torch.float16
        latents = self.prepare_latents(num_images_per_prompt, self.denoising_unet.in_channels, width, height,
                                       video_length, embed_dtype, device, generator, latents)
        ref_image_tensor = self.ref_image_processor.preprocess(ref_image, height=height, width=width)
        pose_cond = self._pose_maps_to_tensor(pose_images, height, width, device)            # [L, 3, H, W]
        video = self.run_device(clip_pixels, ref_image_tensor, pose_cond, latents, num_inference_steps,
                                guidance_scale, context_schedule, context_frames, context_stride, context_overlap,
                                callback, callback_steps, clip_image_embeds, dist_mode)
        images = video.cpu().float().numpy()       # "we always cast to float32" (reference :124-125)
        self.collect_timings()
        if output_type == "tensor":
            images = torch.from_numpy(images)
        if not return_dict:
            return images
        return Pose2VideoPipelineOutput(videos=images)
