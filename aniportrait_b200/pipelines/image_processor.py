"""Image pre-processing used by the pipeline (host side, once per video): the two code paths of diffusers-0.24
VaeImageProcessor.preprocess that the reference exercises (pipeline_pose2vid_long.py:424-452):
  PIL image      -> RGB -> resize(lanczos) -> /255 -> 2x-1
  numpy uint8 HWC -> NO /255 (diffusers' numpy branch does not rescale) -> 2x-1   (the pose maps; values in [-1, 509])
"""
import numpy as np
import PIL.Image
import torch


class VaeImageProcessor:
    def __init__(self, do_resize=True, vae_scale_factor=8, resample="lanczos", do_normalize=True,
                 do_convert_rgb=False, **_):
        self.do_resize, self.vae_scale_factor, self.do_normalize = do_resize, vae_scale_factor, do_normalize
        self.do_convert_rgb = do_convert_rgb

    def _hw(self, image, height, width):
        if height is None:
            height = image.height if isinstance(image, PIL.Image.Image) else image.shape[-2]
        if width is None:
            width = image.width if isinstance(image, PIL.Image.Image) else image.shape[-1]
        f = self.vae_scale_factor
        return height - height % f, width - width % f

    def preprocess(self, image, height=None, width=None) -> torch.Tensor:
        if isinstance(image, (PIL.Image.Image, np.ndarray, torch.Tensor)):
            image = [image]
        if isinstance(image[0], PIL.Image.Image):
            if self.do_convert_rgb:
                image = [i.convert("RGB") for i in image]
            if self.do_resize:
                height, width = self._hw(image[0], height, width)
                image = [i.resize((width, height), resample=PIL.Image.LANCZOS) for i in image]
            arr = np.stack([np.array(i).astype(np.float32) / 255.0 for i in image], axis=0)
            t = torch.from_numpy(arr.transpose(0, 3, 1, 2))
        elif isinstance(image[0], np.ndarray):
            arr = np.concatenate(image, axis=0) if image[0].ndim == 4 else np.stack(image, axis=0)
            t = torch.from_numpy(arr.transpose(0, 3, 1, 2))
            height, width = self._hw(t, height, width)
            if self.do_resize and tuple(t.shape[-2:]) != (height, width):
                t = torch.nn.functional.interpolate(t, size=(height, width))
        else:
            t = torch.cat(image, 0) if image[0].ndim == 4 else torch.stack(image, 0)
            height, width = self._hw(t, height, width)
            if self.do_resize and tuple(t.shape[-2:]) != (height, width):
                t = torch.nn.functional.interpolate(t, size=(height, width))
        do_normalize = self.do_normalize
        if t.min() < 0 and do_normalize:
            do_normalize = False
        if do_normalize:
            t = 2.0 * t - 1.0
        return t
