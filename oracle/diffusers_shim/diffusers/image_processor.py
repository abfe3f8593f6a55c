"""VaeImageProcessor restated from diffusers 0.24.0 image_processor.py (preprocess paths used by the reference:
PIL image -> resize(lanczos) -> /255 -> 2x-1 ; numpy uint8 HxWx3 -> NO /255 -> 2x-1, pipeline_pose2vid_long.py:424-452)."""
import numpy as np
import PIL.Image
import torch


class VaeImageProcessor:
    def __init__(self, do_resize=True, vae_scale_factor=8, resample="lanczos", do_normalize=True,
                 do_binarize=False, do_convert_rgb=False, do_convert_grayscale=False):
        self.config = type("Cfg", (), dict(do_resize=do_resize, vae_scale_factor=vae_scale_factor, resample=resample,
                                           do_normalize=do_normalize, do_convert_rgb=do_convert_rgb))()

    @staticmethod
    def numpy_to_pt(images):
        if images.ndim == 3:
            images = images[..., None]
        return torch.from_numpy(images.transpose(0, 3, 1, 2))

    @staticmethod
    def pil_to_numpy(images):
        if not isinstance(images, list):
            images = [images]
        images = [np.array(image).astype(np.float32) / 255.0 for image in images]
        return np.stack(images, axis=0)

    @staticmethod
    def normalize(images):
        return 2.0 * images - 1.0

    def get_default_height_width(self, image, height=None, width=None):
        if height is None:
            height = image.height if isinstance(image, PIL.Image.Image) else (
                image.shape[2] if isinstance(image, torch.Tensor) else image.shape[1])
        if width is None:
            width = image.width if isinstance(image, PIL.Image.Image) else (
                image.shape[3] if isinstance(image, torch.Tensor) else image.shape[2])
        f = self.config.vae_scale_factor
        return height - height % f, width - width % f

    def resize(self, image, height=None, width=None):
        if isinstance(image, PIL.Image.Image):
            return image.resize((width, height), resample=PIL.Image.LANCZOS)
        if isinstance(image, torch.Tensor):
            if tuple(image.shape[-2:]) == (height, width):
                return image
            return torch.nn.functional.interpolate(image, size=(height, width))
        raise TypeError(type(image))

    def preprocess(self, image, height=None, width=None):
        supported = (PIL.Image.Image, np.ndarray, torch.Tensor)
        if isinstance(image, supported):
            image = [image]
        if isinstance(image[0], PIL.Image.Image):
            if self.config.do_convert_rgb:
                image = [i.convert("RGB") for i in image]
            if self.config.do_resize:
                height, width = self.get_default_height_width(image[0], height, width)
                image = [self.resize(i, height, width) for i in image]
            image = self.numpy_to_pt(self.pil_to_numpy(image))
        elif isinstance(image[0], np.ndarray):
            image = np.concatenate(image, axis=0) if image[0].ndim == 4 else np.stack(image, axis=0)
            image = self.numpy_to_pt(image)
            height, width = self.get_default_height_width(image, height, width)
            if self.config.do_resize:
                image = self.resize(image, height, width)
        elif isinstance(image[0], torch.Tensor):
            image = torch.cat(image, axis=0) if image[0].ndim == 4 else torch.stack(image, axis=0)
            height, width = self.get_default_height_width(image, height, width)
            if self.config.do_resize:
                image = self.resize(image, height, width)
        do_normalize = self.config.do_normalize
        if image.min() < 0 and do_normalize:
            do_normalize = False
        if do_normalize:
            image = self.normalize(image)
        return image
