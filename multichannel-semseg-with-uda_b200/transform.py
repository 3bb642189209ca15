"""Batch / GPU counterparts of the reference's transform.py for the part that follows the PIL decode + `Scale`:

    get_img_transform(img_shape, normalize_way, use_crop)   transform.py:302-314   ToTensor + Normalize
    get_lbl_transform(img_shape, n_class, background_id)    transform.py:317-324   ToLabel + ReLabel
    unnormalize(np_input_img, normalize_way)                transform.py:285-294

The returned callables take DECODED uint8 batches on the GPU ([N,H,W,c] images / [N,H,W] label maps, already at
img_shape - the resize to img_shape is PIL code in the loader) and return exactly the tensors the reference's
per-image transforms produce, stacked: fp32 NCHW images, int64 labels (mcd_b200/pipeline.py, csrc/pipeline.cu).
"""
import torch

from mcd_b200 import pipeline


class _ImgTransform:
    def __init__(self, img_shape, normalize_way):
        self.img_shape, self.normalize_way = tuple(img_shape), normalize_way

    def __call__(self, img_u8):
        if img_u8.dim() == 3:                      # one decoded image [H,W,c] -> [c,H,W] like the reference's Compose
            return self(img_u8.unsqueeze(0))[0]
        assert (img_u8.shape[2], img_u8.shape[1]) == self.img_shape, \
            "images arrive at img_shape (width, height) = %s: `Scale` stays in the loader" % (self.img_shape,)
        return pipeline.transform_images([img_u8], self.normalize_way, out="nchw")


class _LblTransform:
    def __init__(self, img_shape, n_class, background_id):
        self.img_shape, self.n_class, self.background_id = tuple(img_shape), n_class, background_id

    def __call__(self, lbl_u8):
        assert (lbl_u8.shape[-1], lbl_u8.shape[-2]) == self.img_shape
        return pipeline.relabel(lbl_u8, self.n_class, self.background_id)


def get_img_transform(img_shape, normalize_way="imagenet", use_crop=False):
    if normalize_way == "imagenet":
        print("ImageNet Normalization!")
    elif normalize_way != "city":
        print("No normalization...")
    return _ImgTransform(img_shape, normalize_way)


def get_lbl_transform(img_shape, n_class, background_id=255, use_crop=False):
    return _LblTransform(img_shape, n_class, background_id)


def unnormalize(input_img, normalize_way="imagenet"):
    """fp32 [3,H,W] / [N,3,H,W] on the GPU -> uint8 HWC (the reference takes an HWC numpy array and returns a PIL image;
    `Image.fromarray(unnormalize(x).cpu().numpy())` is that image)."""
    if input_img.dim() == 3:
        return pipeline.unnormalize(input_img.unsqueeze(0), normalize_way)[0]
    return pipeline.unnormalize(input_img, normalize_way)
