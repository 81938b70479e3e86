"""Oracle (test infrastructure): the value range / layout step of the reference's input pipeline.

``dataloaders/custom_transforms.py:650-684`` ``Normalize_tf`` (uint8 image -> float32, ``/= 127.5``, ``-= 1.0``: two
separately rounded float32 operations) followed by ``:728-753`` ``ToTensor`` (H x W x C -> C x H x W float32; a 2-D image gets
a channel axis).  Everything before it (PIL crops, flips, elastic deformation, colour jitter, blur) works on uint8 PIL images on
the host and is out of scope; the device-side input kernel starts from those uint8 H x W x C images.  Pinned by
oracle/make_golden.py::case_input, which executes the two reference classes (lifted out of the file with ``ast``) unmodified."""
import numpy as np
import torch


def normalize_to_tensor(image_u8):
    """uint8 [H,W,C] or [H,W] -> float32 torch tensor [C,H,W] in [-1, 1]."""
    img = np.array(image_u8).astype(np.float32)          # Normalize_tf.__call__
    img /= 127.5
    img -= 1.0
    if img.ndim == 2:                                    # ToTensor.__call__
        img = np.expand_dims(img, 2)
    return torch.from_numpy(np.array(img).astype(np.float32).transpose((2, 0, 1)).copy()).float()


def batch(images_u8):
    """uint8 [B,H,W,C] -> float32 [B,C,H,W] (what the DataLoader's default collate stacks)."""
    return torch.stack([normalize_to_tensor(im) for im in images_u8])
