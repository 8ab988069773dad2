"""CPU oracle for frame preprocessing -- TEST INFRASTRUCTURE ONLY (see oracle/git_oracle.py header).

Restates ``image_transform()`` of the reference (/root/reference/src/utils/dataloader.py:18-32, duplicated at
src/real_time_inference.py:16-28): ToTensor -> Resize(224, BICUBIC) -> CenterCrop(224) -> BGR->RGB -> CLIP Normalize,
applied to an OpenCV frame (uint8, H x W x 3, BGR).  With the reference's pinned torchvision 0.16 the resize of a
tensor is ``torch.nn.functional.interpolate(mode='bicubic', align_corners=False)`` WITHOUT antialiasing (the
``antialias`` default only became True in 0.17), without clamping, on the [0,1]-scaled float image; the smaller edge
goes to 224 and the other to int(224 * long / short); the crop offset is round((size - 224) / 2).

PARITY STATUS: pinned against torchvision's own transforms in this image (tests/test_cpu_oracle_and_host.py runs
torchvision.transforms.v2 functional resize with antialias=False + center_crop + normalize on the same frames).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # dataloader.py:27
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)   # dataloader.py:28


def resized_hw(h: int, w: int, size: int = 224):
    """torchvision Resize(int): smaller edge -> size, the other edge -> int(size * long / short)."""
    if h <= w:
        return size, int(size * w / h)
    return int(size * h / w), size


def preprocess_frames(frames_u8: torch.Tensor, size: int = 224) -> torch.Tensor:
    """frames uint8 [N, H, W, 3] (BGR) -> float32 [N, 3, size, size] (RGB, CLIP-normalised)."""
    assert frames_u8.dtype == torch.uint8 and frames_u8.dim() == 4 and frames_u8.shape[-1] == 3
    x = frames_u8.permute(0, 3, 1, 2).float() / 255.0                       # ToTensor
    h, w = x.shape[-2:]
    nh, nw = resized_hw(h, w, size)
    x = F.interpolate(x, size=(nh, nw), mode="bicubic", align_corners=False)  # Resize(224, BICUBIC), no antialias
    top, left = int(round((nh - size) / 2.0)), int(round((nw - size) / 2.0))  # CenterCrop(224)
    x = x[:, :, top:top + size, left:left + size]
    x = x[:, [2, 1, 0]]                                                       # BGR -> RGB
    mean = torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    return (x - mean) / std
