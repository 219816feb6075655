"""Device-side pre-processing of an edit request: the host-side mirror of what the reference scripts do with PIL / numpy /
torchvision before the sampler runs (third "next" row of SURVEY.md 8f).

    reference                                                       here
    get_tensor() / get_tensor_clip()  scripts/inference.py:106-124  get_tensor() / get_tensor_clip()  (uint8 HWC tensors in)
    mask, inpaint_image               scripts/inference.py:311-318  prepare_inpaint(image_u8, mask_u8, binarize=True)
                                      test_bench_dataset.py:89-98   prepare_inpaint(..., binarize=False)
    Resize([h, w])(inpaint_mask)      scripts/inference.py:332      Resize([h, w])(mask)

Inputs are CUDA uint8 tensors in the layout `torch.from_numpy(np.array(pil_image))` gives ([H,W,3] or [B,H,W,3]; masks
[H,W] or [B,H,W]); every output is produced by one kernel launch through the C ABI (`pbe_normalize_u8`,
`pbe_prepare_inpaint_u8`, `pbe_resize_bilinear`) with the reference's fp32 rounding.  There is no CPU path: PIL decoding
and the exemplar's PIL resize to 224x224 stay on the host, as in the reference.
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import torch

from . import _lib

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _require_cuda_u8(t: torch.Tensor, what: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or t.dtype != torch.uint8:
        raise TypeError(f"{what}: expected a uint8 tensor (the bytes PIL / numpy hand over), got {getattr(t, 'dtype', type(t))}")
    if not t.is_cuda:
        raise RuntimeError(f"{what}: pbe_b200 has no CPU path; move the uint8 tensor to the GPU first")
    return t.contiguous()


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class _Normalize:
    """Callable returned by get_tensor() / get_tensor_clip(): ToTensor [+ Normalize] on uint8 HWC CUDA tensors."""

    def __init__(self, mean: Sequence[float], std: Sequence[float]):
        if any(float(s) == 0.0 for s in std):
            raise ValueError("std evaluated to zero, leading to division by zero.")   # torchvision's message
        self.mean = (ctypes.c_float * 3)(*[float(m) for m in mean])
        self.std = (ctypes.c_float * 3)(*[float(s) for s in std])

    def __call__(self, img_u8: torch.Tensor) -> torch.Tensor:
        x = _require_cuda_u8(img_u8, "get_tensor()")
        single = x.dim() == 3
        if single:
            x = x.unsqueeze(0)
        if x.dim() != 4 or x.shape[-1] != 3:
            raise ValueError(f"expected uint8 [H,W,3] or [B,H,W,3], got {tuple(img_u8.shape)}")
        B, H, W, _ = x.shape
        out = torch.empty((B, 3, H, W), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pbe_normalize_u8(x.data_ptr(), out.data_ptr(), B, H, W, self.mean, self.std, _stream(x.device)),
                       "pbe_normalize_u8")
        return out[0] if single else out


def get_tensor(normalize: bool = True, toTensor: bool = True):
    """scripts/inference.py:106-114: ToTensor + Normalize((0.5,)*3, (0.5,)*3); normalize=False is plain ToTensor (x / 255)."""
    if not toTensor:
        raise NotImplementedError("get_tensor(toTensor=False) is never used by the reference scripts")
    return _Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5)) if normalize else _Normalize((0.0, 0.0, 0.0), (1.0, 1.0, 1.0))


def get_tensor_clip(normalize: bool = True, toTensor: bool = True):
    """scripts/inference.py:116-124: ToTensor + Normalize(CLIP mean, CLIP std) for the 224x224 exemplar."""
    if not toTensor:
        raise NotImplementedError("get_tensor_clip(toTensor=False) is never used by the reference scripts")
    return _Normalize(CLIP_MEAN, CLIP_STD) if normalize else _Normalize((0.0, 0.0, 0.0), (1.0, 1.0, 1.0))


def prepare_inpaint(image_u8: torch.Tensor, mask_u8: torch.Tensor, binarize: bool = True
                    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(image_tensor [B,3,H,W], inpaint_mask [B,1,H,W], inpaint_image [B,3,H,W]) as scripts/inference.py:305-318 builds them
    (binarize=True: mask = 1 - m/255 snapped to {0, 1} at 0.5) or as COCOImageDataset.__getitem__ does
    (ldm/data/test_bench_dataset.py:73-100, binarize=False)."""
    img = _require_cuda_u8(image_u8, "prepare_inpaint(image)")
    msk = _require_cuda_u8(mask_u8, "prepare_inpaint(mask)")
    if img.dim() == 3:
        img = img.unsqueeze(0)
    if msk.dim() == 2:
        msk = msk.unsqueeze(0)
    if img.dim() != 4 or img.shape[-1] != 3 or msk.dim() != 3 or msk.shape != img.shape[:3]:
        raise ValueError(f"expected image uint8 [B,H,W,3] and mask uint8 [B,H,W], got {tuple(image_u8.shape)} / {tuple(mask_u8.shape)}")
    if msk.device != img.device:
        raise ValueError("image and mask must live on the same device")
    B, H, W, _ = img.shape
    image = torch.empty((B, 3, H, W), device=img.device, dtype=torch.float32)
    mask = torch.empty((B, 1, H, W), device=img.device, dtype=torch.float32)
    inpaint = torch.empty((B, 3, H, W), device=img.device, dtype=torch.float32)
    with torch.cuda.device(img.device):
        _lib.check(_lib.load().pbe_prepare_inpaint_u8(img.data_ptr(), msk.data_ptr(), B, H, W, 1 if binarize else 0,
                                                      image.data_ptr(), mask.data_ptr(), inpaint.data_ptr(), _stream(img.device)),
                   "pbe_prepare_inpaint_u8")
    return image, mask, inpaint


class Resize:
    """torchvision.transforms.Resize([h, w]) for float CUDA tensors [..., H, W] (bilinear, align_corners=False), as used on
    the mask at scripts/inference.py:332.  antialias=False is what the reference's pinned torchvision 0.12 does for tensors;
    antialias=True is the default of torchvision >= 0.17."""

    def __init__(self, size, antialias: bool = False):
        if isinstance(size, int) or len(size) != 2:
            raise NotImplementedError("Resize: only an explicit [h, w] is used by the reference scripts")
        self.size = (int(size[0]), int(size[1]))
        self.antialias = bool(antialias)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("Resize: pbe_b200 has no CPU path")
        if x.dim() < 2:
            raise ValueError("Resize: expected [..., H, W]")
        xx = x.detach().to(torch.float32).contiguous()
        H, W = xx.shape[-2:]
        nc = xx.numel() // (H * W)
        h, w = self.size
        out = torch.empty(tuple(xx.shape[:-2]) + (h, w), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().pbe_resize_bilinear(xx.data_ptr(), out.data_ptr(), nc, H, W, h, w, 1 if self.antialias else 0,
                                                       _stream(x.device)), "pbe_resize_bilinear")
        return out.to(x.dtype)
