"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's host-side pre-processing of an edit
request, in plain numpy / torch fp32.  Only tests/, __graft_entry__.smoke() and bench.py may import this.

Follows
  * get_tensor() / get_tensor_clip(): torchvision ToTensor + Normalize          scripts/inference.py:106-124,
                                                                                 ldm/data/test_bench_dataset.py:37-61
  * mask = 1 - m/255, 0 / 1 at the 0.5 threshold; inpaint = image * mask        scripts/inference.py:311-318
  * mask = 1 - ToTensor(m) (no threshold); inpaint = image * mask               ldm/data/test_bench_dataset.py:89-98
  * Resize([h, w]) of the float mask                                            scripts/inference.py:332
    = torch.nn.functional.interpolate(mode="bilinear", align_corners=False[, antialias=True]); the algorithm lives in
    ATen (torch 2.11 here; the reference pins torch 1.11 / torchvision 0.12, whose Resize does NOT antialias tensors):
    upsample_bilinear2d and _upsample_bilinear2d_aa are restated below from their published implementation
    (aten/src/ATen/native/UpSample.h area_pixel_compute_source_index, cpu/UpSampleKernel.cpp
    _compute_indices_min_size_weights_aa) and pinned against the live functions in tests/test_oracle_pinned.py.
Pinned: yes (live torchvision transforms + torch.nn.functional.interpolate in the authoring container)."""
import numpy as np
import torch

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
HALF = (0.5, 0.5, 0.5)


def normalize_u8(img_u8: torch.Tensor, mean=HALF, std=HALF) -> torch.Tensor:
    """ToTensor + Normalize on uint8 [B,H,W,3] -> fp32 [B,3,H,W] (torchvision functional: div(255), sub_(mean).div_(std))."""
    x = img_u8.permute(0, 3, 1, 2).to(torch.float32).div(255)
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return x.sub(m).div(s)


def prepare_inpaint(img_u8: torch.Tensor, mask_u8: torch.Tensor, binarize: bool = True):
    """(image [B,3,H,W], mask [B,1,H,W], inpaint_image [B,3,H,W]) from uint8 image [B,H,W,3] and uint8 mask [B,H,W]."""
    image = normalize_u8(img_u8)
    m = mask_u8.numpy()[:, None]
    mask = 1 - m.astype(np.float32) / 255.0            # scripts/inference.py:313
    if binarize:
        mask[mask < 0.5] = 0                           # :314-315
        mask[mask >= 0.5] = 1
    mask_t = torch.from_numpy(mask.astype(np.float32))
    return image, mask_t, image * mask_t


def resize_bilinear(x: torch.Tensor, size, antialias: bool = False) -> torch.Tensor:
    """[N,C,H,W] fp32 -> [N,C,h,w], restating ATen (see module docstring)."""
    N, C, H, W = x.shape
    h, w = size
    src = x.numpy().astype(np.float32)
    f32 = np.float32
    if not antialias:
        def axis(n_in, n_out):
            scale = f32(n_in) / f32(n_out)
            d = np.arange(n_out, dtype=np.float32)
            s = scale * (d + f32(0.5)) - f32(0.5)
            s = np.maximum(s, f32(0)).astype(np.float32)
            i0 = s.astype(np.int64)
            i1 = i0 + (i0 < n_in - 1)
            l1 = (s - i0.astype(np.float32)).astype(np.float32)
            return i0, i1, (f32(1) - l1).astype(np.float32), l1
        y0, y1, ly0, ly1 = axis(H, h)
        x0, x1, lx0, lx1 = axis(W, w)
        top = lx0 * src[:, :, y0][:, :, :, x0] + lx1 * src[:, :, y0][:, :, :, x1]
        bot = lx0 * src[:, :, y1][:, :, :, x0] + lx1 * src[:, :, y1][:, :, :, x1]
        out = ly0[:, None] * top + ly1[:, None] * bot
        return torch.from_numpy(out.astype(np.float32))

    def windows(n_in, n_out):
        scale = f32(n_in) / f32(n_out)
        support = scale if scale >= 1 else f32(1)
        invscale = f32(1.0 / float(scale)) if scale >= 1 else f32(1)
        res = []
        for i in range(n_out):
            center = f32(float(scale) * (i + 0.5))
            lo = max(int(float(f32(center - support)) + 0.5), 0)
            n = min(int(float(f32(center + support)) + 0.5), n_in) - lo
            ws = []
            total = f32(0)
            for j in range(n):
                t = abs(f32((float(f32(f32(j + lo) - center)) + 0.5) * float(invscale)))
                wgt = f32(1) - t if t < 1 else f32(0)
                ws.append(f32(wgt))
                total = f32(total + wgt)
            if total != 0:
                ws = [f32(v / total) for v in ws]
            res.append((lo, ws))
        return res
    wx, wy = windows(W, w), windows(H, h)
    tmp = np.zeros((N, C, H, w), dtype=np.float32)       # width pass first, then height (separable, sequential sums)
    for ox, (lo, ws) in enumerate(wx):
        t = src[:, :, :, lo] * ws[0]
        for j in range(1, len(ws)):
            t = (t + src[:, :, :, lo + j] * ws[j]).astype(np.float32)
        tmp[:, :, :, ox] = t
    out = np.zeros((N, C, h, w), dtype=np.float32)
    for oy, (lo, ws) in enumerate(wy):
        t = tmp[:, :, lo, :] * ws[0]
        for j in range(1, len(ws)):
            t = (t + tmp[:, :, lo + j, :] * ws[j]).astype(np.float32)
        out[:, :, oy, :] = t
    return torch.from_numpy(out)


def synthetic_u8_request(B, H, W, seed=0):
    """uint8 image [B,H,W,3] and a COCOEE-style bbox mask [B,H,W] (255 inside the box to edit, soft 1-pixel border values so
    the 0.5 threshold matters), as PIL would hand them over."""
    g = torch.Generator().manual_seed(seed)
    img = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    mask = torch.zeros(B, H, W, dtype=torch.uint8)
    for b in range(B):
        y0, x0 = int(torch.randint(0, H // 2, (1,), generator=g)), int(torch.randint(0, W // 2, (1,), generator=g))
        hh, ww = int(torch.randint(H // 8, H // 2, (1,), generator=g)), int(torch.randint(W // 8, W // 2, (1,), generator=g))
        mask[b, y0:y0 + hh, x0:x0 + ww] = 255
        mask[b, y0, x0:x0 + ww] = torch.randint(0, 256, (ww,), generator=g, dtype=torch.uint8)   # anti-aliased edge
        mask[b, y0:y0 + hh, x0] = 127
        mask[b, y0 + hh - 1, x0:x0 + ww] = 128
    return img, mask
