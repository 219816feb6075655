"""Pre-processing of an edit request on the GPU (device part) through pbe_b200.preprocess ->
pbe_normalize_u8 / pbe_prepare_inpaint_u8 / pbe_resize_bilinear (C ABI), against the oracle (oracle/preprocess_ref.py, pinned
to the live torchvision transforms).  Byte / fp32-op-sequence work: bit-exact; the antialiased resize within 5e-7 (the
oracle itself sits an ulp from ATen's FMA-contracted CPU kernel)."""
import pytest
import torch

from oracle import preprocess_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("B,H,W", [(1, 512, 512), (3, 64, 48), (2, 224, 224), (1, 7, 5)])
def test_get_tensor_and_get_tensor_clip_bit_exact(dev, B, H, W):
    from pbe_b200 import preprocess as P
    img, _ = R.synthetic_u8_request(B, max(H, 16), max(W, 16), seed=H + W)
    img = img[:, :H, :W].contiguous()
    assert torch.equal(P.get_tensor()(img.to(dev)).cpu(), R.normalize_u8(img))
    assert torch.equal(P.get_tensor_clip()(img.to(dev)).cpu(), R.normalize_u8(img, R.CLIP_MEAN, R.CLIP_STD))
    assert torch.equal(P.get_tensor(normalize=False)(img.to(dev)).cpu(), img.permute(0, 3, 1, 2).float().div(255))
    one = P.get_tensor()(img[0].to(dev))                      # a single HWC image, as the scripts pass it
    assert one.shape == (3, H, W) and torch.equal(one.cpu(), R.normalize_u8(img[:1])[0])


@pytest.mark.parametrize("binarize", [True, False])
@pytest.mark.parametrize("B,H,W", [(1, 512, 512), (2, 96, 80)])
def test_prepare_inpaint_bit_exact(dev, B, H, W, binarize):
    from pbe_b200 import preprocess as P
    img, mask = R.synthetic_u8_request(B, H, W, seed=B + H)
    image, m, inpaint = P.prepare_inpaint(img.to(dev), mask.to(dev), binarize=binarize)
    ri, rm, rp = R.prepare_inpaint(img, mask, binarize=binarize)
    assert torch.equal(image.cpu(), ri) and torch.equal(m.cpu(), rm) and torch.equal(inpaint.cpu(), rp)
    assert set(m.unique().tolist()) <= {0.0, 1.0} if binarize else ((m > 0) & (m < 1)).any()


@pytest.mark.parametrize("H,W,h,w", [(512, 512, 64, 64), (768, 768, 96, 96), (96, 80, 12, 10), (100, 60, 13, 7), (30, 30, 45, 50),
                                     (64, 64, 64, 64)])
@pytest.mark.parametrize("antialias", [False, True])
def test_resize_matches_oracle(dev, H, W, h, w, antialias):
    from pbe_b200 import preprocess as P
    g = torch.Generator().manual_seed(H + w)
    x = torch.rand(2, 1, H, W, generator=g)
    _, mask, _ = R.prepare_inpaint(*R.synthetic_u8_request(2, H, W, seed=3), binarize=True)
    for t in (x, mask):
        out = P.Resize([h, w], antialias=antialias)(t.to(dev)).cpu()
        ref = R.resize_bilinear(t, (h, w), antialias)
        assert out.shape == ref.shape
        if antialias:
            assert (out - ref).abs().max().item() <= 5e-7
        else:
            assert torch.equal(out, ref)          # same fp32 op sequence, no FMA on either side


def test_request_preprocessing_end_to_end(dev):
    """scripts/inference.py:305-332 on one request: image / mask bytes -> inpaint image at 512^2 and the 64^2 latent mask."""
    from pbe_b200 import preprocess as P
    img, mask = R.synthetic_u8_request(1, 512, 512, seed=321)
    _, m, inpaint = P.prepare_inpaint(img.to(dev), mask.to(dev))
    m64 = P.Resize([64, 64])(m)
    ri, rm, rp = R.prepare_inpaint(img, mask)
    assert torch.equal(inpaint.cpu(), rp)
    assert torch.equal(m64.cpu(), R.resize_bilinear(rm, (64, 64)))
    assert m64.shape == (1, 1, 64, 64) and 0.0 < m64.mean().item() < 1.0
