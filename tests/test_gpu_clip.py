"""Conditioning front-end on the GPU, through pbe_b200.FrozenCLIPImageEmbedder -> pbe_clip_encode
(C ABI), against goldens produced by the live transformers tower + the reference mapper, and the fp32 oracle.  Tolerance:
relative L2 <= 1e-2 on the conditioning token (bf16 operands / fp32 accumulation in the tower, fp32 mapper)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _golden(golden_dir, name):
    idx = json.load(open(os.path.join(golden_dir, "golden_index.json")))
    a = np.load(os.path.join(golden_dir, name + ".npy"))
    assert hashlib.sha256(a.astype(np.float32).tobytes()).hexdigest() == idx[name]["sha256"]
    return torch.from_numpy(a), idx[name]


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def _make(cfg, sd, dev):
    from pbe_b200.clip import FrozenCLIPImageEmbedder
    m = FrozenCLIPImageEmbedder(**cfg)
    m.load_state_dict(sd, strict=True)
    return m.to(dev).eval()


@pytest.mark.parametrize("tag", ["small", "v1"])
def test_clip_embed_vs_reference_golden(dev, golden_dir, tag):
    from oracle import clip_ref as K
    g, meta = _golden(golden_dir, f"{tag}_clip_embed")
    cfg = K.SMALL_CLIP_CFG if tag == "small" else K.V1_CLIP_CFG
    m = _make(cfg, K.make_state_dict(cfg, meta["weight_seed"]), dev)
    x = K.synthetic_exemplars(meta["B"], cfg["image_size"], seed=meta["image_seed"])
    z = m(x.to(dev)).cpu()
    assert z.shape == g.shape and torch.isfinite(z).all()
    print(f"{tag} CLIP front-end: rel-L2 = {_rel(z, g):.3e}")
    assert _rel(z, g) <= 1e-2


def test_clip_embed_batch_vs_oracle_on_gpu(dev):
    """ViT-L/14 geometry at a batch that is not a multiple of anything (257 tokens x 5 images = 1285 rows)."""
    from oracle import clip_ref as K
    cfg = K.V1_CLIP_CFG
    sd = K.make_state_dict(cfg, 321)
    m = _make(cfg, sd, dev)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    x = K.synthetic_exemplars(5, 224, seed=77).to(dev)
    z = m.encode(x)
    with torch.no_grad():
        ref = K.encode(sd_dev, cfg, x)
    print(f"v1 CLIP front-end B=5: rel-L2 = {_rel(z, ref):.3e}")
    assert z.shape == (5, 1, 1024)
    assert _rel(z, ref) <= 1e-2
