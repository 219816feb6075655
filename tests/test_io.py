"""Host-side I/O pipeline (pbe_b200/io.py): the reference's test-bench file layout, the prefetching loader and the threaded
PNG writer.  CPU only -- the module is plumbing around the GPU path and must be testable without one."""
import os
import threading
import time

import numpy as np
import pytest
import torch

from pbe_b200 import io as IO


@pytest.fixture(scope="module")
def bench_dir(tmp_path_factory):
    root = str(tmp_path_factory.mktemp("test_bench"))
    ids = IO.write_synthetic_test_bench(root, 7, size=64, ref_size=48, seed=3)
    return root, ids


def test_layout_matches_the_reference_dataset(bench_dir):
    """test_bench_dataset.py:74-91: <dir>/GT_3500/<id:012d>_GT.png, Ref_3500/..._ref.png, Mask_bbox_3500/..._mask.png, id_list.npy."""
    root, ids = bench_dir
    assert np.load(os.path.join(root, "id_list.npy")).tolist() == ids
    ip, rp, mp = IO.request_paths(root, ids[0])
    assert ip.endswith(os.path.join("GT_3500", str(ids[0]).zfill(12) + "_GT.png")) and os.path.exists(ip)
    assert rp.endswith(os.path.join("Ref_3500", str(ids[0]).zfill(12) + "_ref.png")) and os.path.exists(rp)
    assert mp.endswith(os.path.join("Mask_bbox_3500", str(ids[0]).zfill(12) + "_mask.png")) and os.path.exists(mp)
    img, ref, mask = IO.load_triple(root, ids[0])
    assert img.shape == (64, 64, 3) and ref.shape == (224, 224, 3) and mask.shape == (64, 64)       # exemplar resized by PIL
    assert img.dtype == np.uint8 and set(np.unique(mask)) <= {0, 255}


def test_loader_matches_the_reference_preprocessing(bench_dir):
    """The bytes the loader hands over + the device-side formulas of pbe_b200.preprocess (restated on the host here, oracle
    preprocess_ref) == what COCOImageDataset.__getitem__ computes with torchvision (test_bench_dataset.py:79-98)."""
    import torchvision
    from PIL import Image
    from oracle import preprocess_ref as R
    root, ids = bench_dir
    ld = IO.RequestLoader.from_test_bench(root, batch_size=3, pin=False, drop_last=False)
    got = list(ld)
    assert [b[0] for b in got] == [ids[0:3], ids[3:6], ids[6:7]] and len(ld) == 3
    rid = ids[4]
    _, img, ref, mask = got[1]
    tt = torchvision.transforms
    get_tensor = tt.Compose([tt.ToTensor(), tt.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    get_tensor_clip = tt.Compose([tt.ToTensor(), tt.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])
    ip, rp, mp = IO.request_paths(root, rid)
    image_ref = get_tensor(Image.open(ip).convert("RGB"))
    ref_ref = get_tensor_clip(Image.open(rp).resize((224, 224)).convert("RGB"))
    mask_ref = 1 - tt.ToTensor()(Image.open(mp).convert("L"))
    image, m, inpaint = R.prepare_inpaint(img[1:2], mask[1:2], binarize=False)
    assert torch.equal(image[0], image_ref) and torch.equal(m[0], mask_ref) and torch.equal(inpaint[0], image_ref * mask_ref)
    assert torch.equal(R.normalize_u8(ref[1:2], R.CLIP_MEAN, R.CLIP_STD)[0], ref_ref)


def test_loader_shards_prefetches_and_reports_errors(bench_dir):
    root, ids = bench_dir
    a = [i for b in IO.RequestLoader.from_test_bench(root, 2, rank=0, world=2, pin=False) for i in b[0]]
    b = [i for bb in IO.RequestLoader.from_test_bench(root, 2, rank=1, world=2, pin=False) for i in bb[0]]
    assert a == ids[0::2] and b == ids[1::2]                                   # request i -> rank i mod world
    assert len(IO.RequestLoader(ids, 3, lambda r: None, drop_last=True)) == 2  # DataLoader(drop_last=True), inference_test_bench.py:300
    started, gate = [], threading.Event()

    def slow_fetch(rid):
        started.append(rid)
        if len(started) > 4:
            gate.wait(5)
        z = np.zeros((4, 4, 3), np.uint8)
        return z, z, z[..., 0]

    it = iter(IO.RequestLoader(list(range(12)), 2, slow_fetch, prefetch=2, workers=2, pin=False))
    first = next(it)
    time.sleep(0.3)
    assert first[0] == [0, 1] and len(started) >= 4        # the next batches are being decoded while the consumer holds batch 0
    gate.set()
    assert sum(len(b[0]) for b in it) == 10

    def bad_fetch(rid):
        raise FileNotFoundError(f"no such request {rid}")

    with pytest.raises(FileNotFoundError, match="no such request"):
        list(IO.RequestLoader([1, 2], 2, bad_fetch, pin=False))
    with pytest.raises(TypeError, match="uint8"):
        list(IO.RequestLoader([1], 1, lambda r: (np.zeros((2, 2, 3), np.float32),) * 3, pin=False))


def test_writer_round_trips_pixels_and_applies_the_hook(tmp_path):
    from PIL import Image
    g = torch.Generator().manual_seed(0)
    imgs = torch.randint(0, 256, (5, 32, 40, 3), generator=g, dtype=torch.uint8)
    names = [f"{i:012d}" for i in range(5)]
    with IO.ResultWriter(str(tmp_path / "results"), workers=3, max_in_flight=2) as w:
        w.submit(names[:2], imgs[:2])
        w.submit(names[2:], imgs[2:])
    assert w.written == 5 and w.bytes_written > 0
    for i, n in enumerate(names):                                               # PNG is lossless: the exact bytes come back
        back = np.asarray(Image.open(tmp_path / "results" / (n + ".png")))
        assert back.shape == (32, 40, 3) and np.array_equal(back, imgs[i].numpy())
    with IO.ResultWriter(str(tmp_path / "marked"), transform=lambda a: 255 - a) as w2:      # scripts/inference.py put_watermark hook
        w2.submit(["x"], imgs[:1])
    assert np.array_equal(np.asarray(Image.open(tmp_path / "marked" / "x.png")), 255 - imgs[0].numpy())
    with pytest.raises(ValueError, match="uint8"):
        IO.ResultWriter(str(tmp_path / "bad")).submit(["a"], torch.zeros(1, 4, 4, 3))

    def boom(a):
        raise RuntimeError("disk full")

    w3 = IO.ResultWriter(str(tmp_path / "err"), transform=boom)
    w3.submit(["a"], imgs[:1])
    with pytest.raises(RuntimeError, match="disk full"):
        w3.close()
