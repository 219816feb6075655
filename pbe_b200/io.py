"""Host-side I/O around the GPU path, overlapped with it: a prefetching request loader and a threaded PNG writer.

The reference keeps the GPU waiting on the host at both ends of a request: ``COCOImageDataset.__getitem__``
(``ldm/data/test_bench_dataset.py:73-100``: three ``Image.open`` + ``ToTensor`` / ``Normalize`` per triple) behind a
``DataLoader(num_workers=4, pin_memory=True)`` (``scripts/inference_test_bench.py:295-301``), and after the sampler a
synchronous ``.cpu()`` followed by six ``Image.fromarray(...).save(...)`` per image on the main thread
(``scripts/inference_test_bench.py:345-397``, ``scripts/inference.py:362-399``).  At 8 x B200 the denoising path delivers
~77 images/s, i.e. one 512x512 PNG every 13 ms: the host side has to be a pipeline of its own.

* :class:`RequestLoader` — worker threads decode the next batches (PIL releases the GIL while decoding) into PINNED uint8
  tensors, ``prefetch`` batches ahead; the bytes go to the device as they are and ``pbe_b200.preprocess`` does ToTensor /
  Normalize / mask / ``image * mask`` there.  The file layout is the reference's: ``GT_3500/<id>_GT.png``,
  ``Ref_3500/<id>_ref.png`` (resized to 224x224 by PIL, as the dataset does), ``Mask_bbox_3500/<id>_mask.png``, ids from
  ``id_list.npy`` zero-filled to 12 digits.  Request i of the list belongs to rank ``i mod world`` (sharding.py).
* :class:`ResultWriter` — the sampler's uint8 images are copied device->host asynchronously into a pinned buffer; a worker
  thread waits for that copy's CUDA event, optionally applies a per-image transform (the scripts' watermark hook,
  ``put_watermark``), PNG-encodes (zlib releases the GIL) and writes ``results/<id>.png``.  The GPU stream is never
  synchronised by the main thread; back-pressure is a bounded number of images in flight.

Nothing here touches the numerics of the path; it is plumbing, and it is pure host code (no CUDA needed to test it).
"""
from __future__ import annotations

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np
import torch


def request_paths(root: str, rid) -> tuple:
    """(image, exemplar, mask) paths of request ``rid`` in the reference's test-bench layout (test_bench_dataset.py:74-91)."""
    name = str(rid).zfill(12)
    return (os.path.join(root, "GT_3500", name + "_GT.png"), os.path.join(root, "Ref_3500", name + "_ref.png"),
            os.path.join(root, "Mask_bbox_3500", name + "_mask.png"))


def load_triple(root: str, rid, ref_size: int = 224):
    """One edit request as uint8 arrays, decoded the way ``COCOImageDataset.__getitem__`` decodes it: RGB image [H,W,3],
    exemplar resized to ``ref_size`` then RGB [S,S,3] (test_bench_dataset.py:79: ``.resize((224,224)).convert("RGB")``),
    L-mode mask [H,W].  Everything after the decode happens on the device."""
    from PIL import Image
    ip, rp, mp = request_paths(root, rid)
    img = np.asarray(Image.open(ip).convert("RGB"))
    ref = np.asarray(Image.open(rp).resize((ref_size, ref_size)).convert("RGB"))
    mask = np.asarray(Image.open(mp).convert("L"))
    return img, ref, mask


class RequestLoader:
    """Iterate over batches of requests, ``prefetch`` batches ahead of the consumer.

    ``fetch(request_id) -> (img_u8 [H,W,3], ref_u8 [S,S,3], mask_u8 [H,W])`` (numpy or torch uint8) is called on worker
    threads; each batch is stacked into pinned tensors.  Yields ``(ids, img [B,H,W,3], ref [B,S,S,3], mask [B,H,W])``.
    Order is the order of ``ids`` (the reference's loader does not shuffle, inference_test_bench.py:299); like its
    ``drop_last=True`` the trailing partial batch is dropped unless ``drop_last=False``."""

    def __init__(self, ids: Sequence, batch_size: int, fetch: Callable, prefetch: int = 2, workers: int = 4,
                 drop_last: bool = False, pin: Optional[bool] = None):
        self.ids = list(ids)
        self.batch_size = int(batch_size)
        self.fetch = fetch
        self.prefetch = max(1, int(prefetch))
        self.workers = max(1, int(workers))
        self.drop_last = drop_last
        self.pin = torch.cuda.is_available() if pin is None else pin
        n = len(self.ids)
        self.batches: List[list] = [self.ids[i:i + self.batch_size] for i in range(0, n, self.batch_size)]
        if drop_last and self.batches and len(self.batches[-1]) < self.batch_size:
            self.batches.pop()

    @classmethod
    def from_test_bench(cls, root: str, batch_size: int, rank: int = 0, world: int = 1, id_list=None, **kw):
        """The reference's COCOEE test bench on disk (``id_list.npy`` + the three folders), this rank's shard of it."""
        from .sharding import shard_requests
        ids = np.load(os.path.join(root, "id_list.npy")).tolist() if id_list is None else list(id_list)
        mine = [ids[i] for i in shard_requests(len(ids), rank, world)]
        return cls(mine, batch_size, lambda rid: load_triple(root, rid), **kw)

    def __len__(self):
        return len(self.batches)

    def _stack(self, items):
        out = []
        for k in range(3):
            t = torch.stack([torch.from_numpy(np.array(it[k], copy=True)) if not isinstance(it[k], torch.Tensor) else it[k]
                             for it in items])
            if t.dtype != torch.uint8:
                raise TypeError(f"fetch() must return uint8 arrays, got {t.dtype}")
            out.append(t.pin_memory() if self.pin else t)
        return out

    def __iter__(self):
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def produce():
            try:
                with ThreadPoolExecutor(self.workers) as pool:
                    for ids in self.batches:
                        if stop.is_set():
                            return
                        items = list(pool.map(self.fetch, ids))       # one batch decoded by all workers, in request order
                        img, ref, mask = self._stack(items)
                        while not stop.is_set():
                            try:
                                q.put((ids, img, ref, mask), timeout=0.1)
                                break
                            except queue.Full:
                                continue
                q.put(None)
            except BaseException as e:      # surface loader errors in the consumer, not in a dead thread
                q.put(e)

        t = threading.Thread(target=produce, name="pbe-request-loader", daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()


class ResultWriter:
    """Asynchronous ``Image.fromarray(x).save(os.path.join(result_path, id + ".png"))`` (inference_test_bench.py:374-376).

    ``submit(ids, images)``: ``images`` uint8 [B,H,W,3] on the host or on a CUDA device.  Device tensors are copied into a
    pinned buffer with ``non_blocking=True`` and the copy's event is handed to the worker, so the calling thread returns at
    once.  At most ``max_in_flight`` batches are pending (then ``submit`` blocks: back-pressure instead of unbounded memory).
    ``transform(np.ndarray[H,W,3]) -> np.ndarray`` runs on the worker before encoding (the scripts' watermark hook).
    ``close()`` waits for everything and re-raises the first worker error."""

    def __init__(self, out_dir: str, workers: int = 8, max_in_flight: int = 8, compress_level: int = 6,
                 transform: Optional[Callable] = None, suffix: str = ".png"):
        self.out_dir = out_dir
        os.makedirs(out_dir, exist_ok=True)
        self.pool = ThreadPoolExecutor(max(1, int(workers)), thread_name_prefix="pbe-png")
        self.sem = threading.Semaphore(max(1, int(max_in_flight)))
        self.compress_level = int(compress_level)
        self.transform = transform
        self.suffix = suffix
        self.futures: list = []
        self.written = 0
        self.bytes_written = 0
        self._lock = threading.Lock()

    def _encode_one(self, arr: np.ndarray, path: str) -> int:
        from PIL import Image
        if self.transform is not None:
            arr = self.transform(arr)
        Image.fromarray(arr).save(path, compress_level=self.compress_level)
        return os.path.getsize(path)

    def _finish(self, event, host: torch.Tensor, ids) -> None:
        try:
            if event is not None:
                event.synchronize()                         # the device->host copy of THIS batch, nothing else
            arr = host.numpy()
            n = 0
            for i, rid in enumerate(ids):
                n += self._encode_one(arr[i], os.path.join(self.out_dir, str(rid) + self.suffix))
            with self._lock:
                self.written += len(ids)
                self.bytes_written += n
        finally:
            self.sem.release()

    def submit(self, ids: Iterable, images: torch.Tensor) -> None:
        ids = list(ids)
        if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[0] != len(ids):
            raise ValueError(f"expected uint8 [B,H,W,C] with B = {len(ids)}, got {images.dtype} {tuple(images.shape)}")
        self.sem.acquire()
        event = None
        if images.device.type == "cuda":
            host = torch.empty(images.shape, dtype=torch.uint8, pin_memory=True)
            host.copy_(images, non_blocking=True)
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream(images.device))
        else:
            host = images.contiguous()
        self.futures.append(self.pool.submit(self._finish, event, host, ids))

    def close(self) -> int:
        for f in self.futures:
            f.result()                                      # re-raises worker exceptions
        self.futures.clear()
        self.pool.shutdown(wait=True)
        return self.written

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def write_synthetic_test_bench(root: str, n: int, size: int = 512, ref_size: int = 256, seed: int = 0) -> list:
    """A COCOEE-shaped test bench on disk with synthetic content (there is no dataset in this sandbox): ``n`` triples in the
    reference's folder / file-name layout plus ``id_list.npy``.  Returns the ids.  Used by the tests and by
    ``tools/test_bench.py --dataset``."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    ids = [int(v) for v in rng.choice(10 ** 9, size=n, replace=False)]
    for sub in ("GT_3500", "Ref_3500", "Mask_bbox_3500"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    for rid in ids:
        ip, rp, mp = request_paths(root, rid)
        base = rng.integers(0, 256, size=(size // 8, size // 8, 3), dtype=np.uint8)           # blocky: compresses like a photo would not,
        img = np.kron(base, np.ones((8, 8, 1), dtype=np.uint8))                               # but decodes at full size
        img = (img.astype(np.int16) + rng.integers(-8, 9, size=img.shape)).clip(0, 255).astype(np.uint8)
        Image.fromarray(img).save(ip)
        Image.fromarray(rng.integers(0, 256, size=(ref_size, ref_size, 3), dtype=np.uint8)).save(rp)
        m = np.zeros((size, size), dtype=np.uint8)
        h0, w0 = (int(v) for v in rng.integers(0, size // 2, size=2))
        hh, ww = (int(v) for v in rng.integers(size // 8, size // 2, size=2))
        m[h0:h0 + hh, w0:w0 + ww] = 255                                                        # bbox mask: 255 inside the hole
        Image.fromarray(m).save(mp)
    np.save(os.path.join(root, "id_list.npy"), np.asarray(ids))
    return ids
