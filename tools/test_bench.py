"""BASELINE config C3 end to end: a COCOEE-shaped test bench of edit triples (image, bbox mask, exemplar), sharded
request i -> rank i mod world (scripts/inference_test_bench.py:295-397 is the loop being mirrored), every device stage
through the library and the host I/O overlapped with it (pbe_b200/io.py):

    loader threads: PNG decode (or synthetic bytes) -> pinned uint8          [RequestLoader, 2 batches ahead]
    main thread   : H2D -> prepare_inpaint / get_tensor_clip (pbe_b200.preprocess) -> encode_first_stage (VAE) ->
                    get_learned_conditioning + proj_out (CLIP front-end) -> Resize(mask) -> PLMSSampler.sample (50 steps,
                    scale 5) -> decode_to_uint8 (VAE + post-processing) -> async D2H          [never synchronises]
    writer threads: wait for the copy's event -> PNG encode -> results/<id>.png              [ResultWriter]

    python tools/test_bench.py --requests 16 --batch 8                       # one GPU, synthetic bytes, no files
    python tools/test_bench.py --requests 64 --dataset /tmp/tb --save /tmp/out   # PNG triples on disk in, PNGs out
    torchrun --nproc-per-node 8 tools/test_bench.py --requests 3500 --batch 8 --save /tmp/out

Weights are the seeded synthetic ones of the oracles (there is no checkpoint in this sandbox) and so is the data; what is
measured is whole-request throughput including host<->device copies and the host I/O, plus per-stage CUDA-event times.
Rank 0 prints one JSON line."""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--requests", type=int, default=16)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--small", action="store_true", help="narrow CI networks instead of v1.yaml (quick functional check)")
    ap.add_argument("--dataset", default=None, help="test-bench directory in the reference's layout (created with synthetic PNG "
                                                    "triples if it has no id_list.npy); default: synthetic bytes, no files")
    ap.add_argument("--save", default=None, help="write results/<id>.png here (threaded PNG writer)")
    ap.add_argument("--io-workers", type=int, default=8)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from oracle import clip_ref as K, preprocess_ref as R, unet_ref as U, vae_ref as V   # synthetic weights / bytes only
    from pbe_b200 import preprocess as P
    from pbe_b200.clip import FrozenCLIPImageEmbedder
    from pbe_b200.diffusion import LatentDiffusion
    from pbe_b200.samplers import PLMSSampler
    from pbe_b200.sharding import shard_requests
    from pbe_b200.vae import AutoencoderKL

    ucfg, vcfg, kcfg = (U.SMALL_CFG, V.SMALL_VAE_CFG, K.SMALL_CLIP_CFG) if args.small else (U.V1_CFG, V.V1_VAE_CFG, K.V1_CLIP_CFG)
    torch.set_num_threads(max(1, (os.cpu_count() or 8) // world))
    t0 = time.perf_counter()
    model = LatentDiffusion(unet_config=dict(params=dict(ucfg)))
    model.load_state_dict({"model.diffusion_model." + k: v for k, v in U.make_state_dict(ucfg, 321).items()}, strict=False)
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=vcfg["ch"], ch_mult=list(vcfg["ch_mult"]),
              num_res_blocks=vcfg["num_res_blocks"], attn_resolutions=[], dropout=0.0)
    model.first_stage_model = AutoencoderKL(ddconfig=dd, embed_dim=4)
    model.first_stage_model.load_state_dict(V.make_state_dict(vcfg, 321), strict=True)
    model.cond_stage_model = FrozenCLIPImageEmbedder(**kcfg)
    model.cond_stage_model.load_state_dict(K.make_state_dict(kcfg, 321), strict=True)
    g = torch.Generator().manual_seed(5)
    model.proj_out = torch.nn.Linear(kcfg["width"], 768)                                  # latent_diffusion.py:112
    model.proj_out.load_state_dict({"weight": torch.randn(768, kcfg["width"], generator=g) / math.sqrt(kcfg["width"]),
                                    "bias": 0.05 * torch.randn(768, generator=g)})
    model = model.to(dev).eval()
    sampler = PLMSSampler(model)
    setup_s = time.perf_counter() - t0

    from pbe_b200 import io as IO
    f = 2 ** (len(vcfg["ch_mult"]) - 1)
    H = W = args.size
    h, w = H // f, W // f
    S_ref = kcfg["image_size"]
    if args.dataset:
        if rank == 0 and not os.path.exists(os.path.join(args.dataset, "id_list.npy")):
            IO.write_synthetic_test_bench(args.dataset, args.requests, size=args.size, seed=7)     # no dataset in the sandbox
        if world > 1:
            dist.barrier()
        loader = IO.RequestLoader.from_test_bench(args.dataset, args.batch, rank, world, prefetch=2, workers=args.io_workers)
        loader.fetch = lambda rid, root=args.dataset: IO.load_triple(root, rid, ref_size=S_ref)
    else:
        def synth(i):   # what PIL would hand over for request i (bytes only; generated on the loader threads)
            img_u8, mask_u8 = R.synthetic_u8_request(1, H, W, seed=321 + i)
            gg = torch.Generator().manual_seed(1000 + i)
            return img_u8[0], torch.randint(0, 256, (S_ref, S_ref, 3), generator=gg, dtype=torch.uint8), mask_u8[0]
        loader = IO.RequestLoader(shard_requests(args.requests, rank, world), args.batch, synth, prefetch=2, workers=args.io_workers)
    writer = IO.ResultWriter(os.path.join(args.save, "results"), workers=args.io_workers, max_in_flight=4) if args.save else None
    stages = ("h2d", "preprocess", "vae_encode", "clip", "sample", "decode_u8", "d2h")
    ev = lambda: torch.cuda.Event(enable_timing=True)
    torch.manual_seed(4321 + rank)                    # x_T is drawn on the device by the sampler (start_code = None in the script)
    all_marks, done, checksum_t = [], 0, torch.zeros((), dtype=torch.int64, device=dev)
    wall0 = None
    host_keep = []
    nb = len(loader)
    for bi, (ids, img_u8, ref_u8, mask_u8) in enumerate(loader):
        B = len(ids)
        if bi == 1 or nb == 1:
            torch.cuda.synchronize()
            wall0 = time.perf_counter()                                                 # the first batch builds the plans
            done = 0
            all_marks = []
        marks = [ev() for _ in range(len(stages) + 1)]
        with torch.no_grad():
            marks[0].record()
            img_d, mask_d, ref_d = (t.to(dev, non_blocking=True) for t in (img_u8, mask_u8, ref_u8))
            marks[1].record()
            _, m_full, inpaint = P.prepare_inpaint(img_d, mask_d, binarize=False)       # test_bench_dataset.py:89-98
            ref_t = P.get_tensor_clip()(ref_d)
            m_lat = P.Resize([h, w])(m_full)                                            # inference_test_bench.py Resize
            marks[2].record()
            z_inp = model.get_first_stage_encoding(model.encode_first_stage(inpaint).mode())
            marks[3].record()
            c = model.proj_out(model.get_learned_conditioning(ref_t))
            marks[4].record()
            samples, _ = sampler.sample(S=args.steps, conditioning=c, batch_size=B, shape=[4, h, w], verbose=False,
                                        unconditional_guidance_scale=5.0, unconditional_conditioning=model.learnable_vector,
                                        eta=0.0, x_T=None, test_model_kwargs=dict(inpaint_image=z_inp, inpaint_mask=m_lat))
            marks[5].record()
            u8 = model.first_stage_model.decode_to_uint8(samples / model.scale_factor)
            marks[6].record()
            if writer is not None:
                writer.submit([str(i).zfill(12) for i in ids], u8)                      # async D2H + PNG encode off-thread
            else:
                hb = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
                hb.copy_(u8, non_blocking=True)
                host_keep = [hb]
            marks[7].record()
            checksum_t += u8.sum(dtype=torch.int64)
        assert u8.shape == (B, H, W, 3) and u8.dtype == torch.uint8
        all_marks.append((B, marks))
        done += B
    torch.cuda.synchronize()
    if writer is not None:
        writer.close()
    wall = time.perf_counter() - (wall0 or time.perf_counter())
    checksum = int(checksum_t.item()) % (1 << 31)
    acc = {s: 0.0 for s in stages}
    for B, marks in all_marks:
        for i, s in enumerate(stages):
            acc[s] += marks[i].elapsed_time(marks[i + 1])
    total_done = done
    if world > 1:
        t = torch.tensor([float(done), wall], device=dev)
        cnt = t.clone()
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_done, wall = int(cnt[0].item()), t[1].item()
    if rank == 0:
        per_img = {s: round(acc[s] / max(done, 1), 3) for s in stages}
        print(json.dumps({
            "workload": f"BASELINE configs[2]: {args.requests} synthetic edit triples, {H}x{W}, PLMS {args.steps} steps, scale 5, "
                        f"micro-batch {args.batch}, request i -> rank i mod {world}; whole request on the GPU from uint8 bytes "
                        f"to uint8 bytes ({'CI-size' if args.small else 'v1.yaml'} networks, seeded random weights)",
            "n_gpus": world, "requests_timed": total_done, "images_per_sec": round(total_done / wall, 3) if wall > 0 else None,
            "host_io": {"input": "PNG triples decoded by loader threads" if args.dataset else "synthetic bytes from loader threads",
                        "output": (f"{writer.written} PNGs, {writer.bytes_written / 1e6:.1f} MB written by {args.io_workers} "
                                   f"writer threads (rank 0)") if writer is not None else "none (D2H copy only)",
                        "overlap": "loader 2 batches ahead; D2H + PNG encode asynchronous; the main thread never synchronises "
                                   "the stream inside the timed loop"},
            "stage_ms_per_image_rank0": per_img, "setup_s": round(setup_s, 1), "checksum_rank0": checksum}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
