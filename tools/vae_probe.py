"""Time AutoencoderKL.decode / .encode (v1.yaml VAE) through the C ABI at the bench geometry: B latents of 64x64 <-> 512x512."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import vae_ref as V          # synthetic weights / latents only
from pbe_b200.vae import AutoencoderKL

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hw = int(sys.argv[2]) if len(sys.argv) > 2 else 64
out_json = sys.argv[3] if len(sys.argv) > 3 else None
dev = torch.device("cuda:0")
cfg = V.V1_VAE_CFG
dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
          num_res_blocks=2, attn_resolutions=[], dropout=0.0)
vae = AutoencoderKL(ddconfig=dd, embed_dim=4)
vae.load_state_dict(V.make_state_dict(cfg, 321), strict=True)
vae = vae.to(dev).eval()
z = V.synthetic_latents(B, hw, hw, seed=1).to(dev)
for _ in range(3):
    img = vae.decode(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    img = vae.decode(z)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
prof = vae.profile(z); prof = vae.profile(z)
flops = sum(r["flops"] for r in prof)
fam = {}
for r in prof:
    f = fam.setdefault(r["family"], [0.0, 0.0, 0]); f[0] += r["ms"]; f[1] += r["flops"]; f[2] += 1
print(f"VAE decode B={B} {hw}x{hw} -> {8*hw}x{8*hw}: {ms:.2f} ms per batch, {ms/B:.2f} ms per image, "
      f"{flops/B/1e12:.2f} TFLOP per image, {flops/ms/1e9:.0f} TFLOP/s; {len(prof)} ops")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:12s} n={v[2]:4d} {v[0]:8.3f} ms" + (f"  {v[1]/v[0]/1e9:7.0f} TFLOP/s" if v[1] else ""))
for r in sorted(prof, key=lambda r: -r["ms"])[:12]:
    print(f"    {r['name']:34s} {r['ms']*1e3:8.1f} us" + (f"  {r['flops']/r['ms']/1e9:7.0f} TFLOP/s" if r["flops"] else ""))
# ---- encode: B images of 8*hw x 8*hw -> moments
x = V.synthetic_images(B, 8 * hw, 8 * hw, seed=2).to(dev)
for _ in range(3):
    post = vae.encode(x)
torch.cuda.synchronize()
e0.record()
for _ in range(n):
    post = vae.encode(x)
e1.record(); torch.cuda.synchronize()
ems = e0.elapsed_time(e1) / n
eprof = vae.profile(x, encode=True); eprof = vae.profile(x, encode=True)
eflops = sum(r["flops"] for r in eprof)
efam = {}
for r in eprof:
    f = efam.setdefault(r["family"], [0.0, 0.0, 0]); f[0] += r["ms"]; f[1] += r["flops"]; f[2] += 1
print(f"VAE encode B={B} {8*hw}x{8*hw} -> {hw}x{hw}: {ems:.2f} ms per batch, {ems/B:.2f} ms per image, "
      f"{eflops/B/1e12:.2f} TFLOP per image, {eflops/ems/1e9:.0f} TFLOP/s; {len(eprof)} ops")
for k, v in sorted(efam.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:12s} n={v[2]:4d} {v[0]:8.3f} ms" + (f"  {v[1]/v[0]/1e9:7.0f} TFLOP/s" if v[1] else ""))
for r in sorted(eprof, key=lambda r: -r["ms"])[:8]:
    print(f"    {r['name']:34s} {r['ms']*1e3:8.1f} us" + (f"  {r['flops']/r['ms']/1e9:7.0f} TFLOP/s" if r["flops"] else ""))
if out_json:
    json.dump(dict(B=B, hw=hw, ms_per_batch=ems, ms_per_image=ems / B, tflop_per_image=eflops / B / 1e12, ops=eprof),
              open(out_json.replace(".json", "_encode.json"), "w"), indent=1)
if out_json:
    json.dump(dict(B=B, hw=hw, ms_per_batch=ms, ms_per_image=ms / B, tflop_per_image=flops / B / 1e12, ops=prof), open(out_json, "w"), indent=1)
