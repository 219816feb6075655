import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_gpu_sampler_paths as T
from oracle import sampler_ref as S
from pbe_b200.samplers import DDIMSampler
dev = torch.device("cuda:0")
prod, orc = T._ToyProduct(dev), T._ToyOracle()
req = S.synthetic_request(3, 16, 24, seed=77)
for eta, temp in ((1.0, 1.3), (1.0, 1.0), (0.9, 1.3)):
    torch.manual_seed(2024)
    out, inter = DDIMSampler(prod).sample(S=10, eta=eta, temperature=temp, log_every_t=1, **T._kw(req, dev, 3))
    torch.manual_seed(2024)
    rec = []
    ref = S.ddim_sample(orc, 10, req["x_T"], req["c"], req["uc"], 5.0, req["z_inpaint"], req["mask"], record=rec, eta=eta, temperature=temp, rng_device=dev)
    print("eta", eta, "T", temp, "final equal", torch.equal(out.cpu(), ref))
    for k, r in enumerate(rec):
        a = inter["x_inter"][k + 1].cpu(); b = inter["pred_x0"][k + 1].cpu()
        print("  step", k, "index", r["index"], "x_prev maxdiff", (a - r["x_prev"]).abs().max().item(), "nbad", int((a != r["x_prev"]).sum()),
              "pred_x0 maxdiff", (b - r["pred_x0"]).abs().max().item(), "nan", bool(torch.isnan(a).any()), bool(torch.isnan(r["x_prev"]).any()))
