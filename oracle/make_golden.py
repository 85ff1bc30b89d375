"""ORACLE — test infrastructure only (build container only).

Generates tests/golden/*.npz by executing the UNMODIFIED reference code from
/root/reference (ContextUnet.py, code/diffusion_utilities.py, and the functions
ast-lifted from code/train_diffusion_paper.py / train_diffusion_elbo.py) on CPU fp32
with seeded synthetic weights and inputs.  Weights are not stored (86 MB): they are
`torch.manual_seed(seed); ContextUnet(1,128,n_cfeat,64)`, which contextunet_oracle.
init_state_dict reproduces; a checksum is stored so a torch change is detected.

    python oracle/make_golden.py        # rewrites tests/golden/
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import contextunet_oracle as O  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
SEED = 0
NCF = 6


def _norm_keys(sd):
    return [k for k in sd if k.endswith(".1.weight") or k.endswith(".1.bias") or "running_" in k]


def build_models():
    """Reference module with (a) raw seeded init, (b) 'calibrated' norm layers."""
    CU, _ = RH.load_modules()
    torch.manual_seed(SEED)
    raw = CU(1, 128, NCF, 64).eval()
    assert abs(O.state_dict_checksum(raw.state_dict()) - O.state_dict_checksum(O.init_state_dict(SEED, n_cfeat=NCF))) < 1e-6
    cal = CU(1, 128, NCF, 64)
    cal.load_state_dict(O.calibrate_state_dict(raw.state_dict()))
    cal.train()
    torch.manual_seed(SEED + 1)
    with torch.no_grad():
        for _ in range(6):  # populate BatchNorm running statistics with train-mode reference forwards
            xb = torch.randn(4, 1, 64, 64) * 1.5
            cal(xb, torch.rand(4), torch.rand(4, NCF))
    cal.eval()
    return CU, raw, cal


def small_state(sd):
    """The norm-layer tensors that differ from the seeded init (what a test needs to rebuild `cal`)."""
    return {"sd/" + k: sd[k].numpy() for k in _norm_keys(sd)}


def gen_unet(raw, cal):
    out = {}
    torch.manual_seed(SEED + 10)
    x = torch.randn(2, 1, 64, 64)
    c = torch.rand(2, NCF)
    t1 = torch.tensor([0.37])
    tB = torch.tensor([0.9, 0.05])
    out.update(x=x.numpy(), c=c.numpy(), t1=t1.numpy(), tB=tB.numpy())
    for name, m in (("raw", raw), ("cal", cal)):
        for tn, t in (("t1", t1), ("tB", tB)):
            with RH.DrawRecorder() as rec, torch.no_grad():
                y = m(x, t, c)
            out[f"{name}/{tn}/eps"] = y.numpy()
            out[f"{name}/{tn}/shortcut"] = torch.cat(rec.shortcuts[0]).numpy()
        with RH.DrawRecorder() as rec, torch.no_grad():
            y = m(x, t1, None)
        out[f"{name}/cnone/eps"] = y.numpy()
        out[f"{name}/cnone/shortcut"] = torch.cat(rec.shortcuts[0]).numpy()
    out.update(small_state(cal.state_dict()))
    out["checksum_raw"] = np.float64(O.state_dict_checksum(raw.state_dict()))
    np.savez_compressed(os.path.join(GOLD, "unet_eval.npz"), **out)
    print("unet_eval.npz", len(out), "arrays")


def gen_sampler(cal):
    """sample_ddpm (CFG and plain) and sample_ddpm_from_noise(params=None) of train_diffusion_paper.py, T=12."""
    T = 12
    ns = RH.paper_namespace(cal, T, NCF)
    out = {"T": np.int64(T)}
    torch.manual_seed(SEED + 20)
    params = torch.rand(2, NCF)
    out["params"] = params.numpy()
    for tag, gw in (("cfg", 2.0), ("plain", 0.0)):
        torch.manual_seed(SEED + 21)
        with RH.DrawRecorder() as rec:
            x, inter, _, _ = ns["sample_ddpm"](n_sample=2, size=64, device=torch.device("cpu"), params=params,
                                               guide_w=gw)
        out[f"{tag}/x_T"] = rec.randn[0].numpy()
        out[f"{tag}/z"] = np.stack([z.numpy() for z in rec.randn[1:]] + [np.zeros_like(rec.randn[0].numpy())])
        out[f"{tag}/shortcuts"] = np.stack([torch.cat(s).numpy() for s in rec.shortcuts])
        out[f"{tag}/x"] = x.numpy()
        out[f"{tag}/inter"] = inter
        out[f"{tag}/guide_w"] = np.float32(gw)
    torch.manual_seed(SEED + 22)
    noise_images = torch.randn(2, 1, 64, 64)
    with RH.DrawRecorder() as rec:
        x, inter, _, _ = ns["sample_ddpm_from_noise"](noise_images, params=None, save_rate=5, guide_w=2.0)
    out["fromnoise/x_T"] = noise_images.numpy()
    out["fromnoise/z"] = np.stack([z.numpy() for z in rec.randn] + [np.zeros((2, 1, 64, 64), np.float32)])
    out["fromnoise/shortcuts"] = np.stack([torch.cat(s).numpy() for s in rec.shortcuts])
    out["fromnoise/x"] = x.numpy()
    out["fromnoise/inter"] = inter
    # elementwise closures
    torch.manual_seed(SEED + 23)
    xx, nn_, zz, ee = (torch.randn(3, 1, 64, 64) for _ in range(4))
    tt = torch.tensor([1, 7, 12])
    out["ew/x"], out["ew/noise"], out["ew/z"], out["ew/eps"], out["ew/t"] = (xx.numpy(), nn_.numpy(), zz.numpy(),
                                                                             ee.numpy(), tt.numpy())
    out["ew/perturb_vec"] = ns["perturb_input"](xx, tt, nn_).numpy()
    out["ew/perturb_scalar"] = ns["perturb_input"](xx, 5, nn_).numpy()
    out["ew/denoise_t7"] = ns["denoise_add_noise"](xx, 7, ee, zz).numpy()
    out["ew/denoise_t1"] = ns["denoise_add_noise"](xx, 1, ee, 0).numpy()
    for k in ("b_t", "a_t", "ab_t"):
        out["sched/" + k] = ns[k].numpy()
    np.savez_compressed(os.path.join(GOLD, "sampler.npz"), **out)
    print("sampler.npz", len(out), "arrays")


def gen_likelihood(cal):
    """calculate_likelihood / calculate_elbo_and_bpd (paper.py) on a ragged 2-batch loader, T=10; per-batch ELBO (elbo.py)."""
    T = 10
    ns = RH.paper_namespace(cal, T, NCF)
    torch.manual_seed(SEED + 30)
    maps = torch.rand(3, 1, 64, 64)
    prm = torch.rand(3, NCF)
    loader = [(maps[:2], prm[:2]), (maps[2:], prm[2:])]  # ragged last batch, like DataLoader without drop_last
    out = {"T": np.int64(T), "maps": maps.numpy(), "params": prm.numpy()}
    torch.manual_seed(SEED + 31)
    with RH.DrawRecorder() as rec:
        nll = ns["calculate_likelihood"](cal, loader, T, torch.device("cpu"), ns["ab_t"], ns["b_t"], ns["a_t"])
    out["nll"] = np.float64(nll)
    out["nll/noise_b0"] = np.stack([z.numpy() for z in rec.randn[:T]])
    out["nll/noise_b1"] = np.stack([z.numpy() for z in rec.randn[T:]])
    out["nll/shortcuts"] = np.stack([torch.cat(s).numpy() for s in rec.shortcuts])
    torch.manual_seed(SEED + 32)
    with RH.DrawRecorder() as rec:
        elbo, bpd = ns["calculate_elbo_and_bpd"](cal, loader, T, torch.device("cpu"), ns["ab_t"], ns["b_t"], ns["a_t"])
    out["elbo"], out["bpd"] = np.float64(elbo), np.float64(bpd)
    out["elbo/noise_b0"] = np.stack([z.numpy() for z in rec.randn[:10]])
    out["elbo/noise_b1"] = np.stack([z.numpy() for z in rec.randn[10:]])
    out["elbo/shortcuts"] = np.stack([torch.cat(s).numpy() for s in rec.shortcuts])
    torch.manual_seed(SEED + 33)
    pred, noise = torch.randn(3, 1, 64, 64), torch.randn(3, 1, 64, 64)
    tt = torch.tensor([2, 9, 10])
    e, b = ns["calculate_elbo_and_bpd_batch"](maps, pred, noise, tt, ns["b_t"], ns["a_t"], ns["ab_t"], 64 * 64)
    out["eb/pred"], out["eb/noise"], out["eb/t"] = pred.numpy(), noise.numpy(), tt.numpy()
    out["eb/elbo"], out["eb/bpd"] = np.float64(e), np.float64(b)
    np.savez_compressed(os.path.join(GOLD, "likelihood.npz"), **out)
    print("likelihood.npz", len(out), "arrays; nll", nll, "elbo", elbo)


def gen_train(CU, cal):
    """One training step of train_diffusion_paper.py:349-366 (loop body restated verbatim) on the reference module."""
    import torch.nn.functional as F
    T, lr = 1500, 1e-5
    m = CU(1, 128, NCF, 64)
    m.load_state_dict(cal.state_dict())
    m.train()
    optim = torch.optim.Adam(m.parameters(), lr=lr)
    b_t, a_t, ab_t = O.make_schedule(T)
    torch.manual_seed(SEED + 40)
    x = torch.rand(4, 1, 64, 64)
    param = torch.rand(4, NCF)
    with RH.DrawRecorder() as rec:
        optim.zero_grad()
        noise = torch.randn_like(x)
        t = torch.randint(1, T + 1, (x.shape[0],))
        x_pert = ab_t.sqrt()[t, None, None, None] * x + (1 - ab_t[t, None, None, None]) * noise
        pred_noise = m(x_pert, t / T, param)
        loss = F.mse_loss(pred_noise, noise)
        loss.backward()
        optim.step()
    out = {"x": x.numpy(), "param": param.numpy(), "noise": noise.numpy(), "t": t.numpy(),
           "shortcut": torch.cat(rec.shortcuts[0]).numpy(), "loss": np.float64(loss.item()), "lr": np.float64(lr),
           "pred_noise": pred_noise.detach().numpy()}
    sd_after = m.state_dict()
    for name, p in m.named_parameters():
        g = p.grad
        out["gnorm/" + name] = np.float64(g.norm().item())
        out["gslice/" + name] = g.reshape(-1)[:64].numpy().copy()
        out["pslice/" + name] = sd_after[name].reshape(-1)[:64].numpy().copy()
    for k in sd_after:
        if "running_" in k or "num_batches" in k:
            out["bn/" + k] = sd_after[k].numpy()
    np.savez_compressed(os.path.join(GOLD, "train_step.npz"), **out)
    print("train_step.npz", len(out), "arrays; loss", loss.item())


if __name__ == "__main__":
    assert RH.available(), "reference checkout not found"
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    CU, raw, cal = build_models()
    gen_unet(raw, cal)
    gen_sampler(cal)
    gen_likelihood(cal)
    gen_train(CU, cal)
