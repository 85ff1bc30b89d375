"""Checkpoint I/O: `.pth` interchange with the reference layout and bit-exact resume."""
import os

import pytest
import torch

from oracle import contextunet_oracle as O
from tests._util import NCF, cal_sd, make_model

pytestmark = pytest.mark.gpu


def _batch(seed, n=4):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(n, 1, 64, 64, generator=g).cuda(), torch.rand(n, NCF, generator=g).cuda())


def test_pth_roundtrip_reference_layout(tmp_path):
    """A file written the way the reference writes it (torch.save(state_dict)) loads, and ours loads back into a plain
    state_dict with the reference's 156 keys and shapes."""
    import camels_diffusion_model_b200 as cdm
    sd = cal_sd()
    ref_file = tmp_path / "model_epoch_3.pth"
    torch.save(sd, ref_file)  # train_diffusion_paper.py:478
    m = cdm.ContextUnet(1, 128, NCF, 64)
    cdm.load_model(m, str(ref_file))
    m = m.cuda().eval()
    ours = tmp_path / "ours.pth"
    cdm.save_model(m, str(ours))
    back = torch.load(ours, map_location="cpu")  # sample_power_spectra.py:188
    assert list(back.keys()) == list(sd.keys()) and len(back) == 156
    for k in sd:
        assert back[k].shape == sd[k].shape and back[k].dtype == sd[k].dtype and torch.equal(back[k], sd[k]), k


@pytest.mark.parametrize("kind", ["fused_adam", "graphed"])
def test_resume_continues_the_run(tmp_path, kind):
    """2 steps, checkpoint, 1 more step == load the checkpoint into fresh objects, 1 step.
    * what is restored (weights, BatchNorm buffers, Adam moments, step count) is bit-exact;
    * the next step's LOSS is identical: the forward pass is deterministic, so this pins weights, BatchNorm
      buffers, the restored CPU generator (t and shortcut draws) and the Philox noise counter together;
    * the Adam moments after that step agree to the run-to-run noise of the backward pass (the split-K weight
      gradient accumulates with fp32 atomics, as cuDNN's does; a lost moment would be off by ~90 %).
    Parameters after further steps are not compared: at random init Adam's first updates are sign-like, which
    amplifies last-bit gradient noise to a visible fraction of lr in two identical uninterrupted runs as well."""
    import camels_diffusion_model_b200 as cdm
    from camels_diffusion_model_b200 import diffusion as D, train as TR
    T = 1500
    ab_t = D.make_schedule(T)[2]

    def fresh():
        m = make_model(cal_sd()).train()
        if kind == "graphed":
            return m, TR.GraphedTrainStep(m, 4, T, ab_t, lr=1e-4, seed=5)
        return m, TR.FusedAdam(m.parameters(), lr=1e-4)

    def run(m, opt, k0, k1):
        loss = None
        for k in range(k0, k1):
            x, p = _batch(100 + k)
            if kind == "graphed":
                loss = opt(x, p)
            else:
                g = torch.Generator(device="cuda").manual_seed(k)
                noise = torch.randn(x.shape, device="cuda", generator=g)
                loss = TR.training_step(m, opt, x, p, T, ab_t, noise=noise)
        torch.cuda.synchronize()
        return float(loss)

    torch.manual_seed(7)
    m1, o1 = fresh()
    run(m1, o1, 0, 2)
    path = str(tmp_path / "resume.pt")
    cdm.save_checkpoint(path, m1, o1, epoch=3, step=2, extra={"lrate": 1e-4})
    loss1 = run(m1, o1, 2, 3)

    torch.manual_seed(999)  # a different generator state: load_checkpoint must restore the saved one
    m2, o2 = fresh()
    epoch, step, extra = cdm.load_checkpoint(path, m2, o2)
    assert (epoch, step, extra["lrate"]) == (3, 2, 1e-4)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    for k, v in m2.state_dict().items():
        assert torch.equal(v.cpu(), ck["model"][k]), k
    for i, st in o2.state_dict()["state"].items():
        assert torch.equal(st["exp_avg"].cpu(), ck["optim"]["state"][i]["exp_avg"]) and int(st["step"]) == 2
        assert torch.equal(st["exp_avg_sq"].cpu(), ck["optim"]["state"][i]["exp_avg_sq"])
    loss2 = run(m2, o2, 2, 3)
    assert loss1 == loss2, (loss1, loss2)
    # every reduction of the step is fixed-order (BatchNorm statistics, split-K weight gradients through workspaces,
    # GroupNorm sums): the continued run is BIT-identical to the uninterrupted one, parameters included
    rng = torch.get_rng_state()  # both continuations draw their t / shortcut values from the same generator state
    run(m1, o1, 3, 5)
    torch.set_rng_state(rng)
    run(m2, o2, 3, 5)
    for (k, a), b in zip(m1.state_dict().items(), m2.state_dict().values()):
        assert torch.equal(a, b), k
    s1, s2 = o1.state_dict()["state"], o2.state_dict()["state"]
    assert s1.keys() == s2.keys() and len(s1) == 102
    for i in s1:
        assert int(s1[i]["step"]) == int(s2[i]["step"]) == 5
        for key in ("exp_avg", "exp_avg_sq"):
            assert torch.equal(s1[i][key], s2[i][key]), (i, key)


def test_optimizer_state_interchanges_with_torch_adam(tmp_path):
    """FusedAdam state_dict -> torch.optim.Adam (the reference's optimiser, :318) and back."""
    from camels_diffusion_model_b200 import diffusion as D, train as TR
    m = make_model(cal_sd()).train()
    opt = TR.FusedAdam(m.parameters(), lr=1e-4)
    x, p = _batch(1)
    TR.training_step(m, opt, x, p, 1500, D.make_schedule(1500)[2])
    ref_opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    ref_opt.load_state_dict(opt.state_dict())
    st = ref_opt.state[next(iter(m.parameters()))]
    assert int(st["step"]) == 1 and st["exp_avg"].abs().sum() > 0
    opt2 = TR.FusedAdam(m.parameters(), lr=1e-4)
    opt2.load_state_dict(ref_opt.state_dict())
    assert opt2._step == 1
