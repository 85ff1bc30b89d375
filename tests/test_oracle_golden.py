"""CPU: the oracle restatement against vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  fp32 on both sides; tolerance 2e-5 relative L2."""
import numpy as np
import pytest
import torch

from oracle import contextunet_oracle as O
from tests._util import NCF, T, cal_sd, load, raw_sd, rel_l2, split_shortcut

TOL = 2e-5


def test_seeded_weights_reproduce():
    g = load("unet_eval.npz")
    assert abs(O.state_dict_checksum(raw_sd()) - float(g["checksum_raw"])) < 1e-6


@pytest.mark.parametrize("name", ["raw", "cal"])
@pytest.mark.parametrize("tn", ["t1", "tB", "cnone"])
def test_unet_forward(name, tn):
    g = load("unet_eval.npz")
    sd = raw_sd() if name == "raw" else cal_sd()
    x, c = T(g["x"]), T(g["c"])
    t = T(g["tB"]) if tn == "tB" else T(g["t1"])
    with torch.no_grad():
        eps = O.unet_forward(sd, x, t, None if tn == "cnone" else c, split_shortcut(g[f"{name}/{tn}/shortcut"]),
                             n_cfeat=NCF)
    assert rel_l2(eps, g[f"{name}/{tn}/eps"]) < TOL


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("tn", ["t1", "tB"])
def test_unet_forward_other_context_widths(n, tn):
    """BASELINE config 4 (n_cfeat 1..5): seeded init and forward against the reference module at that width."""
    g = load("unet_widths.npz")
    init = O.init_state_dict(5, n_cfeat=n)
    assert abs(O.state_dict_checksum(init) - float(g[f"{n}/checksum_init"])) < 1e-6
    sd = O.calibrate_state_dict(init)
    with torch.no_grad():
        eps = O.unet_forward(sd, T(g[f"{n}/x"]), T(g[f"{n}/{tn}/t"]), T(g[f"{n}/c"]),
                             split_shortcut(g[f"{n}/{tn}/shortcut"]), n_cfeat=n)
    assert rel_l2(eps, g[f"{n}/{tn}/eps"]) < TOL


def test_schedule_and_elementwise():
    g = load("sampler.npz")
    b_t, a_t, ab_t = O.make_schedule(int(g["T"]))
    for k, v in (("b_t", b_t), ("a_t", a_t), ("ab_t", ab_t)):
        assert np.array_equal(v.numpy(), g["sched/" + k])
    x, n, z, e, t = (T(g["ew/" + k]) for k in ("x", "noise", "z", "eps", "t"))
    assert np.array_equal(O.perturb_input(x, t, n, ab_t).numpy(), g["ew/perturb_vec"])
    assert np.array_equal(O.perturb_input(x, 5, n, ab_t).numpy(), g["ew/perturb_scalar"])
    assert np.array_equal(O.denoise_add_noise(x, 7, e, z, b_t, a_t, ab_t).numpy(), g["ew/denoise_t7"])
    assert np.array_equal(O.denoise_add_noise(x, 1, e, 0, b_t, a_t, ab_t).numpy(), g["ew/denoise_t1"])


@pytest.mark.parametrize("tag", ["cfg", "plain", "fromnoise"])
def test_sampler(tag):
    g = load("sampler.npz")
    Tn = int(g["T"])
    sd = cal_sd()
    reps = 2 if tag == "cfg" else 1
    flat = [split_shortcut(s) for s in g[f"{tag}/shortcuts"]]
    shortcuts = [flat[k * reps:(k + 1) * reps] for k in range(Tn)]
    params = None if tag == "fromnoise" else T(g["params"])
    gw = 2.0 if tag != "plain" else 0.0
    with torch.no_grad():
        x, inter = O.sample_ddpm(sd, T(g[f"{tag}/x_T"]), params, gw, Tn, O.make_schedule(Tn), T(g[f"{tag}/z"]),
                                 shortcuts, n_cfeat=NCF, save_rate=5 if tag == "fromnoise" else 20)
    assert rel_l2(x, g[f"{tag}/x"]) < 1e-4
    assert inter.shape == g[f"{tag}/inter"].shape
    assert rel_l2(inter, g[f"{tag}/inter"]) < 1e-4


def test_likelihood_and_elbo():
    g = load("likelihood.npz")
    Tn = int(g["T"])
    sd = cal_sd()
    sched = O.make_schedule(Tn)
    maps, prm = T(g["maps"]), T(g["params"])
    sc = [split_shortcut(s) for s in g["nll/shortcuts"]]
    with torch.no_grad():
        n0 = O.likelihood_batch(sd, maps[:2], prm[:2], Tn, sched, T(g["nll/noise_b0"]), sc[:Tn], n_cfeat=NCF)
        n1 = O.likelihood_batch(sd, maps[2:], prm[2:], Tn, sched, T(g["nll/noise_b1"]), sc[Tn:], n_cfeat=NCF)
    nll = (n0.sum().item() + n1.sum().item()) / 3
    assert abs(nll - float(g["nll"])) / float(g["nll"]) < 1e-4
    sc = [split_shortcut(s) for s in g["elbo/shortcuts"]]
    with torch.no_grad():
        e0 = O.elbo_paper_batch(sd, maps[:2], prm[:2], Tn, sched, T(g["elbo/noise_b0"]), sc[:10], n_cfeat=NCF)
        e1 = O.elbo_paper_batch(sd, maps[2:], prm[2:], Tn, sched, T(g["elbo/noise_b1"]), sc[10:], n_cfeat=NCF)
    elbo = (e0.sum().item() + e1.sum().item()) / 3
    assert abs(elbo - float(g["elbo"])) / float(g["elbo"]) < 1e-4
    assert abs(elbo / (64 * 64 * np.log(2)) - float(g["bpd"])) / float(g["bpd"]) < 1e-4
    e, b = O.elbo_bpd_batch(T(g["eb/pred"]), T(g["eb/noise"]), T(g["eb/t"]), sched[2], 64 * 64)
    assert abs(float(e) - float(g["eb/elbo"])) / float(g["eb/elbo"]) < 1e-5
    assert abs(float(b) - float(g["eb/bpd"])) / float(g["eb/bpd"]) < 1e-5


def test_train_step():
    g = load("train_step.npz")
    sd = cal_sd()
    _, _, ab_t = O.make_schedule(1500)
    loss, grads, stats = O.train_step(sd, T(g["x"]), T(g["param"]), T(g["t"]), T(g["noise"]),
                                      split_shortcut(g["shortcut"]), 1500, ab_t, n_cfeat=NCF)
    assert abs(float(loss) - float(g["loss"])) / float(g["loss"]) < 1e-5
    worst = 0.0
    for name, gr in grads.items():
        ref = float(g["gnorm/" + name])
        if ref < 1e-6:  # conv bias in front of a train-mode BatchNorm: the true gradient is exactly zero
            assert float(gr.norm()) < 1e-6, name
            continue
        worst = max(worst, abs(float(gr.norm()) - ref) / (ref + 1e-12))
        assert rel_l2(gr.reshape(-1)[:64], g["gslice/" + name]) < 2e-2, name  # fp32 re-association through ~40 layers
    assert worst < 1e-3
    # BatchNorm running statistics after the step: momentum 0.1, unbiased variance
    for prefix, (mean, uvar) in stats.items():
        rm = 0.9 * sd[prefix + ".running_mean"] + 0.1 * mean
        rv = 0.9 * sd[prefix + ".running_var"] + 0.1 * uvar
        assert rel_l2(rm, g["bn/" + prefix + ".running_mean"]) < 1e-4
        assert rel_l2(rv, g["bn/" + prefix + ".running_var"]) < 1e-4
    # Adam update of one tensor from the oracle gradient
    name = "out.3.weight"
    p0 = sd[name]
    p1, _, _ = O.adam_step(p0, grads[name], torch.zeros_like(p0), torch.zeros_like(p0), 1, float(g["lr"]))
    assert np.allclose(p1.reshape(-1)[:64].numpy(), g["pslice/" + name], rtol=0, atol=1e-7)


def test_metrics_oracle_matches_reference_statistics():
    """oracle/metrics_oracle.py vs the values the reference's own power_spectrum / compare_distributions
    produced on the reference's generated maps (oracle/make_golden_stats.py)."""
    from oracle import metrics_oracle as MO
    g = load("sampler_stats.npz")
    maps = g["x"][:, 0]
    for i in range(len(maps)):
        k, pk = MO.power_spectrum(maps[i])
        np.testing.assert_allclose(k, g["pk_k"], rtol=1e-14)
        np.testing.assert_allclose(pk, g["pk"][i], rtol=1e-12)
    lo, hi = g["hist_lo"], g["hist_hi"]
    bins, pa, pb = MO.histograms((maps - lo) / (hi - lo), (g["hist_other"] - lo) / (hi - lo))
    assert np.array_equal(bins, g["hist_bins"]) and np.array_equal(pa, g["hist_a"]) and np.array_equal(pb, g["hist_b"])


def test_sampler_draw_replay_matches_reference_run():
    """The draw sequence the GPU trajectory test regenerates from the seed is the one the reference consumed."""
    from tests._util import check_replay, replay_sampler_draws
    g = load("sampler_stats.npz")
    x_T, z, tab = replay_sampler_draws(int(g["seed"]), int(g["B"]), int(g["T"]))
    check_replay(g, x_T, z, tab)


def test_data_oracle_matches_reference_statements():
    """oracle/data_oracle.py vs the reference's own data-preparation statements run on the same synthetic arrays."""
    from oracle import data_oracle as DO
    g = load("data_prep.npz")
    for tag, neg in (("pos", False), ("neg", True)):
        maps = DO.synthetic_maps(11, n=30, size=256, negative=neg)
        params = DO.synthetic_params(12, n_sets=2)
        assert np.array_equal(DO.preprocess_maps(maps)[:4].numpy(), g[f"{tag}/maps"])
        for k in (6, 2, 8) if tag == "pos" else (6,):
            tab, pmin, pmax = DO.normalize_params(params, k)
            assert np.array_equal(tab.numpy(), g[f"{tag}/params{k}"])
        assert np.array_equal(pmin, g[f"{tag}/pmin"]) and np.array_equal(pmax, g[f"{tag}/pmax"])
