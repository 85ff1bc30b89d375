"""CPU-only checks: the C-ABI library builds/loads and exports every symbol the header declares; host
logic (schedule tables, shortcut replay order, snapshot schedule, weight packing); loud failure without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import contextunet_oracle as O
from tests._util import NCF, ROOT, T, load, raw_sd


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from camels_diffusion_model_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    with open(os.path.join(ROOT, "include", "cdm_b200.h")) as fh:
        hdr = fh.read()
    declared = set(re.findall(r"\b(cdm_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    handle = lib.lib()
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in cdm_b200.h but not exported by libcdm_b200.so"
    assert declared == set(lib.EXPORTS)
    assert handle.cdm_version() >= 100


def test_no_gpu_fails_loudly(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.lib().cdm_device_ok() != 0
    assert b"no CPU path" in lib.lib().cdm_last_error() or lib.lib().cdm_last_error()
    import camels_diffusion_model_b200 as cdm
    m = cdm.ContextUnet(1, 128, NCF, 64).eval()
    with pytest.raises(cdm.CdmError):
        m(torch.zeros(1, 1, 64, 64), torch.tensor([0.5]))


def test_module_matches_reference_state_dict_and_seeded_init():
    import camels_diffusion_model_b200 as cdm
    torch.manual_seed(0)
    m = cdm.ContextUnet(1, 128, NCF, 64)
    sd = raw_sd()
    msd = m.state_dict()
    assert list(msd) == list(sd) and len(msd) == 156
    assert all(torch.equal(msd[k], sd[k]) for k in sd)
    assert sum(p.numel() for p in m.parameters()) == 21626881
    m1 = cdm.ContextUnet(1, 128, 1, 64)
    assert sum(p.numel() for p in m1.parameters()) == 21624961


def test_shortcut_draw_order_matches_conv2d_construction():
    import camels_diffusion_model_b200 as cdm
    from camels_diffusion_model_b200 import diffusion as D
    m = cdm.ContextUnet(1, 128, NCF, 64)
    torch.manual_seed(7)
    ref = []
    for _ in range(6):  # T=3 steps x 2 forwards, the order the reference's CFG loop constructs them
        conv = torch.nn.Conv2d(1, 128, kernel_size=1)
        ref.append(torch.cat([conv.weight.detach().view(-1), conv.bias.detach()]))
    torch.manual_seed(7)
    tab = D.draw_shortcut_table(3, 2)
    k = 0
    for i in (3, 2, 1):
        for r in range(2):
            assert torch.equal(tab[i, r].reshape(-1), ref[k])
            k += 1
    torch.manual_seed(7)
    assert torch.equal(m.draw_shortcut(), ref[0])
    lst = [(v[:128], v[128:]) for v in ref]
    assert torch.equal(D.shortcut_table_from_list(lst, 3, 2), tab)


def test_schedule_coef_and_snapshot_tables():
    from camels_diffusion_model_b200 import diffusion as D
    g = load("sampler.npz")
    Tn = int(g["T"])
    b_t, a_t, ab_t = D.make_schedule(Tn, device="cpu")
    assert np.array_equal(ab_t.numpy(), g["sched/ab_t"]) and np.array_equal(b_t.numpy(), g["sched/b_t"])
    coef = D._coef_table(b_t, a_t, ab_t)
    i = 7
    x, e, z = T(g["ew/x"]), T(g["ew/eps"]), T(g["ew/z"])
    assert np.array_equal(((x - e * coef[i, 0]) / coef[i, 1] + coef[i, 2] * z).numpy(), g["ew/denoise_t7"])
    assert D.snapshot_steps(1500) == O.snapshot_steps(1500) and len(D.snapshot_steps(1500)) == 82
    assert D.snapshot_steps(12) == [12, 7, 6, 5, 4, 3, 2, 1]


def test_weight_packing_layouts():
    from camels_diffusion_model_b200 import unet as U
    conv = torch.nn.Conv2d(4, 3, 3, 1, 1)
    w = U._pack_conv3(conv).float()
    assert w.shape == (3, 3, 3, 4)
    assert torch.allclose(w[1, 2, 0, 3], conv.weight[1, 3, 2, 0].detach().to(torch.bfloat16).float())
    ct = torch.nn.ConvTranspose2d(5, 3, 2, 2)
    b = U._pack_convT(ct).float()
    assert b.shape == (12, 5)
    assert torch.allclose(b[(1 * 2 + 0) * 3 + 2, 4], ct.weight[4, 2, 1, 0].detach().to(torch.bfloat16).float())
    bn = torch.nn.BatchNorm2d(3).eval()
    bn.running_mean.uniform_(-1, 1), bn.running_var.uniform_(0.5, 2), bn.weight.data.uniform_(0.5, 2), bn.bias.data.uniform_(-1, 1)
    c2 = torch.nn.Conv2d(4, 3, 3, 1, 1)
    scale, shift = U._fold_bn(c2, bn)
    x = torch.randn(2, 4, 8, 8)
    ref = bn(c2(x))
    got = torch.nn.functional.conv2d(x, c2.weight) * 1  # noqa
    got = torch.nn.functional.conv2d(x, c2.weight, None, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    assert torch.allclose(got, ref, atol=1e-5)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints ONE JSON line with the keys the
    bench contract names; a second rank of a torchrun launch prints nothing and exits 0."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cfg_ddpm_samples_per_sec_64x64_1500steps"
    assert d["unit"] == "samples/s" and d["higher_is_better"] is True and d["value"] > 0
    from oracle import build_ref
    assert d["cpu_baseline"]["kind"] == ("reference" if build_ref.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                        capture_output=True, text=True, env=dict(env, RANK="1", WORLD_SIZE="2"), timeout=60)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_radial_bins_follow_the_reference_rule():
    """metrics._radial_bins is the reference's binning (diffusion_utilities.py:325-356) as a CSR list."""
    from camels_diffusion_model_b200 import metrics as M
    from oracle import metrics_oracle as MO
    for n, dl in ((64, 1.0), (16, 0.5)):
        k_bins, start, items = M._radial_bins(n, dl)
        rng = np.random.RandomState(n)
        power = rng.rand(n * n)
        pk = np.array([power[items[start[b]:start[b + 1]]].mean() if start[b + 1] > start[b] else 0.0
                       for b in range(len(k_bins))]) * dl ** 2
        # feed the same "power" through the oracle's binning by inverting the FFT magnitude: compare bin membership
        kx = 2 * np.pi * np.fft.fftfreq(n, dl)
        kg = np.sqrt(kx[:, None] ** 2 + kx[None, :] ** 2).flatten()
        dk = 2 * np.pi / (n * dl)
        idx = np.array([int(round(v / dk)) for v in kg])
        ref = np.array([power[idx == b].mean() if (idx == b).any() else 0.0 for b in range(len(k_bins))]) * dl ** 2
        np.testing.assert_allclose(pk, ref, rtol=1e-12)
        assert sorted(items.tolist()) == list(range(n * n)) and start[-1] == n * n
        assert len(k_bins) == len(MO.power_spectrum(np.zeros((n, n), np.float32), dl)[0])


def test_header_is_plain_c_and_links(lib, tmp_path):
    """include/cdm_b200.h compiles as C99 and a C program links the library directly (the boundary a non-Python
    host — the reference ported to C++, a Go/Rust FFI — would bind)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "abi", "abi_smoke.c"), "-o", exe, "-L", libdir, "-l:libcdm_b200.so",
           f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "version 200" in r.stdout and "conv3x3(NULL) -> -1" in r.stdout


def test_pack_plan_table_addresses_the_torch_layouts(monkeypatch):
    """The cdm_pack_bf16 table (train._PackPlan) on CPU: emulate the kernel's addressing
    out[i0][i1][i2][i3] = src[off + sum i_k s_k] and compare with the permute / flip expressions it replaces."""
    import camels_diffusion_model_b200 as cdm
    from camels_diffusion_model_b200 import train as TR
    monkeypatch.setattr(TR.L, "pack_bf16", lambda *a: None)
    torch.manual_seed(0)
    m = cdm.ContextUnet(1, 128, 2, 64)
    plan = TR._PackPlan(m)
    tab = plan.table.numpy()
    by_ptr = {p.data_ptr(): p.detach().reshape(-1).numpy() for p in m.parameters()}
    refs = {}
    for name, blk in TR._rcb_list(m):
        for cn, seq in (("c1", blk.conv1), ("c2", blk.conv2)):
            w = seq[0].weight.detach()
            if w.shape[1] != 1:
                refs[f"{name}.{cn}.f"] = w.permute(0, 2, 3, 1).contiguous()
                refs[f"{name}.{cn}.d"] = w.flip(2, 3).permute(1, 2, 3, 0).contiguous()
    w = m.out[0].weight.detach()
    refs["out0.f"], refs["out0.d"] = w.permute(0, 2, 3, 1).contiguous(), w.flip(2, 3).permute(1, 2, 3, 0).contiguous()
    for nm, mod in (("up0", m.up0[0]), ("up1", m.up1.model[0]), ("up2", m.up2.model[0])):
        w = mod.weight.detach()
        refs[nm + ".f"], refs[nm + ".d"] = w.permute(2, 3, 1, 0).contiguous(), w.permute(0, 2, 3, 1).contiguous()
    # up0.0.weight (both layouts) goes through cdm_pack_transpose_bf16 instead of the table: emulate its addressing
    # dst[b][c][r] = src[b][r][c] on the stride arguments
    assert list(plan.P) == list(refs) and len(tab) == 40 and tab.shape[1] == 12 and len(plan.transposes) == 2
    rng = np.random.RandomState(0)
    for key, (src, dst, nb, R, Cc, sb, sr, db, dc) in zip(("up0.f", "up0.d"), plan.transposes):
        flat, ref = by_ptr[src.data_ptr()], refs[key].reshape(-1).numpy()
        assert dst is plan.P[key] and dst.numel() == ref.size == nb * R * Cc
        for _ in range(64):
            b, r, c = rng.randint(nb), rng.randint(R), rng.randint(Cc)
            assert flat[b * sb + r * sr + c] == ref[b * db + c * dc + r], key
    vec = 0
    table_keys = [k for k in plan.P if not k.startswith("up0.")]
    for key, row in zip(table_keys, tab):
        src, _, d1, d2, d3, s0, s1, s2, s3, off, v0, _ = (int(v) for v in row)
        ref = refs[key].reshape(-1).numpy()
        assert v0 == vec and d3 % 8 == 0 and plan.P[key].numel() == ref.size
        vec += ref.size // 8
        for e in rng.randint(0, ref.size, 64):
            i3, t2 = e % d3, e // d3
            i2, t1 = t2 % d2, t2 // d2
            i1, i0 = t1 % d1, t1 // d1
            assert by_ptr[src][off + i0 * s0 + i1 * s1 + i2 * s2 + i3 * s3] == ref[e], key
    assert vec == plan.total_vec


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6])
def test_driver_contexts_match_the_reference_statements(n):
    """Parameter grid / guidance sweep / sensitivity (BASELINE config 4, num_params 1..6): what the drivers hand to
    the sampler equals, bit for bit and call for call, what the reference's module-level statements hand to its
    `sample_ddpm` (vectors recorded by oracle/make_golden_drivers.py exec'ing those lines).  The sampler is a
    recorder here, so no GPU is involved."""
    from camels_diffusion_model_b200 import drivers

    g = load("driver_contexts.npz")
    base = T(g[f"base/{n}"])[0]
    calls = []

    class Recorder:  # the slice of DDPM the drivers use
        n_cfeat = n

        class nn_model:
            h = 64

        def sample_ddpm(self, n_sample=1, size=64, device=None, params=None, guide_w=0.0):
            calls.append((n_sample, params.clone(), float(guide_w)))
            return torch.zeros(n_sample, 1, size, size), None, 0.0, None

    d = Recorder()
    assert torch.equal(drivers.parameter_grid_contexts(base, n), T(g[f"grid/{n}"]))
    _, _, _, ctx = drivers.sample_parameter_grid(d, base)
    assert torch.equal(ctx, T(g[f"grid/{n}"])) and calls[-1][0] == 25 and torch.equal(calls[-1][1], ctx)
    assert calls[-1][2] == 0.0  # the reference's grid call leaves guide_w at its default
    del calls[:]
    out = drivers.guidance_sweep(d, base)
    assert [c[2] for c in calls] == list(g[f"guidance_w/{n}"]) == list(out.keys())
    assert all(c[0] == 5 for c in calls)
    assert torch.equal(torch.stack([c[1] for c in calls]), T(g[f"guidance_params/{n}"]))
    del calls[:]
    assert torch.equal(drivers.sensitivity_contexts(base, n), T(g[f"sensitivity/{n}"]))
    _, ctx, _ = drivers.parameter_sensitivity(d, base, batched=False)  # the reference's call pattern: batch-1 calls
    assert [c[0] for c in calls] == [1] * (5 * n) and all(c[2] == 0.0 for c in calls)
    assert torch.equal(torch.cat([c[1] for c in calls]), T(g[f"sensitivity/{n}"]))
    del calls[:]
    _, ctx_b, _ = drivers.parameter_sensitivity(d, base, batched=True)  # same contexts as one batch
    assert len(calls) == 1 and calls[0][0] == 5 * n and torch.equal(calls[0][1], T(g[f"sensitivity/{n}"]))


def _header_structs():
    """{struct typedef name: [field names]} parsed from include/cdm_b200.h (plain C declarations only)."""
    import re
    src = open(os.path.join(ROOT, "include", "cdm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for body, name in re.findall(r"typedef struct(?:\s+\w+)?\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "const float* a, *b" / "int H, W" / "cdm_forward_args fwd": the declarators after the type
            parts = [p.strip() for p in decl.split(",")]
            first = re.match(r"^(.*?)(\w+)$", parts[0], flags=re.S)
            fields.append(first.group(2))
            fields += [re.sub(r"^[\s\*]*", "", p) for p in parts[1:]]
        out[name] = fields
    return out


def test_ctypes_structs_match_the_c_layout(lib, tmp_path):
    """Every argument struct of include/cdm_b200.h: sizeof and every field's offsetof, as gcc lays them out, equal the
    ctypes mirror in _lib.py (a binding that is one field short makes the library read past the caller's struct)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    structs = _header_structs()
    mirror = {"cdm_conv3x3_args": lib.Conv3x3Args, "cdm_xrank": lib.XrankArgs, "cdm_gemm_args": lib.GemmArgs,
              "cdm_conv_in_args": lib.ConvInArgs, "cdm_conv_out_args": lib.ConvOutArgs,
              "cdm_gn_relu_film_args": lib.GnReluFilmArgs, "cdm_ddpm_step_args": lib.DdpmStepArgs,
              "cdm_perturb_args": lib.PerturbArgs, "cdm_mse_accum_args": lib.MseAccumArgs,
              "cdm_plan_desc": lib.PlanDesc, "cdm_forward_args": lib.ForwardArgs,
              "cdm_sample_step_args": lib.SampleStepArgs, "cdm_gemm_tn_args": lib.GemmTnArgs,
              "cdm_chan_reduce_args": lib.ChanReduceArgs, "cdm_bn_apply_args": lib.BnApplyArgs,
              "cdm_bn_bwd_args": lib.BnBwdArgs, "cdm_gn_bwd_args": lib.GnBwdArgs,
              "cdm_outer_wgrad_args": lib.OuterWgradArgs, "cdm_embed_bwd_args": lib.EmbedBwdArgs}
    assert set(structs) == set(mirror), f"header structs without a ctypes mirror (or vice versa): {set(structs) ^ set(mirror)}"
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "cdm_b200.h"', "int main(void) {"]
    for name, fields in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        lines += [f'  printf("{name}.{f} %zu\\n", offsetof({name}, {f}));' for f in fields]
    lines += ["  return 0;", "}"]
    csrc = tmp_path / "layout.c"
    csrc.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(csrc), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True).stdout.splitlines())
    for name, cls in mirror.items():
        assert int(got[name]) == ctypes_sizeof(cls), f"sizeof({name}): C {got[name]} vs ctypes {ctypes_sizeof(cls)}"
        cfields = structs[name]
        alias = {"inp": "in"}  # `in` is a Python keyword
        pyfields = [alias.get(f[0], f[0]) for f in cls._fields_]
        assert cfields == pyfields, f"{name}: field names/order differ: {cfields} vs {pyfields}"
        for f in cfields:
            pf = {v: k for k, v in alias.items()}.get(f, f)
            assert int(got[f"{name}.{f}"]) == getattr(cls, pf).offset, f"offsetof({name}, {f})"


def ctypes_sizeof(cls):
    import ctypes
    return ctypes.sizeof(cls)


def test_integration_md_stub_is_the_real_struct(lib):
    """The ctypes stub INTEGRATION.md shows a maintainer (cdm_conv3x3_args) has the fields of the binding that the
    layout test above holds to the header — a stub one field short makes the library read past the caller's struct
    (round-1 finding)."""
    import re
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blk = txt[txt.index("class Conv3x3Args"):txt.index("def conv_bn_relu")]
    doc = re.findall(r'\("(\w+)", C\.(\w+)\)', blk)
    real = [(n, t.__name__) for n, t in lib.Conv3x3Args._fields_]
    assert doc == real


def test_oracle_equals_the_live_reference_modules():
    """oracle/_ref (the unmodified reference modules, placed by oracle/build_ref.py) against the oracle restatement,
    live: same seeded weights, same inputs, the same fresh-shortcut draws -> eps identical to fp32 round-off; and the
    reference's own denoise_add_noise / sample_ddpm loop equals the oracle's sampler on a 3-step CFG trajectory."""
    from oracle import build_ref, contextunet_oracle as O
    if not build_ref.available():
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    CU, sps = build_ref.load()
    torch.manual_seed(0)
    ref = CU(in_channels=1, n_feat=128, n_cfeat=6, height=64).eval()
    sd = O.init_state_dict(0, n_cfeat=6)
    assert all(torch.equal(v, sd[k]) for k, v in ref.state_dict().items())
    g = torch.Generator().manual_seed(3)
    x, c = torch.randn(2, 1, 64, 64, generator=g), torch.rand(2, 6, generator=g)
    t = torch.tensor([0.3])
    torch.manual_seed(11)
    with torch.no_grad():
        e_ref = ref(x, t, c)
    torch.manual_seed(11)
    with torch.no_grad():
        e_orc = O.unet_forward(sd, x, t, c, O.draw_shortcut(128), n_cfeat=6)
    assert float((e_ref - e_orc).norm() / e_ref.norm()) < 1e-5
    # the reference's pure-function sampler (code/sample_power_spectra.py:71-110) vs the oracle's, same draw stream
    Tn = 3
    b_t, a_t, ab_t = O.make_schedule(Tn)
    torch.manual_seed(5)
    x_ref = sps.sample_ddpm(ref, n_sample=2, size=64, device=torch.device("cpu"), params=c, guide_w=2.0, timesteps=Tn,
                            b_t=b_t, a_t=a_t, ab_t=ab_t)
    torch.manual_seed(5)
    x_T = torch.randn(2, 1, 64, 64)
    z, shortcuts = torch.zeros(Tn, 2, 1, 64, 64), []
    for i in range(Tn, 0, -1):
        if i > 1:
            z[Tn - i] = torch.randn(2, 1, 64, 64)
        shortcuts.append([O.draw_shortcut(128), O.draw_shortcut(128)])
    with torch.no_grad():
        x_orc, _ = O.sample_ddpm(sd, x_T, c, 2.0, Tn, (b_t, a_t, ab_t), z, shortcuts, n_cfeat=6)
    assert float((x_ref - x_orc).norm() / x_ref.norm()) < 1e-5
