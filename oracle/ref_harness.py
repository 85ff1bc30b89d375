"""ORACLE — test infrastructure only (build container only).

Loads the UNMODIFIED reference code from /root/reference so that golden vectors
can be generated and the restatement in contextunet_oracle.py can be pinned.
Nothing here travels to the GPU box (/root/reference does not exist there); the
-m gpu tests, smoke() and bench.py never import this module.

* ContextUnet.py, code/diffusion_utilities.py import unchanged once stub
  `matplotlib` modules are injected (diffusion_utilities.py:5-6 imports it at top).
* The train_diffusion_*.py scripts parse sys.argv and np.load a dataset that is
  not in the repo at module level, so their functions are lifted with `ast`
   and exec'd in a namespace that supplies the globals they close over.
"""
import ast
import os
import sys
import types

REF = os.environ.get("CDM_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "code"))


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    m = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    anim = types.ModuleType("matplotlib.animation")
    anim.FuncAnimation = object
    anim.PillowWriter = object
    m.pyplot, m.animation = plt, anim
    sys.modules.update({"matplotlib": m, "matplotlib.pyplot": plt, "matplotlib.animation": anim})


def load_modules():
    """Returns (ContextUnet class, diffusion_utilities module) of the reference."""
    _stub_matplotlib()
    for p in (os.path.join(REF, "code"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import diffusion_utilities  # noqa
    import ContextUnet as cu  # noqa
    return cu.ContextUnet, diffusion_utilities


def lift_functions(script, names, namespace):
    """exec the named top-level FunctionDefs of code/<script> inside `namespace`."""
    path = os.path.join(REF, "code", script)
    with open(path) as fh:
        tree = ast.parse(fh.read(), path)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    found = {n.name for n in wanted}
    missing = set(names) - found
    if missing:
        raise KeyError(f"{script}: functions not found: {missing}")
    mod = ast.Module(body=wanted, type_ignores=[])
    exec(compile(mod, path, "exec"), namespace)
    return namespace


def paper_namespace(nn_model, timesteps, n_cfeat, device="cpu"):
    """Namespace with the closures of train_diffusion_paper.py bound to `nn_model`."""
    import time

    import numpy as np
    import torch
    import torch.nn.functional as F
    b_t = (0.02 - 1e-4) * torch.linspace(0, 1, timesteps + 1, device=device) + 1e-4  # :205-214
    a_t = 1 - b_t
    ab_t = torch.cumsum(a_t.log(), dim=0).exp()
    ab_t[0] = 1
    ns = dict(torch=torch, np=np, F=F, time=time, nn_model=nn_model, timesteps=timesteps, n_cfeat=n_cfeat,
              device=torch.device(device), b_t=b_t, a_t=a_t, ab_t=ab_t)
    lift_functions("train_diffusion_paper.py",
                   ["perturb_input", "denoise_add_noise", "sample_ddpm", "sample_ddpm_from_noise",
                    "calculate_likelihood", "calculate_elbo_and_bpd"], ns)
    elbo_ns = dict(torch=torch, np=np, F=F)
    lift_functions("train_diffusion_elbo.py", ["calculate_elbo_and_bpd"], elbo_ns)
    ns["calculate_elbo_and_bpd_batch"] = elbo_ns["calculate_elbo_and_bpd"]
    return ns


class DrawRecorder:
    """Context manager that records, in call order, every tensor the reference draws from the
    global CPU generator through torch.randn / randn_like / rand / randint and the uniform_
    initialisation of the per-forward shortcut conv (diffusion_utilities.py:54)."""

    def __init__(self):
        self.randn, self.shortcuts, self.randint = [], [], []

    def __enter__(self):
        import torch
        self._t = torch
        self._orig = dict(randn_like=torch.randn_like, randn=torch.randn, randint=torch.randint,
                          reset=torch.nn.Conv2d.reset_parameters)
        rec = self

        def randn_like(x, *a, **k):
            z = rec._orig["randn_like"](x, *a, **k)
            rec.randn.append(z.detach().clone())
            return z

        def randn(*a, **k):
            z = rec._orig["randn"](*a, **k)
            rec.randn.append(z.detach().clone())
            return z

        def randint(*a, **k):
            z = rec._orig["randint"](*a, **k)
            rec.randint.append(z.detach().clone())
            return z

        def reset(conv):
            rec._orig["reset"](conv)
            if conv.kernel_size == (1, 1) and conv.in_channels == 1:
                rec.shortcuts.append((conv.weight.detach().view(-1).clone(), conv.bias.detach().clone()))

        torch.randn_like, torch.randn, torch.randint = randn_like, randn, randint
        torch.nn.Conv2d.reset_parameters = reset
        return self

    def __exit__(self, *exc):
        t = self._t
        t.randn_like, t.randn, t.randint = self._orig["randn_like"], self._orig["randn"], self._orig["randint"]
        t.nn.Conv2d.reset_parameters = self._orig["reset"]
        return False
