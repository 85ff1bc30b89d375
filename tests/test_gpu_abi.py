"""The drop-in boundary from the other side: a plain-C program (no Python, no torch) drives one ContextUnet eval
forward through the composite C ABI (cdm_plan_create / cdm_plan_embed / cdm_forward_eval) and must produce exactly
the bits ContextUnet.forward of the Python module produces (which runs through the same calls)."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import contextunet_oracle as O
from tests._util import ROOT, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ncf,B", [(6, 3), (2, 1)])
def test_plain_c_consumer_runs_the_forward(tmp_path, ncf, B):
    import camels_diffusion_model_b200 as cdm
    from camels_diffusion_model_b200 import _lib as L
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda_home, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc / CUDA runtime headers not available")
    exe = str(tmp_path / "plan_forward")
    libdir = os.path.dirname(L.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(cuda_home, "include"),
           os.path.join(ROOT, "tests", "abi", "plan_forward.c"), "-o", exe, "-L", libdir, "-l:libcdm_b200.so",
           "-L", os.path.join(cuda_home, "lib64"), "-lcudart", f"-Wl,-rpath,{libdir}",
           f"-Wl,-rpath,{os.path.join(cuda_home, 'lib64')}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sd = O.init_state_dict(3, n_cfeat=ncf)
    g = torch.Generator().manual_seed(ncf)
    for k, v in sd.items():  # non-trivial BatchNorm statistics / affines so that the folding matters
        if k.endswith("running_mean"):
            v.normal_(0, 0.3, generator=g)
        elif k.endswith("running_var"):
            v.uniform_(0.5, 2.0, generator=g)
    x, c = torch.randn(B, 1, 64, 64, generator=g), torch.rand(B, ncf, generator=g)
    t, sc = torch.tensor([0.37]), torch.rand(256, generator=g) * 2 - 1
    with open(tmp_path / "w.bin", "wb") as fh:
        for name in L.plan_tensor_names():
            fh.write(sd[name].contiguous().numpy().astype(np.float32).tobytes())
    with open(tmp_path / "in.bin", "wb") as fh:
        for v in (x, t, c, sc):
            fh.write(v.contiguous().numpy().astype(np.float32).tobytes())
    r = subprocess.run([exe, str(tmp_path / "w.bin"), str(tmp_path / "in.bin"), str(tmp_path / "eps.bin"), str(ncf), str(B)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "plan_forward ok" in r.stdout, r.stdout + r.stderr
    eps_c = torch.from_numpy(np.fromfile(tmp_path / "eps.bin", dtype=np.float32).reshape(B, 1, 64, 64))
    m = cdm.ContextUnet(1, 128, ncf, 64)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    eps_py = m(x.cuda(), t.cuda(), c.cuda(), shortcut=sc).cpu()
    assert torch.equal(eps_c, eps_py), f"C consumer vs Python module: rel-L2 {rel_l2(eps_c, eps_py):.3e}"
    with torch.no_grad():
        ref = O.unet_forward(sd, x, t, c, (sc[:128], sc[128:]), n_cfeat=ncf)
    assert rel_l2(eps_c, ref) < 1e-2
