"""ORACLE — test infrastructure only.  Recipe for `oracle/_ref/`: the UNMODIFIED reference modules of the hot path,
placed where the GPU box can import them.

The reference is pure Python (no build system, nothing to compile), so "building" it is locating its three modules
where they lie under /root/reference and writing byte-identical files into oracle/_ref/ — git-ignored (never part of
the history, never edited), not gpurun-ignored (it travels to the GPU box like the built libcdm_b200.so):

  ContextUnet.py                  ContextUnet (ContextUnet.py:6-60)
  code/diffusion_utilities.py     ResidualConvBlock / UnetDown / UnetUp / EmbedFC (:13-145), power_spectrum (:302-368)
  code/sample_power_spectra.py    the pure-function sampler sample_ddpm / denoise_add_noise (:64-110)

`load()` imports them (matplotlib is absent from the image and diffusion_utilities.py:5-6 imports it at module top:
stub modules are injected first, as SURVEY.md §8c describes).  Users: bench.py's CPU legs (`--impl reference`,
`cpu_baseline`, kind "reference") and the tests that pin the oracle.  The product never imports this.

    python oracle/build_ref.py        # run by __graft_entry__.build() whenever /root/reference is present
"""
import hashlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("CDM_REFERENCE", "/root/reference")
FILES = ["ContextUnet.py", "code/diffusion_utilities.py", "code/sample_power_spectra.py"]


def build():
    """-> OUT if the reference is present here (files written / refreshed), else None (the GPU box: prebuilt files)."""
    if not os.path.isdir(os.path.join(REF, "code")):
        return OUT if available() else None
    os.makedirs(OUT, exist_ok=True)
    manifest = []
    for rel in FILES:
        with open(os.path.join(REF, rel), "rb") as fh:
            data = fh.read()
        with open(os.path.join(OUT, os.path.basename(rel)), "wb") as fh:
            fh.write(data)
        manifest.append(f"{hashlib.sha256(data).hexdigest()}  {rel}")
    with open(os.path.join(OUT, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(manifest) + "\n")
    return OUT


def available():
    return all(os.path.exists(os.path.join(OUT, os.path.basename(f))) for f in FILES)


def load():
    """-> (ContextUnet class, sample_power_spectra module) of the reference, imported from oracle/_ref."""
    if not available():
        raise ImportError("oracle/_ref is empty: run `python oracle/build_ref.py` where /root/reference exists")
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        anim = types.ModuleType("matplotlib.animation")
        anim.FuncAnimation = anim.PillowWriter = object
        m.pyplot, m.animation = plt, anim
        sys.modules.update({"matplotlib": m, "matplotlib.pyplot": plt, "matplotlib.animation": anim})
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import ContextUnet as cu  # noqa: E402
    import sample_power_spectra as sps  # noqa: E402
    return cu.ContextUnet, sps


if __name__ == "__main__":
    print(build())
