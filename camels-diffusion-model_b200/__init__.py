"""B200-native ContextUnet / DDPM hot path (drop-in for Tengis0618/CAMELS-Diffusion-Model).

Python mirrors the reference's nn.Module / function API; all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI in include/cdm_b200.h.
"""
__version__ = "0.1.0"
