"""B200-native ContextUnet / DDPM hot path (drop-in for Tengis0618/CAMELS-Diffusion-Model).

Python mirrors the reference's nn.Module / function API; all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI in include/cdm_b200.h.
"""
__version__ = "0.1.0"

from ._lib import CdmError  # noqa: F401
from .unet import ContextUnet, EmbedFC, ResidualConvBlock, UnetDown, UnetUp  # noqa: F401
from .diffusion import (DDPM, calculate_elbo_and_bpd, calculate_elbo_and_bpd_batch,  # noqa: F401
                        calculate_likelihood, calculate_likelihood_and_elbo, denoise_add_noise, make_schedule, perturb_input, sample_ddpm)
from .metrics import (compare_distributions, compare_power_spectra, pixel_histograms, power_spectra,  # noqa: F401
                      power_spectrum)
from .data import normalize_params, preprocess_maps  # noqa: F401
from .checkpoint import load_checkpoint, load_model, save_checkpoint, save_model  # noqa: F401
