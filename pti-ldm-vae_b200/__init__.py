"""pti-ldm-vae_b200: the B200-native hot path of Sukikui/PTI-LDM-VAE (AutoencoderKL
encode -> reparameterised sample -> decode, latent head, KL / L1 / L2) behind the reference's
``VAEModel`` API.  Import name: ``pti_ldm_vae_b200`` (see _pkg.py at the repo root)."""
from . import _lib, config, eval_metrics, latent_cache, losses, ops, parallel, trainer, training, transforms  # noqa: F401
from .autoencoderkl import AutoencoderKL, B200AutoencoderKL  # noqa: F401
from .graph import GraphedVAE, PipelinedVAE  # noqa: F401
from .latent_cache import LatentCacheWriter  # noqa: F401
from .loader import load_vae_model  # noqa: F401
from .losses import compute_ar_vae_loss, compute_kl_loss, compute_total_loss, l1_loss, mse_loss  # noqa: F401
from .regression_head import LatentRegressor, regress_from_images  # noqa: F401
from .trainer import TrainStep, flatten_parameters  # noqa: F401
from .training import FlatGrads, TrainRun, VAEFunction  # noqa: F401
from .vae_model import VAEModel  # noqa: F401

build = _lib.build
