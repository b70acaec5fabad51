"""Drop-in for the reference's ``src.models.gaussian_mf_cavi`` (gaussian_mf_cavi.py:9-241): the
Gaussian model WITHOUT biases.  Same kernels as gaussian_mf_cavi_bias with NULL bias vectors and no
bias passes (two passes per iteration, gaussian_mf_cavi.py:121-178)."""
from __future__ import annotations

from dataclasses import dataclass

from .gaussian_mf_cavi_bias import GaussianMFCAVI as _WithBias
from .ratings import DEFAULT_SEG_LEN


@dataclass
class GaussianMFCAVIConfig:
    n_factors: int = 10          # K (latent dimension)
    sigma2: float = 1.0          # observation noise variance σ²
    eta_theta2: float = 1.0      # prior variance for user factors η_θ²
    eta_beta2: float = 1.0       # prior variance for item factors η_β²
    max_iter: int = 20           # maximum CAVI iterations
    tol: float = 1e-3            # tolerance for validation RMSE improvement
    random_state: int = 42
    verbose: bool = True


class GaussianMFCAVI(_WithBias):
    """Gaussian Matrix Factorization with mean-field VI (CAVI updates), no biases, B200 engine."""

    _table_names = ("m_theta", "V_theta", "m_beta", "V_beta")
    _with_bias = False

    def __init__(self, config: GaussianMFCAVIConfig, device=None, seg_len=DEFAULT_SEG_LEN, shard=None):
        super().__init__(config, device=device, seg_len=seg_len, shard=shard)
