"""Posterior object returned by AutoencoderKL.encode - same attributes/methods as the reference's
DiagonalGaussianDistribution (src/prediff/utils/distributions.py:26-71): parameters, mean, logvar (clamped to
[-30, 20]), std, var, sample(), mode(). Only tensor plumbing; the moments come from the CUDA encoder."""
from typing import Optional

import torch


class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor, deterministic: bool = False):
        self.parameters = parameters
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device,
                            dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self) -> torch.Tensor:
        return self.mean


def reference_distribution_class():
    """If the reference package is importable (drop-in use under the reference's own LatentDiffusion, whose
    cond_stage_forward does an isinstance check, latent_diffusion.py:370), return its class instead."""
    try:
        from prediff.utils.distributions import DiagonalGaussianDistribution as RefDist  # type: ignore
        return RefDist
    except Exception:
        return DiagonalGaussianDistribution
