"""`cytofpy_b200.model` — same public names as reference `cytofpy.model`
(cytofpy/model/__init__.py:1-2 re-exports everything in fit.py / Model.py)."""
import torch

# The reference flips torch's default dtype to float64 at import
# (cytofpy/model/Model.py:16, fit.py:10, model_param.py:9, vae.py:5) and its callers rely on
# it (e.g. torch.tensor(...) literals in driver scripts).  A drop-in does the same.
torch.set_default_dtype(torch.float64)

from .priors import (compute_Z, default_priors, gen_beta_est, logit, prob_miss,  # noqa: E402
                     solve_beta)
from .variational import ModelParam, VAE  # noqa: E402
from .core import Model  # noqa: E402
from .train import check_nan_in_grad, device_minibatch, fit, update  # noqa: E402
from . import post_proc  # noqa: E402

__all__ = ['compute_Z', 'default_priors', 'gen_beta_est', 'logit', 'prob_miss', 'solve_beta',
           'ModelParam', 'VAE', 'Model', 'check_nan_in_grad', 'device_minibatch', 'fit', 'update', 'post_proc']
