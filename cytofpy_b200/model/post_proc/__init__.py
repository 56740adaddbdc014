"""Posterior post-processing; API mirror of reference cytofpy/model/post_proc/__init__.py:1-3."""
from . import sample_params
from . import lam_post
from .sample_params import theta
from .lam_post import sample

__all__ = ['sample_params', 'lam_post', 'theta', 'sample']
