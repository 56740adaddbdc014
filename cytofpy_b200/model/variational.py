"""Mean-field variational family over the global parameters and the
per-(sample, marker) imputation parameters.

API mirror of reference cytofpy/model/model_param.py:30-92 (ModelParam) and
cytofpy/model/vae.py:7-41 (VAE); the arithmetic is restated with explicit
formulas (no torch.distributions transforms) because the same formulas are what
the fused device code evaluates.  Everything is fp64, on the model's device.
"""
import math

import torch
from torch.distributions import Normal

F64 = torch.float64
SUPPORTS = ('positive', 'unit_interval', 'simplex', 'real')
_TINY = torch.finfo(F64).tiny
_ONE_MINUS_EPS = 1.0 - torch.finfo(F64).eps
HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def stick_breaking(x):
    """R^n -> interior of the (n+1)-simplex.  Same map as
    torch.distributions.transforms.StickBreakingTransform (used by reference
    model_param.py:12,62), including its clipped sigmoid."""
    n = x.shape[-1]
    shift = torch.log(torch.arange(n, 0, -1, dtype=x.dtype, device=x.device))
    z = torch.clamp(torch.sigmoid(x - shift), min=_TINY, max=_ONE_MINUS_EPS)
    rest = torch.cumprod(1 - z, -1)
    lead = torch.cat([z, torch.ones_like(z[..., :1])], -1)
    tail = torch.cat([torch.ones_like(z[..., :1]), rest], -1)
    return lead * tail


def stick_breaking_logdet(x, y):
    """log|det J| of stick_breaking at x (y = stick_breaking(x)); reference
    model_param.py:77-78."""
    n = x.shape[-1]
    xs = x - torch.log(torch.arange(n, 0, -1, dtype=x.dtype, device=x.device))
    return (-xs + torch.nn.functional.logsigmoid(xs) + torch.log(y[..., :-1])).sum(-1)


class ModelParam:
    """One block of global parameters: vp = stack([m, log_s]) (reference
    model_param.py:30-44).  When `m`/`s` are omitted they are drawn N(0,1) from
    torch's global CPU generator in the reference's order, so a seeded
    construction gives the reference's initial state."""

    def __init__(self, size, support, m=None, s=None, device=None):
        assert support in SUPPORTS
        m = torch.randn(size, dtype=F64) if m is None else torch.as_tensor(m, dtype=F64)
        log_s = torch.randn(size, dtype=F64) if s is None else torch.as_tensor(s, dtype=F64).log()
        self.vp = torch.nn.Parameter(torch.stack([m, log_s]).to(device))
        self.size = size
        self.support = support

    def __deepcopy__(self, memo):
        # snapshots (fit's trace / backup, reference fit.py:122-133) copy only this block even
        # when vp is a view into the fused step's flat parameter buffer
        new = ModelParam.__new__(ModelParam)
        new.vp = torch.nn.Parameter(self.vp.detach().clone())
        new.size, new.support = self.size, self.support
        return new

    def to(self, device):
        """Detached copy of this block on `device` (fit returns CPU copies like the reference)."""
        new = ModelParam.__new__(ModelParam)
        new.vp = torch.nn.Parameter(self.vp.detach().to(device, copy=True))
        new.size, new.support = self.size, self.support
        return new

    # pickles hold only this block, on the CPU (vp may be a CUDA view into the fused step's flat buffer:
    # the default pickle would carry that whole storage and need a GPU to load); the reference's callers
    # pickle.dump fit's result (tests/test_model.py:41-43, sims/cb/cb.py:164)
    def __getstate__(self):
        return {'vp': self.vp.detach().cpu().clone(), 'size': self.size, 'support': self.support}

    def __setstate__(self, state):
        self.vp = torch.nn.Parameter(state['vp'])
        self.size, self.support = state['size'], state['support']

    # -- q(real) = Normal(loc, scale); reference model_param.py:46-52
    def loc_scale(self):
        if self.support in ('simplex', 'unit_interval'):
            return torch.sigmoid(self.vp[0]) * 20 - 10, torch.sigmoid(self.vp[1]) * 10
        return self.vp[0], torch.exp(self.vp[1])

    def dist(self):
        loc, scale = self.loc_scale()
        return Normal(loc, scale)

    def real_sample(self, n=torch.Size([])):
        return self.dist().rsample(n)

    def transform(self, real):
        if self.support == 'unit_interval':
            return torch.sigmoid(real)
        if self.support == 'simplex':
            return stick_breaking(real)
        if self.support == 'positive':
            return torch.exp(real)
        return real

    def logabsdetJ(self, real, param):
        if self.support == 'unit_interval':
            return torch.log(param) + torch.log1p(-param)
        if self.support == 'simplex':
            return stick_breaking_logdet(real, param)
        if self.support == 'positive':
            return real
        return 0.0

    def log_q(self, real):
        loc, scale = self.loc_scale()
        zed = (real - loc) / scale
        return (-0.5 * zed * zed - torch.log(scale) - HALF_LOG_2PI).sum()


def _rebuild_vae(mean, log_sd):
    v = VAE(mean.shape[1])
    with torch.no_grad():
        v.mean.copy_(mean)
        v.log_sd.copy_(log_sd)
    return v


class VAE(torch.nn.Module):
    """Imputation parameters of one sample: Normal(mean_j, exp(log_sd_j)) per
    marker (reference vae.py:7-15; despite the name it is not a network).

    Training never calls `forward`: the fused CUDA kernel imputes, evaluates
    log_qy and back-propagates into `mean` / `log_sd` in one pass.  `forward` is
    kept for API compatibility (reference vae.py:17-41) as a small elementwise
    torch expression on the caller's tensors."""

    def __init__(self, J, mean_init=-3, sd_init=.1, device=None):
        super().__init__()
        self.mean = torch.nn.Parameter(torch.full((1, J), float(mean_init), dtype=F64, device=device))
        self.log_sd = torch.nn.Parameter(torch.full((1, J), math.log(sd_init), dtype=F64, device=device))

    def to_cpu_copy(self):
        new = VAE(self.mean.shape[1])
        with torch.no_grad():
            new.mean.copy_(self.mean.detach().cpu())
            new.log_sd.copy_(self.log_sd.detach().cpu())
        return new

    def __reduce__(self):       # pickle as a CPU copy of the two rows (see ModelParam.__getstate__)
        return (_rebuild_vae, (self.mean.detach().cpu().clone(), self.log_sd.detach().cpu().clone()))

    def forward(self, y_mini, m_mini, N_full, scale=1):
        from torch.distributions import normal as _tdn
        miss = m_mini.to(device=self.mean.device)
        y0 = torch.where(miss, torch.zeros((), dtype=F64, device=miss.device),
                         y_mini.to(device=miss.device, dtype=F64))
        eps = _tdn._standard_normal(y0.shape, dtype=F64, device=miss.device)
        y_imp = torch.where(miss, self.mean + torch.exp(self.log_sd) * eps, y0)
        lq = torch.where(miss, -0.5 * eps * eps - self.log_sd - HALF_LOG_2PI,
                         torch.zeros((), dtype=F64, device=miss.device)).sum()
        return y_imp, lq * (N_full / y0.shape[0]) * scale
