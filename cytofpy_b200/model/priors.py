"""Priors and small helpers of the model; API mirror of reference
cytofpy/model/Model.py:18-112 (logit, compute_Z, prob_miss, solve_beta,
gen_beta_est, default_priors)."""
import numpy as np
import torch
from torch.distributions import Beta, Dirichlet, Gamma, Uniform
from torch.distributions.log_normal import LogNormal

F64 = torch.float64


def logit(p):
    return torch.log(p) - torch.log1p(-p)


def compute_Z(v, H, tau):
    """Binary feature matrix 1[v_k > H_jk] in the forward pass with the
    gradient of the tempered sigmoid (straight-through), reference
    Model.py:22-34.  The threshold is evaluated here in fp64; the CUDA kernel
    only ever sees the resulting bit mask."""
    soft = torch.sigmoid((logit(v[None, :]) - logit(H)) / tau)
    hard = (v > H).to(F64)
    return (hard - soft).detach() + soft


def prob_miss(y, b0, b1, b2):
    """P(missing | y), reference Model.py:36-38."""
    return torch.sigmoid(b0 + b1 * y + b2 * y ** 2)


def solve_beta(y, p):
    """Coefficients of the quadratic logit through the (y, p) points,
    reference Model.py:40-45 (known answer Model.py:47-49)."""
    y = np.asarray(y, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    design = np.stack([np.ones_like(y), y, y * y], axis=1)
    return torch.tensor(np.linalg.solve(design, np.log(p) - np.log1p(-p)), dtype=F64)


def gen_beta_est(yi, y_quantiles, p_bounds):
    """reference Model.py:51-54: anchor points = percentiles of the negative,
    observed expression values (NaN < 0 is False, so missing entries drop out)."""
    yi = torch.as_tensor(yi)
    neg = yi[yi < 0]
    return solve_beta(np.percentile(neg.cpu().numpy(), y_quantiles), p_bounds)


def default_priors(y, K: int = 30, L=None, y_quantiles=[0, 35, 70], p_bounds=[.05, .8, .05],
                   y_bounds=None):
    """reference Model.py:60-112 (same keys, same distributions)."""
    I = len(y)
    J = y[0].size(1)
    for yi in y:
        assert yi.size(1) == J
    N = [yi.size(0) for yi in y]
    if L is None:
        L = [5, 5]
    b = torch.zeros((3, I), dtype=F64)
    for i in range(I):
        b[:, i] = gen_beta_est(y[i].flatten(), y_quantiles, p_bounds) if y_bounds is None \
            else solve_beta(np.array(y_bounds), p_bounds)
    one = lambda n: torch.ones(n, dtype=F64)
    t = lambda x: torch.tensor(float(x), dtype=F64)
    return {'I': I, 'J': J, 'N': N, 'L': L, 'K': K,
            'delta0': Gamma(t(1), t(1)), 'delta1': Gamma(t(1), t(1)),
            'sig2': LogNormal(t(-1), t(.1)),
            'eta0': Dirichlet(one(L[0]) / L[0]), 'eta1': Dirichlet(one(L[1]) / L[1]),
            'alpha': Gamma(t(.1), t(.1)), 'H': Uniform(t(0), t(1)),
            'b0': b[0].clone(), 'b1': b[1].clone(), 'b2': b[2].clone(),
            'noisy_var': 10.0,
            'W': Dirichlet(one(K) / K),
            'eps': Beta(t(1), t(99))}


def dist_to(d, device):
    """Re-instantiate a torch Distribution with fp64 parameters on `device`
    (callers replace entries of the priors dict by arbitrary distributions,
    e.g. reference sims/cb/cb.py:117-123)."""
    if isinstance(d, LogNormal):
        return LogNormal(d.loc.to(device=device, dtype=F64), d.scale.to(device=device, dtype=F64),
                         validate_args=False)
    kw = {k: torch.as_tensor(getattr(d, k)).to(device=device, dtype=F64) for k in d.arg_constraints}
    return type(d)(**kw, validate_args=False)
