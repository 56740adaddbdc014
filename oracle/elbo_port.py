"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

fp64 CPU restatement ("port") of CytofPy's ADVI ELBO path, used only as
 * the parity checker in tests/ and __graft_entry__.smoke(), and
 * the timed CPU arm of bench.py (`cpu_baseline`, `--impl reference`).
Nothing under cytofpy_b200/ may import this module.

Pinning: the reference's own tests hold no golden vector for this path
(reference tests/test_model.py:14 asserts only 1+1==2), so the pin is the live
reference itself: oracle/gen_golden.py imports /root/reference in the build
container, runs Model.forward + backward under injected noise, and commits the
inputs/outputs under tests/golden/*.npz.  tests/test_oracle_golden.py checks
this port against every one of those fixtures (<= 1e-12 relative), and
`solve_beta` against the one known answer printed in reference
cytofpy/model/Model.py:47-49.

Reference lines followed (all paths relative to /root/reference):
  variational family / sampling    cytofpy/model/model_param.py:46-55
  constraining transforms           cytofpy/model/model_param.py:57-71
  log|det J|                        cytofpy/model/model_param.py:73-89
  entropy term log q                cytofpy/model/model_param.py:91-92, Model.py:280-287
  imputation ("VAE")                cytofpy/model/vae.py:17-41
  compute_Z (straight-through)      cytofpy/model/Model.py:22-34
  missingness probability           cytofpy/model/Model.py:36-38
  loglike                           cytofpy/model/Model.py:210-278
  log_prior                         cytofpy/model/Model.py:289-304
  priors                            cytofpy/model/Model.py:60-112
  forward / update                  cytofpy/model/Model.py:339-346, fit.py:32-48

The primitive arithmetic (Normal.log_prob, StickBreakingTransform, prior
log_probs, logsumexp, Adam) comes from the same third-party dependency the
reference calls: torch (un-pinned in reference setup.py:13; 2.11.0 here).

The port is written as pure functions over plain dicts (no nn.Module, no
ModelParam class), but it deliberately keeps the reference's *computational
shape* — dense (B,J,L) and (B,J,K) fp64 intermediates and torch autograd for
the backward — so that timing it is a fair stand-in for the reference's CPU
path when /root/reference is not on the machine.
"""
import math

import numpy as np
import torch
from torch.distributions import Beta, Dirichlet, Gamma, Normal, Uniform
from torch.distributions.log_normal import LogNormal
from torch.distributions.transforms import StickBreakingTransform

F64 = torch.float64
_STICK = StickBreakingTransform(0)

# support of every global block, reference Model.py:169-192
SUPPORT = {
    'delta0': 'positive', 'delta1': 'positive', 'eps': 'unit_interval',
    'sig2': 'positive', 'eta0': 'simplex', 'eta1': 'simplex', 'W': 'simplex',
    'alpha': 'positive', 'v': 'unit_interval', 'H': 'unit_interval',
}


def key_order(model_noisy=True):
    """Insertion order of Model.mp (reference Model.py:169-192); also the order
    in which Normal.rsample consumes noise in Model.sample_reals (:306-310)."""
    keys = ['delta0', 'delta1']
    if model_noisy:
        keys.append('eps')
    return keys + ['sig2', 'eta0', 'eta1', 'W', 'alpha', 'v', 'H']


def block_shapes(I, J, K, L0, L1, model_noisy=True):
    """Shape of vp[0] for every global block (reference Model.py:169-192)."""
    s = {'delta0': (L0,), 'delta1': (L1,), 'sig2': (I,),
         'eta0': (I, J, L0 - 1), 'eta1': (I, J, L1 - 1), 'W': (I, K - 1),
         'alpha': (1,), 'v': (K,), 'H': (J, K)}
    if model_noisy:
        s['eps'] = (I,)
    return {k: s[k] for k in key_order(model_noisy)}


def solve_beta(y_bounds, p_bounds):
    """Quadratic-logit missingness coefficients through three (y, p) points
    (reference Model.py:40-45)."""
    yb = np.asarray(y_bounds, dtype=np.float64)
    p = np.asarray(p_bounds, dtype=np.float64)
    design = np.stack([np.ones_like(yb), yb, yb * yb], axis=1)
    return np.linalg.solve(design, np.log(p) - np.log1p(-p))


def prior_table(K, L0, L1, spec=None):
    """Priors of reference default_priors (Model.py:90-112), minus the sizes; `spec`
    ({key: [family, *hyper-parameters]}, the notation of oracle/gen_golden.py) replaces entries
    the way a caller of the reference does (reference sims/cb/cb.py:117-123)."""
    one = lambda n: torch.ones(n, dtype=F64)
    tab = {
        'delta0': Gamma(torch.tensor(1., dtype=F64), torch.tensor(1., dtype=F64)),
        'delta1': Gamma(torch.tensor(1., dtype=F64), torch.tensor(1., dtype=F64)),
        'sig2': LogNormal(torch.tensor(-1., dtype=F64), torch.tensor(.1, dtype=F64)),
        'eta0': Dirichlet(one(L0) / L0),
        'eta1': Dirichlet(one(L1) / L1),
        'alpha': Gamma(torch.tensor(.1, dtype=F64), torch.tensor(.1, dtype=F64)),
        'H': Uniform(torch.tensor(0., dtype=F64), torch.tensor(1., dtype=F64)),
        'W': Dirichlet(one(K) / K),
        'eps': Beta(torch.tensor(1., dtype=F64), torch.tensor(99., dtype=F64)),
    }
    fam = {'Gamma': Gamma, 'Beta': Beta, 'LogNormal': LogNormal, 'Uniform': Uniform}
    for k, (name, *args) in (spec or {}).items():
        a = [torch.tensor(float(v), dtype=F64) for v in args]
        if name == 'Dirichlet':       # one number c: Dirichlet(c * ones(n) / n)
            n = tab[k].concentration.numel()
            tab[k] = Dirichlet(a[0] * one(n) / n)
        else:
            tab[k] = fam[name](*a)
    return tab


def init_state(I, J, K, L0, L1, model_noisy=True, use_stick_break=True,
               y_mean_init=-3.0, y_sd_init=0.1, generator=None):
    """Initial variational parameters, reference Model.py:169-208 and
    model_param.py:31-42 (random N(0,1) where the reference passes no m / s)."""
    def rn(shape):
        return torch.randn(shape, dtype=F64, generator=generator)

    def block(shape, m=None, s=None):
        m = rn(shape) if m is None else torch.as_tensor(m, dtype=F64).expand(shape).clone()
        ls = rn(shape) if s is None else torch.as_tensor(s, dtype=F64).expand(shape).clone().log()
        return torch.stack([m, ls]).requires_grad_(True)

    shp = block_shapes(I, J, K, L0, L1, model_noisy)
    st = {}
    st['delta0'] = block(shp['delta0'], 1.0, 1.0)
    st['delta1'] = block(shp['delta1'], 1.0, 1.0)
    if model_noisy:
        st['eps'] = block(shp['eps'], 1.0 / 100.0, .001)   # Beta(1,99).mean
    st['sig2'] = block(shp['sig2'], -1.0, .1)
    st['eta0'] = block(shp['eta0'])
    st['eta1'] = block(shp['eta1'])
    st['W'] = block(shp['W'])
    st['alpha'] = block(shp['alpha'])
    st['v'] = block(shp['v'], .99 if use_stick_break else None)
    st['H'] = block(shp['H'])
    st = {k: st[k] for k in key_order(model_noisy)}
    st['vae_mean'] = [torch.full((1, J), float(y_mean_init), dtype=F64, requires_grad=True)
                      for _ in range(I)]
    st['vae_log_sd'] = [torch.full((1, J), math.log(y_sd_init), dtype=F64, requires_grad=True)
                        for _ in range(I)]
    return st


def state_parameters(state, model_noisy=True):
    """Parameter list in the reference's registration order
    (model_param.py:116-119 then Model.py:202-208)."""
    out = [state[k] for k in key_order(model_noisy)]
    for m, s in zip(state['vae_mean'], state['vae_log_sd']):
        out += [m, s]
    return out


def loc_scale(vp, support):
    """Mean-field Normal over the unconstrained real, model_param.py:46-52."""
    if support in ('simplex', 'unit_interval'):
        return torch.sigmoid(vp[0]) * 20 - 10, torch.sigmoid(vp[1]) * 10
    return vp[0], torch.exp(vp[1])


def constrain(real, support):
    """model_param.py:57-71"""
    if support == 'unit_interval':
        return torch.sigmoid(real)
    if support == 'simplex':
        return _STICK(real)
    if support == 'positive':
        return torch.exp(real)
    return real


def log_abs_det_jac(real, param, support):
    """model_param.py:73-89"""
    if support == 'unit_interval':
        return torch.log(param) + torch.log1p(-param)
    if support == 'simplex':
        return _STICK.log_abs_det_jacobian(real, param)
    if support == 'positive':
        return real
    return 0.0


def straight_through_Z(v, H, tau):
    """Model.py:22-34: forward value 1[v>H], gradient of sigmoid((logit v - logit H)/tau)."""
    lg = lambda p: torch.log(p) - torch.log1p(-p)
    soft = torch.sigmoid((lg(v[None, :]) - lg(H)) / tau)
    hard = (v > H).to(F64)
    return (hard - soft).detach() + soft


def impute_rows(y_rows, miss, mean, log_sd, eps, N_full, scale):
    """vae.py:17-41 with the (B,J) standard-normal draw `eps` supplied."""
    y0 = torch.where(miss, torch.zeros_like(y_rows), y_rows)
    mf = miss.to(F64)
    loc = y0 * (1 - mf) + mean * mf
    sd = torch.exp(log_sd) * mf
    y_imp = loc + eps * sd
    lq = Normal(loc[miss], sd[miss], validate_args=False).log_prob(y_imp[miss]).sum()
    return y_imp, lq * (N_full / y_rows.shape[0]) * scale


def loglike_port(params, y_imp, miss, cfg):
    """Model.py:210-278, one pass per sample with the reference's dense
    (B,J,L) and (B,J,K) intermediates."""
    sig = torch.sqrt(params['sig2'])
    mu0 = -torch.cumsum(params['delta0'], 0)
    mu1 = torch.cumsum(params['delta1'], 0)
    v = torch.cumprod(params['v'], 0) if cfg['use_stick_break'] else params['v']
    total = 0.0
    for i in range(cfg['I']):
        yi = y_imp[i]
        lm0 = torch.logsumexp(
            Normal(mu0[None, None, :], sig[i]).log_prob(yi[:, :, None]) + torch.log(params['eta0'][i:i + 1]), 2)
        lm1 = torch.logsumexp(
            Normal(mu1[None, None, :], sig[i]).log_prob(yi[:, :, None]) + torch.log(params['eta1'][i:i + 1]), 2)
        Z = straight_through_Z(v, params['H'], cfg['tau'])
        gated = Z[None] * lm1[:, :, None] + (1 - Z[None]) * lm0[:, :, None]
        f = gated.sum(1) + torch.log(params['W'][i:i + 1])
        cell = torch.logsumexp(f, 1)
        if cfg['model_noisy']:
            e = params['eps'][i]
            quiet = cell + torch.log1p(-e)
            noisy = Normal(torch.tensor(0., dtype=F64), cfg['noisy_sd']).log_prob(yi).sum(1) + torch.log(e)
            s_i = torch.stack([quiet, noisy]).logsumexp(0).sum(0)
        else:
            s_i = cell.sum(0)
        if bool(miss[i].sum() > 0):
            u = cfg['b0'][i] + cfg['b1'][i] * yi + cfg['b2'][i] * yi ** 2
            s_i = s_i + torch.log(torch.sigmoid(u)[miss[i]]).sum()
        total = total + s_i * (cfg['N'][i] / yi.shape[0])
    return total


def elbo_port(state, y, idx, noise, cfg):
    """Model.forward (Model.py:339-346) under supplied noise.

    state : init_state()-style dict          y   : list of (N_i,J) fp64, NaN = missing
    idx   : list of I integer index arrays   cfg : make_cfg() dict
    noise : {'<key>': tensor like vp[0], ..., 'y': [ (B_i,J) ... ]}
    returns dict(elbo, ll, lp, lq, log_qy) of 0-d fp64 tensors (graph attached)
    and 'y_imp' (list of imputed minibatches)."""
    keys = key_order(cfg['model_noisy'])
    pri = prior_table(cfg['K'], cfg['L0'], cfg['L1'], cfg.get('prior_spec'))
    reals, params = {}, {}
    lq = 0.0
    for k in keys:
        loc, sc = loc_scale(state[k], SUPPORT[k])
        reals[k] = loc + noise[k] * sc
        params[k] = constrain(reals[k], SUPPORT[k])
        lq = lq + Normal(loc, sc).log_prob(reals[k]).sum()
    y_imp, miss, log_qy = [], [], 0.0
    for i in range(cfg['I']):
        rows = y[i][torch.as_tensor(np.asarray(idx[i]), dtype=torch.long)]
        mi = torch.isnan(rows)
        yi, lqi = impute_rows(rows, mi, state['vae_mean'][i], state['vae_log_sd'][i],
                              noise['y'][i], cfg['N'][i], cfg['scale'])
        y_imp.append(yi)
        miss.append(mi)
        log_qy = log_qy + lqi
    lq = lq + log_qy
    ll = loglike_port(params, y_imp, miss, cfg)
    lp = 0.0
    for k in keys:
        if k == 'v':
            conc = params['alpha'] if cfg['use_stick_break'] else params['alpha'] / cfg['K']
            t = Beta(conc, torch.tensor(1., dtype=F64), validate_args=False).log_prob(params['v'])
        else:
            t = pri[k].log_prob(params[k])
        lp = lp + (t + log_abs_det_jac(reals[k], params[k], SUPPORT[k])).sum()
    return {'elbo': ll + lp - lq, 'll': ll, 'lp': lp, 'lq': lq,
            'log_qy': log_qy, 'y_imp': y_imp, 'params': params}


def make_cfg(y, K, L0, L1, tau=0.1, use_stick_break=True, model_noisy=True, scale=1.0,
             y_bounds=(-5., -3.5, -2.), p_bounds=(.05, .8, .05), noisy_var=10.0, prior_spec=None):
    I = len(y)
    beta = solve_beta(y_bounds, p_bounds)
    return {
        'I': I, 'J': int(y[0].shape[1]), 'K': int(K), 'L0': int(L0), 'L1': int(L1),
        'N': [int(t.shape[0]) for t in y], 'tau': float(tau),
        'use_stick_break': bool(use_stick_break), 'model_noisy': bool(model_noisy),
        'scale': float(scale), 'noisy_sd': torch.sqrt(torch.tensor(noisy_var, dtype=F64)),
        'b0': torch.full((I,), beta[0], dtype=F64), 'b1': torch.full((I,), beta[1], dtype=F64),
        'b2': torch.full((I,), beta[2], dtype=F64),
        'prior_spec': dict(prior_spec or {}),
    }


def draw_noise(cfg, idx, generator=None):
    shp = block_shapes(cfg['I'], cfg['J'], cfg['K'], cfg['L0'], cfg['L1'], cfg['model_noisy'])
    nz = {k: torch.randn(s, dtype=F64, generator=generator) for k, s in shp.items()}
    nz['y'] = [torch.randn((len(ix), cfg['J']), dtype=F64, generator=generator) for ix in idx]
    return nz


def update_port(opt, state, y, idx, noise, cfg):
    """fit.py:32-48: loss = -elbo/Nsum, backward, NaN scrub of the 10 global
    blocks only, optimizer step."""
    out = elbo_port(state, y, idx, noise, cfg)
    loss = -out['elbo'] / sum(cfg['N'])
    opt.zero_grad()
    loss.backward()
    fixed = False
    with torch.no_grad():
        for k in key_order(cfg['model_noisy']):
            bad = torch.isnan(state[k].grad)
            if bool(bad.any()):
                state[k].grad[bad] = 0.0
                fixed = True
    opt.step()
    return out, fixed
