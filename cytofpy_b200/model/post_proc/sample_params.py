"""One posterior draw of the model parameters in the form the analysis code wants;
API mirror of reference cytofpy/model/post_proc/sample_params.py:4-51."""
import torch


def theta(mod, priors, y=None):
    """Returns (theta dict, mod).  theta holds detached mu0, mu1, eta0, eta1, sig, W, Z, eps,
    noisy_sd and 'y' = list of (N_i, J) matrices: the imputed data of this draw (emit kernel)
    or the caller's `y`."""
    idx = [range(n) for n in mod.N_local]
    params = mod.sample_params(idx)
    det = lambda k: params[k].detach()
    v, H = det('v'), det('H')
    vz = torch.cumprod(v, 0) if mod.use_stick_break else v
    out = {
        'mu0': -torch.cumsum(det('delta0'), 0), 'mu1': torch.cumsum(det('delta1'), 0),
        'eta0': det('eta0'), 'eta1': det('eta1'), 'sig': torch.sqrt(det('sig2')),
        'W': det('W'), 'Z': (vz > H).to(v.dtype),
        'noisy_sd': torch.sqrt(torch.tensor(priors['noisy_var'], dtype=v.dtype)),
    }
    if mod.model_noisy:
        out['eps'] = det('eps')
    out['y'] = [params['y'][i].detach() if y is None else y[i] for i in range(mod.I)]
    return out, mod
