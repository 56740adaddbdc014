"""Build a cytofpy_b200 Model in the state of a golden fixture and replay the
fixture's noise through the same torch hook the reference's draws go through."""
import contextlib

import numpy as np
import torch
import torch.distributions.normal as tdn

import cytofpy_b200 as cy

F64 = torch.float64


@contextlib.contextmanager
def replay_noise(tensors, target):
    """FIFO patch of torch.distributions.normal._standard_normal."""
    queue = list(tensors)
    orig = tdn._standard_normal

    def fake(shape, dtype=None, device=None, **kw):
        t = queue.pop(0)
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t.to(device=target, dtype=dtype)
    tdn._standard_normal = fake
    try:
        yield queue
    finally:
        tdn._standard_normal = orig


def apply_prior_spec(pri, spec):
    """Replace entries of a priors dict the way reference sims/cb/cb.py:117-123 does; `spec` is the
    {key: [family, *hyper-parameters]} record of a golden fixture (oracle/gen_golden.py)."""
    import torch.distributions as D
    for k, (fam, *args) in spec.items():
        a = [torch.tensor(float(v), dtype=F64) for v in args]
        if fam == 'Dirichlet':
            n = pri[k].concentration.numel()
            pri[k] = D.Dirichlet(a[0] * torch.ones(n, dtype=F64) / n)
        else:
            pri[k] = getattr(D, fam)(*a)
    return pri


def model_with_engine(factory):
    """Test-only Model subclass whose per-cell engine is `factory` (e.g. tests/_stub_engine.OracleEngine,
    the fp64 oracle on the CPU) -- the product class itself has no such seam."""
    class StubbedModel(cy.model.Model):
        def _make_engine(self, *args, **kwargs):
            return factory(*args, **kwargs)
    return StubbedModel


def build_model(case, device, engine_factory=None, noise='external'):
    y = [t.clone() for t in case.y]
    pri = cy.model.default_priors(y, K=case.K, L=[case.L0, case.L1], y_bounds=[-5., -3.5, -2.])
    apply_prior_spec(pri, getattr(case, 'prior_spec', {}))
    cls = cy.model.Model if engine_factory is None else model_with_engine(engine_factory)
    mod = cls(y=y, priors=pri, tau=case.tau, verbose=-1, use_stick_break=case.stick,
              model_noisy=case.noisy, scale=case.scale, device=device, noise=noise)
    assert list(mod.mp.keys()) == case.keys
    with torch.no_grad():
        for k in case.keys:
            mod.mp[k].vp.copy_(torch.tensor(case.z['vp_' + k], dtype=F64))
        for i in range(case.I):
            mod.y_vae[i].mean.copy_(torch.tensor(case.z['vae_mean_%d' % i], dtype=F64))
            mod.y_vae[i].log_sd.copy_(torch.tensor(case.z['vae_log_sd_%d' % i], dtype=F64))
    return mod


def noise_queue(case, t=None):
    nz = case.noise(t)
    return [nz[k] for k in case.keys] + list(nz['y'])


def model_grads(mod, case):
    g = {k: mod.mp[k].vp.grad.detach().cpu().numpy() for k in case.keys}
    for i in range(case.I):
        g['vae_mean_%d' % i] = mod.y_vae[i].mean.grad.detach().cpu().numpy()
        g['vae_log_sd_%d' % i] = mod.y_vae[i].log_sd.grad.detach().cpu().numpy()
    return g
