"""TEST INFRASTRUCTURE — generates tests/golden/*.npz from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

For every case it builds the reference `cytofpy.model.Model`, optionally trains
it for a few hundred reference `update` steps (to reach mixed quiet/noisy
responsibilities), then runs ONE `Model.forward` + `elbo.backward()` with all
Normal.rsample noise replaced by recorded tensors (the FIFO patch of
torch.distributions.normal._standard_normal described in SURVEY.md section 8c) and
records inputs (data, idx, variational state, noise) and outputs (elbo, ll,
lp, lq, log_qy, every .grad).  The "traj" case records three reference
`update` steps (Adam) instead.

The only shim applied to the reference is
torch.distributions.Distribution.set_default_validate_args(False), needed on
torch >= 1.8 because reference vae.py:29,33 builds Normal(., 0).
"""
import copy
import json
import os
import sys

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def _import_reference():
    sys.path.insert(0, REF)
    torch.distributions.Distribution.set_default_validate_args(False)
    import cytofpy  # noqa: flips torch default dtype to float64 (reference Model.py:16)
    return cytofpy


class NoiseTape:
    """Replaces torch.distributions.normal._standard_normal by a recorder /
    replayer so that the draws of one Model.forward are known."""

    def __init__(self, gen):
        self.gen = gen
        self.tape = []

    def __enter__(self):
        import torch.distributions.normal as tdn
        self._mod = tdn
        self._orig = tdn._standard_normal

        def fake(shape, dtype, device):
            t = torch.randn(tuple(shape), dtype=dtype, generator=self.gen)
            self.tape.append(t.clone())
            return t
        tdn._standard_normal = fake
        return self

    def __exit__(self, *a):
        self._mod._standard_normal = self._orig


def make_data(cytofpy, N, J, L0, L1, nan_frac_neg, seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    sim = cytofpy.util.simdata(N=N, J=J, a_W=[100.] * 5, L0=L0, L1=L1)
    y = copy.deepcopy(sim['data']['y'])
    if nan_frac_neg > 0:
        g = torch.Generator().manual_seed(seed + 1000)
        for yi in y:
            drop = (yi < 0) & (torch.rand(yi.shape, generator=g) < nan_frac_neg)
            yi[drop] = float('nan')
    return y


def apply_prior_spec(pri, spec):
    """spec: {key: [family, *hyper-parameters]} -> torch distributions put into the reference's
    priors dict the way reference sims/cb/cb.py:117-123 does.  A Dirichlet takes ONE number c and
    means Dirichlet(c * ones(n) / n) with n the size of the default entry."""
    import torch.distributions as D
    for k, (fam, *args) in spec.items():
        a = [torch.tensor(float(v), dtype=torch.float64) for v in args]
        if fam == 'Dirichlet':
            n = pri[k].concentration.numel()
            pri[k] = D.Dirichlet(a[0] * torch.ones(n, dtype=torch.float64) / n)
        elif fam == 'LogNormal':
            pri[k] = D.LogNormal(*a)
        else:
            pri[k] = getattr(D, fam)(*a)
    return pri


def run_case(cytofpy, name, N, J, K, L0, L1, mb, tau, stick, noisy, scale,
             nan_frac, train_steps=0, seed=0, traj=0, priors=None, train_lr=5e-2):
    y = make_data(cytofpy, N, J, L0, L1, nan_frac, seed)
    pri = cytofpy.model.default_priors(y, K=K, L=[L0, L1], y_bounds=[-5., -3.5, -2.])
    if priors:
        apply_prior_spec(pri, priors)
    torch.manual_seed(seed + 1)
    np.random.seed(seed + 1)
    mod = cytofpy.model.Model(y=y, priors=pri, tau=tau, verbose=-1, use_stick_break=stick,
                              model_noisy=noisy, scale=scale)
    rng = np.random.RandomState(seed + 2)

    def draw_idx():
        return [np.arange(n) if mb >= n else rng.choice(n, mb, replace=False) for n in N]

    if train_steps:
        opt = torch.optim.Adam(mod.parameters(), lr=train_lr)
        for _ in range(train_steps):
            cytofpy.model.update(opt, mod, draw_idx())

    rec = {'meta_N': np.array(N), 'meta_J': J, 'meta_K': K, 'meta_L0': L0, 'meta_L1': L1,
           'meta_tau': tau, 'meta_stick': int(stick), 'meta_noisy': int(noisy),
           'meta_scale': scale, 'meta_traj': traj, 'meta_priors': json.dumps(priors or {})}
    for i, yi in enumerate(y):
        rec['y_%d' % i] = yi.numpy()
    keys = list(mod.mp.keys())
    rec['meta_keys'] = np.array(keys)
    for k in keys:
        rec['vp_' + k] = mod.mp[k].vp.detach().numpy().copy()
    for i in range(len(N)):
        rec['vae_mean_%d' % i] = mod.y_vae[i].mean.detach().numpy().copy()
        rec['vae_log_sd_%d' % i] = mod.y_vae[i].log_sd.detach().numpy().copy()

    gen = torch.Generator().manual_seed(seed + 3)
    if not traj:
        idx = draw_idx()
        with NoiseTape(gen) as tape:
            elbo, ll, lp, lq = mod(idx)
        mod.zero_grad()
        elbo.backward()
        assert len(tape.tape) == len(keys) + len(N)
        for k, t in zip(keys, tape.tape[:len(keys)]):
            assert tuple(t.shape) == tuple(mod.mp[k].vp[0].shape)
            rec['noise_' + k] = t.numpy()
        for i in range(len(N)):
            rec['idx_%d' % i] = np.asarray(idx[i], dtype=np.int64)
            rec['noise_y_%d' % i] = tape.tape[len(keys) + i].numpy()
        rec['out_elbo'] = float(elbo.item())
        rec['out_ll'], rec['out_lp'], rec['out_lq'] = ll, lp, lq
        rec['out_log_qy'] = float(mod.log_qy)
        for k in keys:
            rec['grad_' + k] = mod.mp[k].vp.grad.numpy().copy()
        for i in range(len(N)):
            rec['grad_vae_mean_%d' % i] = mod.y_vae[i].mean.grad.numpy().copy()
            rec['grad_vae_log_sd_%d' % i] = mod.y_vae[i].log_sd.grad.numpy().copy()
    else:
        lr = 0.1
        rec['meta_lr'] = lr
        opt = torch.optim.Adam(mod.parameters(), lr=lr)
        elbos = []
        for t in range(traj):
            idx = draw_idx()
            with NoiseTape(gen) as tape:
                elbo, fixed, ll, lp, lq = cytofpy.model.update(opt, mod, idx)
            elbos.append(float(elbo.item()))
            for k, tt in zip(keys, tape.tape[:len(keys)]):
                rec['noise_%d_%s' % (t, k)] = tt.numpy()
            for i in range(len(N)):
                rec['idx_%d_%d' % (t, i)] = np.asarray(idx[i], dtype=np.int64)
                rec['noise_y_%d_%d' % (t, i)] = tape.tape[len(keys) + i].numpy()
        rec['out_elbos'] = np.array(elbos)
        for k in keys:
            rec['final_vp_' + k] = mod.mp[k].vp.detach().numpy().copy()
        for i in range(len(N)):
            rec['final_vae_mean_%d' % i] = mod.y_vae[i].mean.detach().numpy().copy()
            rec['final_vae_log_sd_%d' % i] = mod.y_vae[i].log_sd.detach().numpy().copy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print('wrote', name, 'elbo' if not traj else 'elbos', rec.get('out_elbo', rec.get('out_elbos')))


CASES = [
    # name, N, J, K, L0, L1, mb, tau, stick, noisy, scale, nan_frac, train_steps, seed, traj
    dict(name='a_reftest_fresh', N=[300, 100, 200], J=8, K=10, L0=3, L1=3, mb=100, tau=.1,
         stick=True, noisy=True, scale=1.0, nan_frac=0.0),
    dict(name='b_reftest_trained', N=[300, 100, 200], J=8, K=10, L0=3, L1=3, mb=100, tau=.1,
         stick=True, noisy=True, scale=1.0, nan_frac=0.0, train_steps=150),
    dict(name='c_nan_nostick_tau001', N=[300, 100, 200], J=8, K=10, L0=3, L1=3, mb=100, tau=.001,
         stick=False, noisy=True, scale=0.7, nan_frac=0.3, train_steps=60),
    dict(name='d_cfg3like_fresh', N=[96, 64, 80], J=32, K=30, L0=5, L1=3, mb=48, tau=.001,
         stick=False, noisy=True, scale=0.1, nan_frac=0.3),
    dict(name='e_cfg2like_quiet', N=[96, 64, 80], J=32, K=30, L0=5, L1=3, mb=48, tau=.1,
         stick=True, noisy=False, scale=1.0, nan_frac=0.3),
    dict(name='f_fullbatch_mixed', N=[150, 70], J=8, K=6, L0=4, L1=3, mb=100000, tau=.1,
         stick=True, noisy=True, scale=0.7, nan_frac=0.3, train_steps=200),
    dict(name='g_cfg5like', N=[40, 30], J=40, K=50, L0=8, L1=8, mb=24, tau=.1,
         stick=True, noisy=True, scale=1.0, nan_frac=0.2),
    dict(name='h_quiet_trained', N=[150, 70], J=8, K=6, L0=4, L1=3, mb=50, tau=.1,
         stick=False, noisy=False, scale=1.0, nan_frac=0.3, train_steps=100),
    dict(name='i_cfg4like_trained', N=[64, 64, 64, 64], J=32, K=30, L0=5, L1=5, mb=100000, tau=.1,
         stick=True, noisy=True, scale=1.0, nan_frac=0.0, train_steps=40),
    # the cord-blood prior set of reference sims/cb/cb.py:115-123 (sig2 ~ Gamma(.1, 1)) at cfg3 shapes,
    # trained to a mixed quiet / noisy state; gathered minibatch, NaN, no stick-breaking, tau = .001
    dict(name='j_cb_priors_cfg3', N=[96, 64, 80], J=32, K=30, L0=5, L1=3, mb=48, tau=.001,
         stick=False, noisy=True, scale=0.1, nan_frac=0.3, train_steps=1500, seed=4, train_lr=0.05,
         priors={'sig2': ['Gamma', .1, 1.], 'alpha': ['Gamma', .1, .1], 'delta0': ['Gamma', 1., 1.],
                 'delta1': ['Gamma', 1., 1.], 'eps': ['Beta', 1., 99.]}),
    # non-default hyper-parameters of every recognised family at cfg1 shapes (K = 5, L = 3/3, full batch)
    dict(name='k_priors_cfg1', N=[120, 80, 100], J=32, K=5, L0=3, L1=3, mb=100000, tau=.1,
         stick=True, noisy=True, scale=1.0, nan_frac=0.2, train_steps=400, seed=5,
         priors={'sig2': ['Gamma', .5, 2.], 'alpha': ['Gamma', 2., .1], 'delta0': ['Gamma', 2., .5],
                 'delta1': ['Gamma', 1.5, 2.], 'eps': ['Beta', 2., 50.], 'eta0': ['Dirichlet', 2.1],
                 'eta1': ['Dirichlet', .6], 'W': ['Dirichlet', 3.5]}),
    # trained mixed state at cfg2 shapes (J = 32, K = 30, L = 5/3) with the LogNormal sig2 prior moved
    dict(name='l_cfg2_trained_mixed', N=[96, 64, 80], J=32, K=30, L0=5, L1=3, mb=48, tau=.1,
         stick=True, noisy=True, scale=1.0, nan_frac=0.0, train_steps=150, seed=6,
         priors={'sig2': ['LogNormal', -.5, .3]}),
    dict(name='t_traj3', N=[300, 100, 200], J=8, K=10, L0=3, L1=3, mb=100, tau=.1,
         stick=True, noisy=True, scale=1.0, nan_frac=0.3, traj=3),
]


def main():
    cytofpy = _import_reference()
    only = sys.argv[1:]
    for c in CASES:
        if not only or c['name'] in only:
            run_case(cytofpy, **c)


if __name__ == '__main__':
    main()
