"""GPU: the callers' side of fit() -- what the reference's drivers do with its result (pickle it, rebuild a Model
from it: tests/test_model.py:41-43, sims/cb/cb.py:164,173-179) and the cord-blood file format feeding it
(cytofpy/util/readCB.py:3-40 -> fit)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

import cytofpy_b200 as cy
from tests.test_host_logic_cpu import _write_cb_file

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _small_fit(max_iter=25, **kw):
    torch.manual_seed(0)
    np.random.seed(0)
    y = cy.util.simdata_with_missing_values(N=[300, 100, 200], J=8, a_W=[100.] * 5, L0=3, L1=3, seed=0, rate=0.2)['data']['y']
    pri = cy.model.default_priors(y, K=10, L=[3, 3], y_bounds=[-5., -3.5, -2.])
    out = cy.model.fit(y, max_iter=max_iter, lr=1e-1, print_freq=1000, eps_conv=0, priors=pri, minibatch_size=100,
                       tau=0.1, trace_every=5, backup_every=5, verbose=0, seed=1, **kw)
    return y, pri, out


def test_fit_result_pickles_and_loads_without_a_gpu(tmp_path):
    y, pri, out = _small_fit()
    assert set(out) == {'elbo', 'trace', 'mp', 'tau', 'vae', 'use_stick_break', 'y', 'priors', 'model_noisy'}
    assert len(out['elbo']) == 26 and len(out['trace']) == 6
    assert all(not blk.vp.is_cuda for blk in out['mp'].values()) and all(not v.mean.is_cuda for v in out['vae'])
    path = str(tmp_path / 'out.p')
    pickle.dump(out, open(path, 'wb'))
    assert os.path.getsize(path) < 2_000_000
    code = r'''
import os, sys, pickle
os.environ['CUDA_VISIBLE_DEVICES'] = ''
sys.path.insert(0, %r)
import torch
out = pickle.load(open(%r, 'rb'))
assert not torch.cuda.is_available()
W = out['mp']['W'].dist().mean                      # sims/cb/cb.py:282-283 does this on trace entries
assert W.shape == (3, 9) and torch.isfinite(W).all()
assert torch.isfinite(out['trace'][-1]['H'].dist().mean).all()
yi, _ = out['vae'][0](out['y'][0], torch.isnan(out['y'][0]), out['y'][0].shape[0])     # cb.py:334
assert not torch.isnan(yi).any()
print('ok', float(W.sum()))
''' % (ROOT, path)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith('ok'), r.stdout + r.stderr[-2000:]


def test_model_restored_from_fit_result_like_the_reference_drivers():
    """`mod = Model(...); mod.mp = out['mp']; mod.y_vae = out['vae']` (sims/cb/cb.py:173-179), then posterior draws"""
    y, pri, out = _small_fit()
    out = pickle.loads(pickle.dumps(out))
    dev = torch.device('cuda', 0)
    mod = cy.model.Model(y=[t.clone() for t in y], priors=pri, tau=out['tau'], use_stick_break=out['use_stick_break'],
                         model_noisy=out['model_noisy'], verbose=-1, device=dev)
    mod.mp = out['mp']
    mod.y_vae = out['vae']
    for k in out['mp']:
        assert mod.mp[k].vp.is_cuda and torch.equal(mod.mp[k].vp.detach().cpu(), out['mp'][k].vp.detach())
    for a, b in zip(mod.y_vae, out['vae']):
        assert a.mean.is_cuda and torch.equal(a.mean.detach().cpu(), b.mean.detach())
    idx = [np.random.choice(n, 5) for n in mod.N]
    post = mod.sample_params(idx)
    assert post['W'].shape == (3, 10) and len(post['y']) == 3 and post['y'][0].shape == (5, 8)
    theta, _ = cy.model.post_proc.sample_params.theta(mod, pri)
    lam = cy.model.post_proc.lam_post.sample(theta, mod)
    assert [int(t.numel()) for t in lam] == mod.N and all(int(t.max()) <= 10 for t in lam)
    # the elbo of the restored state continues where the fit left off (same order of magnitude, finite)
    elbo, ll, lp, lq = mod([np.arange(n) for n in mod.N])
    assert np.isfinite(elbo.item()) and abs(elbo.item() - out['elbo'][-1]) < 0.5 * abs(out['elbo'][-1])


def test_cord_blood_file_to_fit(tmp_path):
    """f4: file -> readCB -> preprocess -> fit on the GPU (cord-blood shapes: 3 samples, 32 markers, NA entries,
    minibatch 500, no stick-breaking, tau = .001, scale = .1 as in sims/cb/cb.py:156-159)."""
    N, J = [14000, 12500, 13600], 32
    p = str(tmp_path / 'cb.txt')
    _write_cb_file(p, N, J, seed=3, na_rate=0.3)
    data = cy.util.readCB(p)
    cy.util.preprocess(data, rm_cells_below=-6.0)
    y = data['y']
    assert sum(data['N']) > 25000 and all(bool(torch.isnan(t).any()) for t in y)
    pri = cy.model.default_priors(y, K=30, L=[5, 3], y_quantiles=[0, 25, 50], p_bounds=[.05, .8, .05])
    pri['sig2'] = torch.distributions.Gamma(torch.tensor(.1, dtype=torch.float64), torch.tensor(1., dtype=torch.float64))
    out = cy.model.fit(y, max_iter=50, lr=1e-2, print_freq=1000, eps_conv=0, priors=pri, minibatch_size=500, tau=0.001,
                       trace_every=50, backup_every=10, verbose=0, seed=2, use_stick_break=False, scale=.1)
    e = np.array(out['elbo'])
    assert len(e) == 51 and np.isfinite(e).all()
    assert e[-10:].mean() > e[:10].mean()                    # 50 Adam steps improve the ELBO
    assert out['mp']['sig2'].vp.shape == (2, 3)


def test_out_of_range_minibatch_indices_raise_like_the_reference():
    y = [torch.randn(50, 8, dtype=torch.float64), torch.randn(40, 8, dtype=torch.float64)]
    pri = cy.model.default_priors(y, K=4, L=[3, 3], y_bounds=[-5., -3.5, -2.])
    mod = cy.model.Model(y=y, priors=pri, verbose=-1, device=torch.device('cuda', 0))
    with pytest.raises(IndexError):
        mod([np.arange(10), np.array([0, 40])])
    with pytest.raises(IndexError):
        mod([np.array([-51]), np.arange(5)])
    elbo, *_ = mod([np.array([-1, 0, 49]), np.arange(40)])          # negative indices wrap like numpy / torch
    assert np.isfinite(elbo.item())
