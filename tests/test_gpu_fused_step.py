"""GPU: the fully fused step (global part as two fp64 kernels + the per-cell kernel + Adam,
cytofpy_b200/model/fused.py) against the golden fixtures of the live reference."""
import numpy as np
import pytest
import torch

import cytofpy_b200 as cy
from cytofpy_b200.model.fused import FusedStep, recognise_priors
from tests._golden import Case, case_names, rel
from tests._model_util import build_model

pytestmark = pytest.mark.gpu
F64 = torch.float64
ELBO_TOL, GRAD_TOL = 1e-5, 1e-4


def _noise(c, dev, t=None):
    nz = c.noise(t)
    g = torch.cat([nz[k].reshape(-1) for k in c.keys]).to(dev)
    cell = torch.cat(nz['y'], 0).to(dev, torch.float32).contiguous()
    return g, cell


@pytest.mark.parametrize('name', case_names())
def test_fused_step_matches_reference_fixture(name):
    c = Case(name)
    dev = torch.device('cuda', 0)
    mod = build_model(c, dev)
    fs = FusedStep(mod, lr=0.1)
    assert fs.R == sum(int(np.prod(c.z['noise_' + k].shape)) for k in c.keys)
    g, cell = _noise(c, dev)
    theta0 = fs.theta.clone()
    elbo, ll, lp, lq, scrub = fs.step(c.idx(), global_noise=g, cell_noise=cell, update=False).tolist()
    assert torch.equal(fs.theta, theta0)                      # update=False leaves parameters alone
    assert abs(elbo - c.out('elbo')) <= ELBO_TOL * abs(c.out('elbo'))
    assert abs(ll - c.out('ll')) <= ELBO_TOL * abs(c.out('ll'))
    # lp contains log1p(-sigmoid(x)) terms (reference model_param.py:75) that amplify a 1-ulp
    # difference between the device's and the host's exp() by up to 1e9: 1e-7 is what fp64 allows
    assert abs(lp - c.out('lp')) <= 1e-7 * max(1, abs(c.out('lp')))
    assert abs(lq - c.out('lq')) <= ELBO_TOL * max(1, abs(c.out('lq')))
    assert scrub == 0.0
    scale = -1.0 / sum(c.N)                                   # fused grads are d(-elbo/Nsum)
    ref = c.grads()
    for k in c.keys:
        got = mod.mp[k].vp.grad.cpu().numpy() / scale
        assert rel(got, ref[k]) < GRAD_TOL, (k, rel(got, ref[k]))
    for i in range(c.I):
        assert rel(mod.y_vae[i].mean.grad.cpu().numpy() / scale, ref['vae_mean_%d' % i]) < GRAD_TOL
        assert rel(mod.y_vae[i].log_sd.grad.cpu().numpy() / scale, ref['vae_log_sd_%d' % i]) < GRAD_TOL


def test_fused_trajectory_matches_reference_adam():
    c = Case('t_traj3')
    dev = torch.device('cuda', 0)
    mod = build_model(c, dev)
    fs = FusedStep(mod, lr=float(c.z['meta_lr']))
    elbos = []
    for t in range(c.traj):
        g, cell = _noise(c, dev, t)
        elbos.append(fs.step(c.idx(t), global_noise=g, cell_noise=cell)[0].item())
    assert rel(elbos, c.z['out_elbos']) < ELBO_TOL
    for k in c.keys:     # model.mp[k].vp are views of theta
        assert rel(mod.mp[k].vp.detach().cpu().numpy(), c.z['final_vp_' + k]) < GRAD_TOL, k
    for i in range(c.I):
        assert rel(mod.y_vae[i].mean.detach().cpu().numpy(), c.z['final_vae_mean_%d' % i]) < GRAD_TOL
        assert rel(mod.y_vae[i].log_sd.detach().cpu().numpy(), c.z['final_vae_log_sd_%d' % i]) < GRAD_TOL


def test_fused_philox_global_noise_and_nan_scrub():
    c = Case('a_reftest_fresh')
    dev = torch.device('cuda', 0)
    mod = build_model(c, dev, noise='philox')
    fs = FusedStep(mod, lr=0.05, seed=3)
    draws = []
    for t in range(40):
        fs.step(c.idx(), update=False)
        draws.append(fs.stash[fs.R:2 * fs.R].clone())
    e = torch.stack(draws).cpu().numpy()
    assert abs(e.mean()) < 0.02 and abs(e.std() - 1) < 0.02
    assert not np.allclose(e[0], e[1])
    # a NaN in the likelihood gradient of a global block is zeroed and flagged; parameters stay finite
    before = fs.theta.clone()
    sc = fs.step(c.idx())
    assert sc[4].item() == 0.0 and torch.isfinite(fs.theta).all() and not torch.equal(fs.theta, before)
    mod.mp['alpha'].vp.data[0] = float('nan')
    sc = fs.step(c.idx())
    assert sc[4].item() == 1.0
    assert torch.isfinite(mod.mp['H'].vp).all()


def test_recognise_priors_falls_back_on_unknown_family():
    from torch.distributions import Normal
    c = Case('a_reftest_fresh')
    pri = cy.model.default_priors(c.y, K=c.K, L=[c.L0, c.L1], y_bounds=[-5., -3.5, -2.])
    assert recognise_priors(pri, True) is not None
    pri['sig2'] = torch.distributions.Gamma(.1, 1.)
    assert recognise_priors(pri, True)['sig2_family'] == 1
    pri['alpha'] = Normal(0., 1.)
    assert recognise_priors(pri, True) is None


def test_step_from_host_matches_resident_step():
    """e2e path (sample-by-sample upload from pinned memory overlapped with compute) gives the
    same ELBO / gradients as the resident step for the same step counter and seed."""
    c = Case('i_cfg4like_trained')
    dev = torch.device('cuda', 0)
    outs = []
    for mode in ('resident', 'host'):
        mod = build_model(c, dev, noise='philox')
        fs = FusedStep(mod, lr=0.05, seed=11)
        y_pin = torch.cat([t.float() for t in c.y], 0).pin_memory()
        idx = [range(n) for n in c.N]
        if mode == 'resident':
            sc = fs.step(idx, update=False)
        else:
            mod.engine.y_dev.zero_()                       # the host path must re-upload everything
            sc = fs.step_from_host(y_pin, update=False)
        outs.append((sc.clone().cpu().numpy(), fs.grad.clone().cpu().numpy()))
    assert rel(outs[1][0][:4], outs[0][0][:4]) < 1e-9
    assert rel(outs[1][1], outs[0][1]) < 1e-6


def test_step_from_host_alternates_device_buffers():
    """consecutive host-fed steps upload into alternating device copies of the matrix (so the next upload
    overlaps the tail of the previous step): every call must see exactly the matrix it was given, and the
    resident matrix afterwards is the last one uploaded (a captured graph of the old address is dropped)."""
    c = Case('i_cfg4like_trained')
    dev = torch.device('cuda', 0)
    y0 = torch.cat([t.float() for t in c.y], 0)
    mats = [y0, y0 * 0.5, y0 + 0.25, y0 * 0.5]
    idx = [range(n) for n in c.N]
    ref = []
    for ymat in mats:                      # resident reference: a fresh model per matrix, same step counter
        mod = build_model(c, dev, noise='philox')
        fs = FusedStep(mod, lr=0.05, seed=11)
        mod.engine.y_dev.copy_(ymat.to(dev))
        for _ in range(len(ref) + 1):      # advance the step counter without touching the parameters
            fs.step(idx, update=False)
        ref.append(fs.step(idx, update=False).clone().cpu().numpy())
    mod = build_model(c, dev, noise='philox')
    fs = FusedStep(mod, lr=0.05, seed=11)
    fs.step(idx, update=False)             # step 1 (capture() would otherwise take a training step first)
    fs.capture(None)
    assert fs.graph is not None
    got = []
    for ymat in mats:
        got.append(fs.step_from_host(ymat.pin_memory(), update=False).clone().cpu().numpy())
    assert fs.graph is None
    for a, b in zip(got, ref):
        assert rel(a[:4], b[:4]) < 1e-9
    assert torch.equal(mod.engine.y_dev.cpu().nan_to_num(), mats[-1].nan_to_num())
