"""CPU: the oracle (oracle/elbo_port.py, oracle/cells_closed_form.py) against the
golden fixtures recorded from the live reference, and against the one
known-answer value the reference tree holds (Model.py:47-49)."""
import numpy as np
import pytest
import torch

from oracle import cells_closed_form as ccf
from oracle import elbo_port as port
from tests._golden import Case, case_names, rel

F64 = torch.float64
ROOT_DIR = __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))


def test_solve_beta_known_answer():
    # reference cytofpy/model/Model.py:47-49 ("exact: -8.35785565, -6.49610001, -1.08268334")
    b = port.solve_beta([-5.0, -3.0, -1.0], [.05, .8, .05])
    assert np.allclose(b, [-8.35785565, -6.49610001, -1.08268334], atol=5e-9)
    y = np.array([-5.0, -3.0, -1.0])
    p = 1 / (1 + np.exp(-(b[0] + b[1] * y + b[2] * y * y)))
    assert np.allclose(p, [.05, .8, .05], atol=1e-12)


@pytest.mark.parametrize('name', case_names())
def test_port_matches_reference_fixture(name):
    c = Case(name)
    st = c.state()
    out = port.elbo_port(st, c.y, c.idx(), c.noise(), c.cfg())
    for k in ('elbo', 'll', 'lp', 'lq', 'log_qy'):
        ref = c.out(k)
        assert abs(float(out[k]) - ref) <= 1e-11 * max(1.0, abs(ref)), (k, float(out[k]), ref)
    out['elbo'].backward()
    g = c.grads()
    for k in c.keys:
        assert rel(st[k].grad.numpy(), g[k]) < 1e-10, k
    for i in range(c.I):
        assert rel(st['vae_mean'][i].grad.numpy(), g['vae_mean_%d' % i]) < 1e-10
        assert rel(st['vae_log_sd'][i].grad.numpy(), g['vae_log_sd_%d' % i]) < 1e-10


def kernel_inputs_from_port(c, st, nz, cfg, idx):
    """Derive the kernel-level inputs (what the CUDA entry point consumes) from
    the port's transformed parameters."""
    out = port.elbo_port(st, c.y, idx, nz, cfg)
    P = out['params']
    v = torch.cumprod(P['v'], 0) if cfg['use_stick_break'] else P['v']
    g = {
        'mu0': (-torch.cumsum(P['delta0'], 0)).detach().numpy(),
        'mu1': torch.cumsum(P['delta1'], 0).detach().numpy(),
        'sig': torch.sqrt(P['sig2']).detach().numpy(),
        'logeta0': torch.log(P['eta0']).detach().numpy(),
        'logeta1': torch.log(P['eta1']).detach().numpy(),
        'logW': torch.log(P['W']).detach().numpy(),
        'Z': (v[None, :] > P['H']).detach().numpy().astype(np.float64),
        'b': np.stack([cfg['b0'].numpy(), cfg['b1'].numpy(), cfg['b2'].numpy()], 1),
        'vm': np.stack([m.detach().numpy()[0] for m in st['vae_mean']]),
        'vls': np.stack([s.detach().numpy()[0] for s in st['vae_log_sd']]),
    }
    if cfg['model_noisy']:
        g['eps'] = P['eps'].detach().numpy()
    rows = [c.y[i].numpy()[np.asarray(idx[i])] for i in range(c.I)]
    eps = [nz['y'][i].numpy() for i in range(c.I)]
    fac = np.array([cfg['N'][i] / max(1, len(idx[i])) for i in range(c.I)])
    return out, g, rows, eps, fac


@pytest.mark.parametrize('name', case_names())
def test_closed_form_matches_fixture_and_autograd(name):
    c = Case(name)
    st, cfg, idx, nz = c.state(), c.cfg(), c.idx(), c.noise()
    out, g, rows, eps, fac = kernel_inputs_from_port(c, st, nz, cfg, idx)
    ll, lqy, block, y_imp = ccf.cells_fwdbwd(rows, eps, g, fac, cfg['scale'],
                                             float(cfg['noisy_sd']), cfg['model_noisy'])
    assert abs(ll - c.out('ll')) <= 1e-10 * abs(c.out('ll'))
    assert abs(lqy - c.out('log_qy')) <= 1e-10 * max(1.0, abs(c.out('log_qy')))
    # gradient block vs torch autograd of the port's likelihood at the kernel inputs
    I, J, K, L0, L1 = c.I, c.J, c.K, c.L0, c.L1
    lay = ccf.block_layout(I, J, K, L0, L1)
    t = {k: torch.tensor(v, dtype=F64, requires_grad=True) for k, v in g.items() if k != 'b'}
    ll_t, lqy_t = _torch_cells(t, rows, eps, fac, cfg, g['b'])
    ll_t.backward(retain_graph=True)
    sec = lambda n: block[lay[n][0]:lay[n][0] + lay[n][1]]
    pairs = [('dmu0', 'mu0'), ('dmu1', 'mu1'), ('dsig', 'sig'), ('dlogeta0', 'logeta0'),
             ('dlogeta1', 'logeta1'), ('dlogW', 'logW'), ('dZ', 'Z'), ('dvm', 'vm'), ('dvls', 'vls')]
    if cfg['model_noisy']:
        pairs.append(('deps', 'eps'))
    for s, k in pairs:
        ref = t[k].grad.numpy().ravel()
        assert rel(sec(s), ref) < 1e-9 or np.max(np.abs(sec(s) - ref)) < 1e-9, (s, rel(sec(s), ref))
    for k in t:
        t[k].grad = None
    lqy_t.backward()
    assert rel(sec('dlqy_dvls'), t['vls'].grad.numpy().ravel()) < 1e-12 or \
        np.max(np.abs(sec('dlqy_dvls'))) == 0


def _torch_cells(t, rows, eps, fac, cfg, b):
    """Independent torch fp64 statement of the per-cell part at kernel inputs
    (autograd supplies the gradient the closed form must match)."""
    LOG2PI = float(np.log(2 * np.pi))
    ll, lqy = 0.0, 0.0
    for i in range(len(rows)):
        yd = torch.tensor(rows[i], dtype=F64)
        miss = torch.isnan(yd)
        e = torch.tensor(eps[i], dtype=F64)
        y = torch.where(miss, t['vm'][i][None] + torch.exp(t['vls'][i])[None] * e, torch.nan_to_num(yd))
        s = t['sig'][i]
        A = []
        for z, (mu, le) in enumerate(((t['mu0'], t['logeta0']), (t['mu1'], t['logeta1']))):
            d = y[:, :, None] - mu[None, None, :]
            a = le[i][None] - torch.log(s) - .5 * LOG2PI - d * d / (2 * s * s)
            A.append(torch.logsumexp(a, 2))
        f = A[0].sum(1, keepdim=True) + (A[1] - A[0]) @ t['Z'] + t['logW'][i][None]
        L = torch.logsumexp(f, 1)
        if cfg['model_noisy']:
            sd = float(cfg['noisy_sd'])
            zz = (-.5 * LOG2PI - np.log(sd) - y * y / (2 * sd * sd)).sum(1) + torch.log(t['eps'][i])
            L = torch.logaddexp(L + torch.log1p(-t['eps'][i]), zz)
        u = b[i, 0] + b[i, 1] * y + b[i, 2] * y * y
        ll = ll + fac[i] * (L.sum() + torch.nn.functional.logsigmoid(u)[miss].sum())
        lqy = lqy + fac[i] * cfg['scale'] * (-.5 * e * e - t['vls'][i][None] - .5 * LOG2PI)[miss].sum()
    return ll, lqy


@pytest.mark.skipif(not __import__('os').path.isdir('/root/reference/cytofpy'),
                    reason='needs the live reference (build container only)')
def test_lam_post_port_matches_reference():
    """oracle/lam_post_port.py vs the probabilities the live reference hands to its own sampler:
    `Categorical` inside cytofpy.model.post_proc.lam_post is replaced by a recorder, so the
    captured tensors ARE the `lam_probs` of reference lam_post.py:51-53 (no restated formula)."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, '/root/reference'); sys.path.insert(0, %r)
torch.distributions.Distribution.set_default_validate_args(False)
import cytofpy
from cytofpy.model.post_proc import lam_post as ref_lam_post
from oracle import lam_post_port
torch.manual_seed(0); np.random.seed(0)
y = cytofpy.util.simdata(N=[60, 40], J=6, a_W=[100.]*3, L0=3, L1=2)['data']['y']
y[0][0, 1] = float('nan')
pri = cytofpy.model.default_priors(y, K=4, L=[3, 2], y_bounds=[-5., -3.5, -2.])
mod = cytofpy.model.Model(y=y, priors=pri, verbose=-1)
th, _ = cytofpy.model.post_proc.sample_params.theta(mod, pri)
captured = []
real_categorical = ref_lam_post.Categorical
class Recorder(real_categorical):
    def __init__(self, probs=None, **kw):
        captured.append(probs.detach().clone().numpy())
        super().__init__(probs, **kw)
ref_lam_post.Categorical = Recorder
lam = ref_lam_post.sample(th, mod)
ref_lam_post.Categorical = real_categorical
assert len(captured) == mod.I and all(c.shape == (n, 5) for c, n in zip(captured, [60, 40]))
assert all(int(l.max()) <= 4 for l in lam)
tn = {k: (v.numpy() if torch.is_tensor(v) else v) for k, v in th.items() if k != 'y'}
tn['noisy_sd'] = float(th['noisy_sd']); tn['y'] = [t.numpy() for t in th['y']]
got = lam_post_port.label_probs(tn)
print(max(float(np.abs(a - b).max()) for a, b in zip(got, captured)))
''' % ROOT_DIR
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert float(out.stdout.strip().splitlines()[-1]) < 1e-12
