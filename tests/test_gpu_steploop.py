"""GPU: the small step-loop kernels behind the C ABI — fused Adam with NaN scrub
(reference fit.py:24-30,83) and the without-replacement minibatch sampler
(reference fit.py:96-102) — plus an end-to-end `fit` smoke run."""
import ctypes as C

import numpy as np
import pytest
import torch

import cytofpy_b200 as cy
from cytofpy_b200 import _lib

pytestmark = pytest.mark.gpu
F64 = torch.float64


def _p(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


def test_adam_kernel_matches_torch_adam_and_scrubs_nan():
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    n, n_scrub = 3528, 3336
    p0 = torch.randn(n, dtype=F64, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=0.1)
    p = p0.clone().to(dev)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    for t in range(1, 6):
        grad = torch.randn(n, dtype=F64, generator=g) * 10
        ref.grad = (grad * -0.001).clone()
        if t == 3:
            grad[5] = float('nan')          # inside the scrubbed range
            ref.grad[5] = 0.0
        opt.step()
        use_dev_counter = t >= 4
        if t == 4:
            step_dev.fill_(3)
        grad_dev = grad.to(dev)            # held until the synchronize below (the call is asynchronous)
        _lib.check(lib.cytof_adam_step_f64(_p(p), _p(grad_dev), _p(m), _p(v), n, n_scrub, -0.001, 0.1,
                                           0.9, 0.999, 1e-8, t, _p(step_dev) if use_dev_counter else _p(None),
                                           _p(flag), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        torch.cuda.synchronize()
        assert flag.item() == (1 if t >= 3 else 0)
        assert torch.allclose(p.cpu(), ref.detach(), rtol=1e-12, atol=1e-14), t
    assert step_dev.item() == 5
    # a NaN outside the scrubbed range propagates (reference leaves imputation grads alone)
    grad = torch.zeros(n, dtype=F64)
    grad[n - 1] = float('nan')
    _lib.check(lib.cytof_adam_step_f64(_p(p), _p(grad.to(dev)), _p(m), _p(v), n, n_scrub, 1.0, 0.1, 0.9,
                                       0.999, 1e-8, 6, _p(None), _p(None), C.c_void_p(0)))
    assert torch.isnan(p[n - 1]).item() and not torch.isnan(p[:n - 1]).any().item()


@pytest.mark.parametrize('N,m', [(40000, 2000), (41474, 500), (17, 16), (2, 1), (1000003, 1000)])
def test_sampler_without_replacement(N, m):
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    out = torch.empty(m, dtype=torch.int32, device=dev)

    def draw(seed, step, sample):
        _lib.check(lib.cytof_sample_minibatch(N, m, seed, step, sample, _p(out), C.c_void_p(0)))
        return out.cpu().numpy().copy()
    a = draw(1, 0, 0)
    assert a.min() >= 0 and a.max() < N and len(np.unique(a)) == m        # distinct, in range
    assert np.array_equal(a, draw(1, 0, 0))                               # deterministic per key
    if N > 100:
        assert not np.array_equal(a, draw(1, 1, 0)) and not np.array_equal(a, draw(1, 0, 1))
        assert not np.array_equal(a, draw(2, 0, 0))


def test_sampler_is_roughly_uniform_and_full_batch_is_identity():
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    N, m = 1000, 100
    out = torch.empty(m, dtype=torch.int32, device=dev)
    hits = np.zeros(N)
    for step in range(2000):
        _lib.check(lib.cytof_sample_minibatch(N, m, 7, step, 0, _p(out), C.c_void_p(0)))
        hits[out.cpu().numpy()] += 1
    # each index expected 200 times, sd ~ 13.4; allow 6 sigma
    assert abs(hits.mean() - 200) < 1e-9 and hits.min() > 120 and hits.max() < 280
    full = torch.empty(N, dtype=torch.int32, device=dev)
    _lib.check(lib.cytof_sample_minibatch(N, N + 5, 7, 0, 0, _p(full), C.c_void_p(0)))
    assert np.array_equal(full.cpu().numpy(), np.arange(N))


def test_sampler_pair_cooccurrence_and_positions():
    """Second-order checks of the keyed Feistel sampler: over many steps every PAIR of indices must land in the same
    minibatch about as often as under uniform sampling without replacement (a weak mixing function shows up here
    long before it shows in the marginals), and the position of an index inside the minibatch must be uniform too
    (the kernel walks consecutive minibatch rows, so position bias would correlate neighbours)."""
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    N, m, steps = 60, 12, 6000
    out = torch.empty(m, dtype=torch.int32, device=dev)
    co = np.zeros((N, N))
    pos = np.zeros((N, m))
    for step in range(steps):
        _lib.check(lib.cytof_sample_minibatch(N, m, 11, step, 2, _p(out), C.c_void_p(0)))
        a = out.cpu().numpy()
        ind = np.zeros(N)
        ind[a] = 1
        co += np.outer(ind, ind)
        pos[a, np.arange(m)] += 1
    # P(i and j together) = m (m-1) / (N (N-1)) = 0.0373 -> 223.7 expected, binomial sd 14.7
    off = co[~np.eye(N, dtype=bool)]
    expect = steps * m * (m - 1) / (N * (N - 1))
    sd = np.sqrt(expect * (1 - expect / steps))
    assert abs(off.mean() - expect) < 1e-9                       # exact: every draw has m (m - 1) ordered pairs
    assert off.min() > expect - 6 * sd and off.max() < expect + 6 * sd, (off.min(), off.max(), expect, sd)
    z = (off - expect) / sd
    assert 0.8 < z.std() < 1.2, z.std()                          # neither clumped nor suspiciously regular
    # position uniformity: each (index, position) cell expects steps / N = 100 hits, sd ~ 9.9
    e2 = steps / N
    assert pos.min() > e2 - 6 * np.sqrt(e2) and pos.max() < e2 + 6 * np.sqrt(e2), (pos.min(), pos.max())


def test_fit_with_the_reference_numpy_sampler_stream():
    """sampler='numpy' keeps the reference's np.random.choice minibatch stream on the host (fit.py:96-102): the
    minibatches fit draws are exactly the reference's for the same seed, and the fit trains."""
    torch.manual_seed(0)
    np.random.seed(0)
    N = [300, 100, 200]
    y = cy.util.simdata(N=N, J=8, a_W=[100.] * 5, L0=3, L1=3)['data']['y']
    pri = cy.model.default_priors(y, K=10, L=[3, 3], y_bounds=[-5., -3.5, -2.])
    drawn = []
    orig = np.random.choice

    def spy(n, size, replace):
        r = orig(n, size, replace=replace)
        drawn.append(np.array(r))
        return r
    np.random.choice = spy
    try:
        out = cy.model.fit(y, max_iter=6, lr=1e-1, print_freq=100, eps_conv=0, priors=pri, minibatch_size=50, tau=0.1,
                           trace_every=50, backup_every=50, verbose=0, seed=7, sampler='numpy')
    finally:
        np.random.choice = orig
    assert len(out['elbo']) == 7 and np.isfinite(out['elbo']).all()
    # what the reference's loop draws for seed 7: np.random.seed(seed) at the top of fit, then per step and sample
    np.random.seed(7)
    expect = [orig(n, 50, replace=False) for _ in range(7) for n in N]
    assert len(drawn) == len(expect) and all(np.array_equal(a, b) for a, b in zip(drawn, expect))


def test_fit_runs_and_improves_elbo(capsys):
    """the reference's own smoke test (tests/test_model.py:12-43) against the drop-in, with an
    assertion the reference lacks: the ELBO must improve."""
    torch.manual_seed(0)
    np.random.seed(0)
    data = cy.util.simdata(N=[300, 100, 200], L0=3, L1=3, J=8, a_W=[100.] * 5, seed=0)
    y = [t.clone() for t in data['data']['y']]
    pri = cy.model.default_priors(y, K=10, L=[3, 3], y_bounds=[-5., -3.5, -2.])
    out = cy.model.fit(y, max_iter=120, lr=1e-1, print_freq=40, eps_conv=1e-6, priors=pri,
                       minibatch_size=100, tau=0.1, verbose=0, seed=1)
    assert set(out) == {'elbo', 'trace', 'mp', 'tau', 'vae', 'use_stick_break', 'y', 'priors', 'model_noisy'}
    e = np.array(out['elbo'])
    assert len(e) == 121 and np.isfinite(e).all()
    # the reference itself moves from about -15,000 to about -14,600 on this problem in the
    # first iterations (its own log, tests/test_model.py settings)
    assert e[-20:].mean() > e[:3].mean() + 150
    assert len(out['trace']) >= 30 and 'W' in out['trace'][0]
    assert out['trace'][0]['W'].dist().mean.shape == (3, 9)


def test_fit_with_run_ahead_equals_step_by_step_fit(capsys):
    """fit(pipeline=True) queues step t+1 before it reads the result of step t (except around snapshots):
    ELBO history, trace and backup must be exactly those of the step-by-step loop."""
    outs = []
    for pipeline in (True, False):
        torch.manual_seed(0)
        np.random.seed(0)
        data = cy.util.simdata_with_missing_values(N=[300, 100, 200], L0=3, L1=3, J=8, a_W=[100.] * 5, seed=0,
                                                   rate=0.1)
        y = [t.clone() for t in data['data']['y']]
        pri = cy.model.default_priors(y, K=10, L=[3, 3], y_bounds=[-5., -3.5, -2.])
        outs.append(cy.model.fit(y, max_iter=47, lr=1e-1, print_freq=40, eps_conv=1e-12, priors=pri,
                                 minibatch_size=100, tau=0.1, verbose=0, seed=1, backup_every=7,
                                 trace_every=5, pipeline=pipeline))
    a, b = outs
    assert a['elbo'] == b['elbo'] and len(a['elbo']) == 48
    assert len(a['trace']) == len(b['trace']) == 10
    for ta, tb in zip(a['trace'], b['trace']):
        for key in ta:
            assert torch.equal(ta[key].vp, tb[key].vp)
    for key in a['mp']:
        assert torch.equal(a['mp'][key].vp, b['mp'][key].vp)
    for va, vb in zip(a['vae'], b['vae']):
        assert torch.equal(va.mean, vb.mean) and torch.equal(va.log_sd, vb.log_sd)


def test_posterior_label_sampler_matches_oracle():
    """post_proc.theta + post_proc.sample (reference post_proc/lam_post.py:6-55): class
    probabilities vs the fp64 oracle, and the categorical draws vs those probabilities."""
    from oracle import lam_post_port
    from tests._golden import Case
    from tests._model_util import build_model
    c = Case('c_nan_nostick_tau001')
    dev = torch.device('cuda', 0)
    mod = build_model(c, dev, noise='philox')
    pri = cy.model.default_priors(c.y, K=c.K, L=[c.L0, c.L1], y_bounds=[-5., -3.5, -2.])
    theta, _ = cy.model.post_proc.theta(mod, pri)
    assert [tuple(t.shape) for t in theta['y']] == [(n, c.J) for n in c.N]
    assert not any(torch.isnan(t).any().item() for t in theta['y'])          # imputed
    for t, y0 in zip(theta['y'], c.y):                                        # observed entries untouched
        obs = ~torch.isnan(y0)
        assert torch.allclose(t.cpu()[obs], y0[obs].float().double(), atol=1e-6)
    lam, probs = cy.model.post_proc.sample(theta, mod, return_probs=True, seed=5)
    th_np = {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in theta.items() if k != 'y'}
    th_np['noisy_sd'] = float(theta['noisy_sd'])
    th_np['y'] = [t.cpu().numpy() for t in theta['y']]
    ref = lam_post_port.label_probs(th_np)
    for i in range(c.I):
        p = probs[i].cpu().numpy()
        assert p.shape == (c.N[i], c.K + 1)
        assert np.max(np.abs(p - ref[i])) < 2e-5
        assert np.allclose(p.sum(1), 1.0, atol=1e-5)
        li = lam[i].cpu().numpy()
        assert li.dtype == np.int64 and li.min() >= 0 and li.max() <= c.K
        sure = ref[i].max(1) > 0.999
        assert (li[sure] == ref[i].argmax(1)[sure]).mean() > 0.99
    # repeated draws follow the probabilities (cell 0 of sample 0)
    draws = np.stack([cy.model.post_proc.sample(theta, mod, seed=100 + s)[0].cpu().numpy() for s in range(300)])
    k_star = int(ref[0][:, 1:].mean(0).argmax()) + 1
    emp = (draws == k_star).mean()
    assert abs(emp - ref[0][:, k_star].mean()) < 0.03


def _small_model(dev, nan_rate):
    torch.manual_seed(0)
    np.random.seed(0)
    N, J, K, L = [600, 500, 700], 16, 8, [3, 3]
    sim = cy.util.simdata_with_missing_values(N=N, J=J, a_W=[100.] * 5, L0=L[0], L1=L[1], seed=0, rate=nan_rate)
    y = sim['data']['y']
    pri = cy.model.default_priors(y, K=K, L=L, y_bounds=[-5., -3.5, -2.])
    return cy.model.Model(y=[t.clone() for t in y], priors=pri, tau=0.1, verbose=-1, device=dev,
                          noise='philox', seed=3)


@pytest.mark.parametrize('mb,nan_rate', [(200, 0.2), (None, 0.0)])
def test_cuda_graph_step_equals_host_driven_steps(mb, nan_rate):
    """One captured CUDA graph (device sampler + global fwd + per-cell kernel + finalise + global bwd /
    Adam, all step-dependent inputs read from a device counter) must reproduce, bit for bit, the
    host-driven steps that pass the same step numbers through the descriptors."""
    from cytofpy_b200.model.fused import FusedStep
    from cytofpy_b200.model.train import device_minibatch
    dev = torch.device('cuda', 0)
    steps = 6
    ref_mod = _small_model(dev, nan_rate)
    ref = FusedStep(ref_mod, lr=0.05, seed=3)
    hist_ref = []
    for _ in range(steps):
        s = ref.t + 1      # the step number the graph reads from its device counter
        idx = [range(n) for n in ref_mod.N_local] if mb is None else device_minibatch(ref_mod, mb, s, ref.g.seed)
        hist_ref.append(ref.step(idx).tolist())
    g_mod = _small_model(dev, nan_rate)
    gs = FusedStep(g_mod, lr=0.05, seed=3)
    s = gs.t + 1
    idx = [range(n) for n in g_mod.N_local] if mb is None else device_minibatch(g_mod, mb, s, gs.g.seed)
    hist = [gs.step(idx).tolist()]             # step 1 host-driven, then capture
    gs.capture(mb)
    for _ in range(steps - 1):
        hist.append(gs.graph_step().tolist())
    assert gs.t == steps
    assert np.array_equal(np.array(hist), np.array(hist_ref))
    assert torch.equal(gs.theta, ref.theta)
    gs.release_graph()
    # and host-driven steps continue seamlessly after the graph is released
    s = gs.t + 1
    idx = [range(n) for n in g_mod.N_local] if mb is None else device_minibatch(g_mod, mb, s, gs.g.seed)
    a = gs.step(idx).tolist()
    s = ref.t + 1
    idx = [range(n) for n in ref_mod.N_local] if mb is None else device_minibatch(ref_mod, mb, s, ref.g.seed)
    assert a == ref.step(idx).tolist()


@pytest.mark.parametrize('nan_rate', [0.2, 0.0])
def test_graph_steps_reading_the_matrix_from_pinned_host_memory(nan_rate):
    """FusedStep.use_host_matrix: the per-cell kernel gathers the minibatch rows straight from a pinned host
    matrix (no staging copy); the captured step must reproduce the device-resident one bit for bit."""
    from cytofpy_b200.model.fused import FusedStep
    dev = torch.device('cuda', 0)
    mb, steps = 200, 5
    ref_mod = _small_model(dev, nan_rate)
    ref = FusedStep(ref_mod, lr=0.05, seed=3)
    ref.capture(mb)
    hist_ref = [ref.graph_step().tolist() for _ in range(steps)]
    mod = _small_model(dev, nan_rate)
    fs = FusedStep(mod, lr=0.05, seed=3)
    y_pin = mod.engine.y_dev.cpu().pin_memory()
    fs.capture(mb)                       # takes its first (host-driven) step from the resident copy, like `ref`
    fs.use_host_matrix(y_pin)
    assert fs.graph is None
    mod.engine_resident = fs._y_resident
    fs._y_resident.zero_()               # the resident copy must not be read from here on
    fs.capture(mb)
    hist = [fs.graph_step().tolist() for _ in range(steps)]
    assert hist == hist_ref
    assert torch.equal(fs.theta, ref.theta)
    fs.use_device_matrix()
    assert mod.engine.y_dev.is_cuda
    with pytest.raises(ValueError):
        fs.use_host_matrix(y_pin[:-1])


def test_multi_sample_sampler_equals_per_sample_calls():
    """cytof_sample_minibatch_multi (one launch for all samples, host step or device step counter)
    draws exactly the indices of the per-sample cytof_sample_minibatch calls."""
    dev = torch.device('cuda', 0)
    lib = _lib.load()
    N, m = [40000, 41474, 17, 500], [2000, 500, 16, 800]      # last sample: m >= N -> 0..N-1 in order
    I = len(N)
    seed, step = 12345, 7
    take = [min(a, b) for a, b in zip(m, N)]
    ref = torch.empty(sum(take), dtype=torch.int32, device=dev)
    off = 0
    for i in range(I):
        _lib.check(lib.cytof_sample_minibatch(N[i], m[i], seed, step, i, C.c_void_p(ref.data_ptr() + 4 * off),
                                              C.c_void_p(0)))
        off += take[i]
    n_arr, m_arr = (C.c_int64 * I)(*N), (C.c_int64 * I)(*m)
    out = torch.full_like(ref, -1)
    _lib.check(lib.cytof_sample_minibatch_multi(I, n_arr, m_arr, seed, step, C.c_void_p(0), _p(out), C.c_void_p(0)))
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    step_dev = torch.tensor([step - 1], dtype=torch.int64, device=dev)      # device counter: step = *step_dev + 1
    out2 = torch.full_like(ref, -1)
    _lib.check(lib.cytof_sample_minibatch_multi(I, n_arr, m_arr, seed, 0, _p(step_dev), _p(out2), C.c_void_p(0)))
    torch.cuda.synchronize()
    assert torch.equal(out2, ref)
    assert torch.equal(out[-500:].cpu(), torch.arange(500, dtype=torch.int32))


def test_graph_step_with_in_kernel_sampler_equals_the_sampler_kernel(monkeypatch):
    """capture() for minibatches lets the per-cell kernel draw the rows (no sampler node, no index buffer);
    CYTOF_SAMPLER_KERNEL=1 keeps the separate sampler launch.  Same rows, so the same parameters bit for bit."""
    from cytofpy_b200.model.fused import FusedStep
    thetas, launches = [], []
    for env in ('0', '1'):
        monkeypatch.setenv('CYTOF_SAMPLER_KERNEL', env)
        torch.manual_seed(3)
        s = cy.util.synth_cytof([3000, 2000, 2500], 32, [100.] * 5, 5, 3, seed=6)
        y = [torch.from_numpy(a) for a in s['y']]
        priors = cy.model.default_priors([t.double() for t in y], K=30, L=[5, 3], y_bounds=[-5., -3.5, -2.])
        model = cy.model.Model(y=y, priors=priors, verbose=-1, device=torch.device('cuda', 0), noise='philox', seed=1)
        fs = FusedStep(model, lr=1e-2, seed=1)
        fs.capture(400)
        for _ in range(6):
            sc = fs.graph_step()
        torch.cuda.synchronize()
        assert torch.isfinite(sc).all()
        thetas.append(fs.theta.clone())
        launches.append(fs._graph_launches)
        fs.release_graph()
    assert torch.equal(thetas[0], thetas[1])
    assert launches[0] == launches[1] - 1
