"""CPU: host-side logic of the product (variational family, transforms, priors,
autograd plumbing of the C-ABI call, update loop) with the CUDA engine replaced
by an oracle-backed stub, against the golden fixtures of the live reference."""
import numpy as np
import pytest
import torch
from torch.distributions.transforms import StickBreakingTransform

import cytofpy_b200 as cy
from cytofpy_b200.model import variational as V
from tests._golden import Case, case_names, rel
from tests._model_util import build_model, model_grads, noise_queue, replay_noise
from tests._stub_engine import OracleEngine

F64 = torch.float64


def test_public_names_match_reference_surface():
    for name in ('fit', 'update', 'Model', 'default_priors', 'prob_miss', 'compute_Z', 'solve_beta',
                 'gen_beta_est', 'ModelParam', 'VAE', 'check_nan_in_grad'):
        assert hasattr(cy.model, name), name
    assert callable(cy.model.fit)
    for name in ('simdata', 'readCB', 'preprocess'):
        assert hasattr(cy.util, name)


def test_solve_beta_known_answer():
    b = cy.model.solve_beta(np.array([-5.0, -3.0, -1.0]), np.array([.05, .8, .05]))
    assert np.allclose(b.numpy(), [-8.35785565, -6.49610001, -1.08268334], atol=5e-9)
    p = cy.model.prob_miss(torch.tensor([-5.0, -3.0, -1.0]), b[0], b[1], b[2])
    assert np.allclose(p.numpy(), [.05, .8, .05], atol=1e-12)


def test_stick_breaking_matches_torch_transform():
    g = torch.Generator().manual_seed(0)
    x = torch.randn((3, 5, 4), dtype=F64, generator=g) * 3
    sb = StickBreakingTransform(0)
    y = V.stick_breaking(x)
    assert torch.allclose(y, sb(x), rtol=0, atol=1e-15)
    assert torch.allclose(V.stick_breaking_logdet(x, y), sb.log_abs_det_jacobian(x, sb(x)), atol=1e-12)
    assert torch.allclose(y.sum(-1), torch.ones(3, 5, dtype=F64))


def test_compute_Z_is_binary_with_soft_gradient():
    v = torch.tensor([.3, .6, .9], dtype=F64, requires_grad=True)
    H = torch.tensor([[.5, .5, .5], [.2, .7, .95]], dtype=F64, requires_grad=True)
    Z = cy.model.compute_Z(v, H, 0.1)
    assert torch.equal(Z.detach().round(), torch.tensor([[0., 1., 1.], [1., 0., 0.]], dtype=F64))
    Z.sum().backward()
    assert v.grad.abs().sum() > 0 and H.grad.abs().sum() > 0


@pytest.mark.parametrize('name', case_names())
def test_model_forward_backward_on_stub_engine(name):
    """Model.forward + backward (global part in torch, per-cell part via the oracle stub)
    reproduces the live reference: tolerance 1e-5 (elbo) / 1e-4 (grads) is the north-star
    bar; with the fp32 hand-off of the globals it lands around 1e-7."""
    c = Case(name)
    mod = build_model(c, 'cpu', engine_factory=OracleEngine)
    with replay_noise(noise_queue(c), 'cpu') as q:
        elbo, ll, lp, lq = mod(c.idx())
    assert not q
    assert abs(elbo.item() - c.out('elbo')) <= 1e-5 * abs(c.out('elbo'))
    assert abs(ll - c.out('ll')) <= 1e-5 * abs(c.out('ll'))
    assert abs(lp - c.out('lp')) <= 1e-9 * max(1, abs(c.out('lp')))
    assert abs(lq - c.out('lq')) <= 1e-5 * max(1, abs(c.out('lq')))
    elbo.backward()
    ref, got = c.grads(), model_grads(mod, c)
    for k in ref:
        assert rel(got[k], ref[k]) < 1e-4, (k, rel(got[k], ref[k]))


def test_update_trajectory_matches_reference():
    c = Case('t_traj3')
    mod = build_model(c, 'cpu', engine_factory=OracleEngine)
    opt = torch.optim.Adam(mod.parameters(), lr=float(c.z['meta_lr']))
    elbos = []
    for t in range(c.traj):
        with replay_noise(noise_queue(c, t), 'cpu'):
            elbo, fixed, ll, lp, lq = cy.model.update(opt, mod, c.idx(t))
        elbos.append(elbo.item())
        assert fixed is False
    assert rel(elbos, c.z['out_elbos']) < 1e-5
    for k in c.keys:
        assert rel(mod.mp[k].vp.detach().numpy(), c.z['final_vp_' + k]) < 1e-4, k


def test_nan_scrub_in_update():
    c = Case('a_reftest_fresh')
    mod = build_model(c, 'cpu', engine_factory=OracleEngine)

    class Poison(torch.optim.SGD):
        def step(self):
            pass
    opt = Poison(mod.parameters(), lr=0.0)
    orig = mod.log_prior

    def bad_prior(reals, params, idx=None):
        return orig(reals, params, idx) + (params['alpha'] * float('nan')).sum()
    mod.log_prior = bad_prior
    with replay_noise(noise_queue(c), 'cpu'):
        elbo, fixed, *_ = cy.model.update(opt, mod, c.idx())
    assert fixed is True
    assert not torch.isnan(mod.mp['alpha'].vp.grad).any()


def test_simdata_and_missing_injection_shapes():
    d = cy.util.simdata(N=[30, 10, 20], J=8, L0=3, L1=3, seed=0)
    assert [tuple(t.shape) for t in d['data']['y']] == [(30, 8), (10, 8), (20, 8)]
    s = cy.util.synth_cytof([2000, 1000], 32, [100.] * 5, 5, 3, seed=1)
    beta = cy.model.solve_beta([-5., -3.5, -2.], [.05, .8, .05]).numpy()
    cy.util.inject_missing(s['y'], beta, rate=0.1, seed=2)
    frac = np.isnan(np.concatenate(s['y'])).mean()
    assert 0.02 < frac < 0.12


def test_readcb_roundtrip(tmp_path):
    p = tmp_path / 'cb.txt'
    p.write_text('N 2 1\nCD3 CD4\n1.5 NA\n-7.0 2.0\n0.25 -1.0\n')
    d = cy.util.readCB(str(p))
    assert d['N'] == [2, 1] and d['markers'] == ['CD3', 'CD4']
    assert torch.isnan(d['y'][0][0, 1]) and d['m'][0][0, 1]
    cy.util.preprocess(d, rm_cells_below=-6.0)
    assert d['N'] == [1, 1] and d['y'][0].shape == (1, 2)


@pytest.mark.parametrize('max_iter,backup_every,trace_every', [
    (1000, 10, 20), (47, 7, 5), (30, 10, 0), (25, 0, 0), (12, 1, 3), (100, 64, 50), (5, 10, 10), (0, 10, 10)])
def test_pipelined_fit_schedule_never_runs_past_a_snapshot(max_iter, backup_every, trace_every):
    """fit(pipeline=True) replays multi-step graphs between parameter snapshots (train.py): the schedule must
    cover steps 0..max_iter exactly once, start host-driven, and a snapshot step may only be the LAST step of
    a chunk -- otherwise the backup / trace copies would hold later parameters than the reference's."""
    from cytofpy_b200.model import train
    chunk = train.chunk_length(backup_every, trace_every)
    assert 1 <= chunk <= 32
    sched = train.step_schedule(max_iter, backup_every, trace_every, chunk)
    assert sched[0] == (0, 1)
    covered = [t + j for t, n in sched for j in range(n)]
    assert covered == list(range(max_iter + 1))
    for t, n in sched:
        assert n in (1, chunk)
        for j in range(n - 1):
            assert not train.is_snapshot_step(t + j, backup_every, trace_every), (t, n, j)
    periods = [v for v in (backup_every, trace_every) if v > 0]
    if chunk > 1 and max_iter >= 4 * chunk and all(v % chunk == 0 for v in periods):
        # aligned snapshot periods (the defaults: 10 and max_iter / 50): all but the edges run inside chunks
        assert sum(n for _, n in sched if n > 1) >= max_iter + 1 - 2 * chunk


def _write_cb_file(path, N, J, seed=0, na_rate=0.2, low_rate=0.01):
    """A file in the cord-blood text format (reference cytofpy/util/readCB.py:3-28): header
    `N n_1 ... n_I`, marker names, one row per cell, `NA` for missing."""
    rng = np.random.default_rng(seed)
    with open(path, 'w') as f:
        f.write('N ' + ' '.join(str(n) for n in N) + '\n')
        f.write(' '.join('M%d' % j for j in range(J)) + '\n')
        for n in N:
            y = rng.normal(0.0, 2.0, size=(n, J))
            y[rng.random((n, J)) < low_rate] = -7.5           # cells preprocess() must drop
            miss = (y < 0) & (rng.random((n, J)) < na_rate)
            for row, mrow in zip(y, miss):
                f.write(' '.join('NA' if mm else repr(float(v)) for v, mm in zip(row, mrow)) + '\n')


def test_readcb_and_preprocess_match_the_reference_reader(tmp_path):
    """f4: the same file through cytofpy_b200.util.readCB / preprocess and through the reference's reader
    (imported from /root/reference when it is there: build container only) gives identical tensors."""
    import os
    import sys
    N, J = [700, 450, 520], 32
    p = tmp_path / 'cb.txt'
    _write_cb_file(str(p), N, J)
    d = cy.util.readCB(str(p))
    assert d['N'] == N and len(d['markers']) == J and [tuple(t.shape) for t in d['y']] == [(n, J) for n in N]
    assert all(t.dtype == torch.float64 for t in d['y'])
    raw_nan = [int(torch.isnan(t).sum()) for t in d['y']]
    assert all(c > 0 for c in raw_nan)
    cy.util.preprocess(d, rm_cells_below=-6.0)
    assert all(a < b for a, b in zip(d['N'], N))              # some cells went away
    for yi, mi in zip(d['y'], d['m']):
        assert bool(((yi > -6.0) | torch.isnan(yi)).all()) and torch.equal(mi, torch.isnan(yi))
    if os.path.isdir('/root/reference/cytofpy'):
        import subprocess
        code = r"""
import sys, torch, pickle
sys.path.insert(0, '/root/reference')
from cytofpy.util.readCB import readCB, preprocess
d = readCB(%r); preprocess(d, rm_cells_below=-6.0)
pickle.dump({'N': d['N'], 'y': [t.double() for t in d['y']], 'm': d['m'], 'markers': d['markers']}, open(%r, 'wb'))
""" % (str(p), str(tmp_path / 'ref.p'))
        out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        import pickle
        ref = pickle.load(open(str(tmp_path / 'ref.p'), 'rb'))
        assert ref['N'] == d['N'] and ref['markers'] == d['markers']
        for a, b, ma, mb in zip(d['y'], ref['y'], d['m'], ref['m']):
            # the reference parses into torch's default dtype at import order; values agree to fp32 at least
            assert torch.allclose(torch.nan_to_num(a, nan=99.0), torch.nan_to_num(b, nan=99.0), rtol=1e-6, atol=0)
            assert torch.equal(ma, mb)


def test_model_param_and_vae_pickle_as_cpu_copies():
    """f3: the blocks fit() returns pickle by value, on the CPU, whatever they were views of."""
    import pickle
    from cytofpy_b200.model.variational import ModelParam, VAE
    flat = torch.arange(40, dtype=torch.float64)
    blk = ModelParam((2, 5), 'simplex')
    blk.vp = torch.nn.Parameter(flat[10:30].view(2, 2, 5))          # a view into a larger buffer
    data = pickle.dumps({'W': blk, 'vae': [VAE(4, mean_init=-2.5, sd_init=.3)]})
    assert len(data) < 4000                                          # not the whole 40-element storage + module state
    back = pickle.loads(data)
    assert torch.equal(back['W'].vp.detach(), blk.vp.detach()) and back['W'].support == 'simplex'
    assert back['W'].dist().mean.shape == (2, 5)
    v = back['vae'][0]
    assert torch.allclose(v.mean, torch.full((1, 4), -2.5, dtype=torch.float64))
    y = torch.tensor([[0.5, float('nan'), 1.0, -1.0]], dtype=torch.float64)
    yi, lq = v(y, torch.isnan(y), 10)
    assert yi.shape == (1, 4) and not torch.isnan(yi).any() and yi[0, 0] == 0.5


def test_on_gpu_node_restores_the_affinity_and_never_raises():
    """util.on_gpu_node: without NVML / a GPU it changes nothing and yields 0; with one it must hand the calling
    thread its previous CPU mask back (the bench's CPU-baseline leg runs later in the same process)."""
    import os
    import cytofpy_b200 as cy
    before = os.sched_getaffinity(0)
    with cy.util.on_gpu_node(0) as n:
        inside = os.sched_getaffinity(0)
        assert n == 0 or (n == len(inside) and inside <= before)
    assert os.sched_getaffinity(0) == before
    assert cy.util.gpu_local_cpus(99) == set()      # no such device: empty, no exception
