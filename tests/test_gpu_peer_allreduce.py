"""GPU, world_size = 2: the one-shot all-reduce over peer-memory mailboxes that is fused into the
finalise kernel (`cytof_elbo_fwdbwd_allreduce`, cytofpy_b200/model/peer.py).

Both ranks run as separate processes and share cuda:0 (CUDA IPC works between processes on one
device, so the test also runs on a one-GPU box); gloo only carries the 64-byte IPC handles.  Each
rank holds half of every sample's rows of a golden fixture and runs the fused step with the
fixture's noise: every rank must end with the reference's ELBO / gradients, the two ranks must
agree bit for bit, and the parameters after an Adam step must stay identical (replicas in sync)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests._golden import Case, rel

pytestmark = pytest.mark.gpu
F64 = torch.float64
CASE = 'f_fullbatch_mixed'


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir, mode):
    import cytofpy_b200 as cy
    from cytofpy_b200.model.fused import FusedStep
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    c = Case(CASE)
    lo = [n * rank // world for n in c.N]
    hi = [n * (rank + 1) // world for n in c.N]
    y = [t[a:b].clone() for t, a, b in zip(c.y, lo, hi)]
    pri = cy.model.default_priors(c.y, K=c.K, L=[c.L0, c.L1], y_bounds=[-5., -3.5, -2.])
    mod = cy.model.Model(y=y, priors=pri, tau=c.tau, verbose=-1, use_stick_break=c.stick,
                         model_noisy=c.noisy, scale=c.scale, device=dev, noise='external',
                         N_global=c.N, row_global=lo, dist_group=dist.group.WORLD)
    with torch.no_grad():
        for k in c.keys:
            mod.mp[k].vp.copy_(torch.tensor(c.z['vp_' + k], dtype=F64))
        for i in range(c.I):
            mod.y_vae[i].mean.copy_(torch.tensor(c.z['vae_mean_%d' % i], dtype=F64))
            mod.y_vae[i].log_sd.copy_(torch.tensor(c.z['vae_log_sd_%d' % i], dtype=F64))
    fs = FusedStep(mod, lr=0.05, allreduce=mode)
    assert (fs.peers is not None) == (mode == 'peer')
    nz = c.noise()
    g = torch.cat([nz[k].reshape(-1) for k in c.keys]).to(dev)
    cell = torch.cat([nz['y'][i][lo[i]:hi[i]] for i in range(c.I)], 0).to(dev, torch.float32).contiguous()
    idx = [np.arange(b - a) for a, b in zip(lo, hi)]
    out = {}
    sc = fs.step(idx, global_noise=g, cell_noise=cell, update=False).tolist()
    out['scalars'] = np.array(sc)
    out['packed'] = fs.packed.cpu().numpy().copy()
    scale = -1.0 / sum(c.N)
    for k in c.keys:
        out[k] = mod.mp[k].vp.grad.cpu().numpy() / scale
    for i in range(c.I):
        out['vae_mean_%d' % i] = mod.y_vae[i].mean.grad.cpu().numpy() / scale
        out['vae_log_sd_%d' % i] = mod.y_vae[i].log_sd.grad.cpu().numpy() / scale
    for _ in range(3):      # a few real updates: the replicas must stay bit-identical
        fs.step(idx, global_noise=g, cell_noise=cell, update=True)
    out['theta'] = fs.theta.cpu().numpy().copy()
    np.savez(os.path.join(out_dir, '%s_rank%d.npz' % (mode, rank)), **out)
    torch.cuda.synchronize()
    if fs.peers is not None:
        dist.barrier()
        fs.peers.close()
    dist.destroy_process_group()


def _graph_worker(rank, world, port, out_dir):
    """multi-step CUDA graph (FusedStep.capture(repeat=3) / graph_steps) with the fused peer all-reduce
    against single-step replays: same rows, same parameters, on both ranks"""
    import cytofpy_b200 as cy
    from cytofpy_b200.model.fused import FusedStep
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    c = Case(CASE)
    lo = [n * rank // world for n in c.N]
    hi = [n * (rank + 1) // world for n in c.N]
    pri = cy.model.default_priors(c.y, K=c.K, L=[c.L0, c.L1], y_bounds=[-5., -3.5, -2.])
    out = {}
    for tag, repeat in (('multi', 3), ('single', 1)):
        torch.manual_seed(5)
        y = [t[a:b].clone() for t, a, b in zip(c.y, lo, hi)]
        mod = cy.model.Model(y=y, priors=pri, tau=c.tau, verbose=-1, use_stick_break=c.stick,
                             model_noisy=c.noisy, scale=c.scale, device=dev, noise='philox', seed=9,
                             N_global=c.N, row_global=lo, dist_group=dist.group.WORLD)
        fs = FusedStep(mod, lr=0.05, seed=9, allreduce='peer')
        assert fs.peers is not None
        rows = [fs.step([np.arange(b - a) for a, b in zip(lo, hi)]).tolist()]
        fs.capture(None, repeat=repeat)
        if repeat > 1:
            for _ in range(2):
                rows += fs.graph_steps().tolist()
        else:
            for _ in range(6):
                rows.append(fs.graph_step().tolist())
        assert fs.t == 7 and fs.peers.seq == 7
        out[tag + '_rows'] = np.array(rows)
        out[tag + '_theta'] = fs.theta.cpu().numpy().copy()
        torch.cuda.synchronize()
        dist.barrier()
        fs.peers.close()
    np.savez(os.path.join(out_dir, 'graph_rank%d.npz' % rank), **out)
    dist.destroy_process_group()


def test_multi_step_graph_with_peer_allreduce_two_ranks(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    world = 2
    mp.spawn(_graph_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(os.path.join(str(tmp_path), 'graph_rank%d.npz' % r)) for r in range(world)]
    for o in outs:
        assert np.isfinite(o['multi_rows']).all() and o['multi_rows'].shape == (7, 5)
        assert np.array_equal(o['multi_rows'], o['single_rows'])
        assert np.array_equal(o['multi_theta'], o['single_theta'])
    for k in outs[0].files:
        assert np.array_equal(outs[0][k], outs[1][k]), k


def _run(tmp_path, mode):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), mode), nprocs=world, join=True)
    return [np.load(os.path.join(str(tmp_path), '%s_rank%d.npz' % (mode, r))) for r in range(world)]


def test_fused_peer_allreduce_two_ranks(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    c = Case(CASE)
    ref = c.grads()
    outs = _run(tmp_path, 'peer')
    for o in outs:
        elbo, ll = o['scalars'][0], o['scalars'][1]
        assert np.isfinite(o['packed']).all()
        assert abs(elbo - c.out('elbo')) <= 1e-5 * abs(c.out('elbo'))
        assert abs(ll - c.out('ll')) <= 1e-5 * abs(c.out('ll'))
        for k in ref:
            assert rel(o[k], ref[k]) < 1e-4, k
    for k in outs[0].files:      # identical on both ranks, bit for bit
        assert np.array_equal(outs[0][k], outs[1][k]), k


def _stalled_worker(rank, world, port, out_dir):
    """rank 1 stalls on the host for longer than the exchange's ~2 s bound: rank 0 must give up, and BOTH ranks must
    report the failure code (no update, NaN results) instead of one of them training on alone"""
    import time
    import cytofpy_b200 as cy
    from cytofpy_b200.model.fused import FusedStep
    from cytofpy_b200.model.train import check_exchange
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    c = Case(CASE)
    lo = [n * rank // world for n in c.N]
    hi = [n * (rank + 1) // world for n in c.N]
    y = [t[a:b].clone() for t, a, b in zip(c.y, lo, hi)]
    pri = cy.model.default_priors(c.y, K=c.K, L=[c.L0, c.L1], y_bounds=[-5., -3.5, -2.])
    torch.manual_seed(5)
    mod = cy.model.Model(y=y, priors=pri, tau=c.tau, verbose=-1, use_stick_break=c.stick, model_noisy=c.noisy,
                         scale=c.scale, device=dev, noise='philox', seed=9, N_global=c.N, row_global=lo,
                         dist_group=dist.group.WORLD)
    fs = FusedStep(mod, lr=0.05, seed=9, allreduce='peer')
    idx = [np.arange(b - a) for a, b in zip(lo, hi)]
    good = fs.step(idx).tolist()                 # a normal step first: the exchange works
    torch.cuda.synchronize()
    theta1 = fs.theta.clone()
    dist.barrier()
    if rank == 1:
        time.sleep(3.5)                          # no GPU work is queued by this rank while rank 0 waits
    bad = fs.step(idx).tolist()
    torch.cuda.synchronize()
    later = fs.step(idx).tolist()                # sticky: every later step of every rank fails the same way
    raised = False
    try:
        check_exchange(later[4])
    except RuntimeError:
        raised = True
    np.savez(os.path.join(out_dir, 'stall_rank%d.npz' % rank), good=np.array(good), bad=np.array(bad),
             later=np.array(later), same=np.array([float(torch.equal(fs.theta, theta1))]), raised=np.array([float(raised)]))
    dist.barrier()
    fs.peers.close()
    dist.destroy_process_group()


def test_stalled_peer_poisons_every_rank(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    world = 2
    mp.spawn(_stalled_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        o = np.load(os.path.join(str(tmp_path), 'stall_rank%d.npz' % r))
        assert np.isfinite(o['good']).all() and o['good'][4] in (0.0, 1.0)
        for key in ('bad', 'later'):
            assert o[key][4] == 2.0, (r, key, o[key])
            assert np.isnan(o[key][0]) and np.isnan(o[key][1])
        assert o['same'][0] == 1.0          # no Adam update was applied from the failed exchange on
        assert o['raised'][0] == 1.0
