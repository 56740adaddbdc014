"""CPU, world_size = 2 over gloo: the cell-sharded multi-GPU path of `Model` (each rank holds
half of every sample's rows, fac_i from GLOBAL counts, one all-reduce of the packed gradient
block) must reproduce the single-process result.  The CUDA engine is replaced by the
oracle-backed stub, so this exercises exactly the host-side sharding / collective logic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests._golden import Case, rel
from tests._model_util import build_model, model_grads, model_with_engine, noise_queue, replay_noise
from tests._stub_engine import OracleEngine

import cytofpy_b200 as cy

F64 = torch.float64
CASE = 'f_fullbatch_mixed'


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard(c, rank, world):
    lo = [n * rank // world for n in c.N]
    hi = [n * (rank + 1) // world for n in c.N]
    return lo, hi


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    c = Case(CASE)
    lo, hi = _shard(c, rank, world)
    y = [t[a:b].clone() for t, a, b in zip(c.y, lo, hi)]
    pri = cy.model.default_priors(c.y, K=c.K, L=[c.L0, c.L1], y_bounds=[-5., -3.5, -2.])
    mod = model_with_engine(OracleEngine)(y=y, priors=pri, tau=c.tau, verbose=-1, use_stick_break=c.stick,
                                          model_noisy=c.noisy, scale=c.scale, device='cpu', noise='external',
                                          N_global=c.N, row_global=lo, dist_group=dist.group.WORLD)
    with torch.no_grad():
        for k in c.keys:
            mod.mp[k].vp.copy_(torch.tensor(c.z['vp_' + k], dtype=F64))
        for i in range(c.I):
            mod.y_vae[i].mean.copy_(torch.tensor(c.z['vae_mean_%d' % i], dtype=F64))
            mod.y_vae[i].log_sd.copy_(torch.tensor(c.z['vae_log_sd_%d' % i], dtype=F64))
    # the fixture is a full batch in order: this rank's cells are rows lo..hi, and its slice of the
    # recorded per-cell noise is the same rows
    q = noise_queue(c)
    nk = len(c.keys)
    q = q[:nk] + [q[nk + i][lo[i]:hi[i]] for i in range(c.I)]
    idx = [np.arange(b - a) for a, b in zip(lo, hi)]
    with replay_noise(q, 'cpu'):
        elbo, ll, lp, lq = mod(idx)
    elbo.backward()
    g = model_grads(mod, c)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), elbo=elbo.item(), ll=ll, lp=lp, lq=lq, **g)
    dist.destroy_process_group()


def test_two_ranks_reproduce_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    c = Case(CASE)
    ref = c.grads()
    outs = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(world)]
    for o in outs:      # every rank ends with the identical, complete result
        assert abs(float(o['elbo']) - c.out('elbo')) <= 1e-5 * abs(c.out('elbo'))
        assert abs(float(o['ll']) - c.out('ll')) <= 1e-5 * abs(c.out('ll'))
        for k in ref:
            assert rel(o[k], ref[k]) < 1e-4, k
    for k in ref:
        assert np.array_equal(outs[0][k], outs[1][k]), k
