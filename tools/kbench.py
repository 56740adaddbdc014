"""Times the fused ELBO kernel alone (CUDA events, L2 flushed) on BASELINE cfg shapes with
random but valid globals — the quick A/B harness for kernel tuning builds:

    CYTOF_B200_LIB=cytofpy_b200/_C/variant.so python tools/kbench.py [--workload cfg4] [--nan 0.0]
"""
import argparse
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cytofpy_b200.model.engine import Batch, CellEngine  # noqa: E402

SH = {'cfg1': (3, 2000, 32, 5, 3, 3), 'cfg2': (3, 40000, 32, 30, 5, 3), 'cfg4': (8, 125000, 32, 30, 5, 5),
      'cfg5': (8, 1250000, 40, 50, 8, 8), 'cfg5s': (8, 125000, 40, 50, 8, 8),
      'j32k50': (8, 125000, 32, 50, 5, 5), 'j64k64': (8, 125000, 64, 64, 5, 5), 'j56k40': (8, 125000, 56, 40, 5, 3), 'l88': (8, 125000, 32, 30, 8, 8), 'l75': (8, 125000, 32, 30, 7, 5), 'j24k40': (8, 125000, 24, 40, 5, 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='cfg4')
    ap.add_argument('--nan', type=float, default=0.0)
    ap.add_argument('--mb', type=int, default=0)
    ap.add_argument('--n', type=int, default=0, help='cells per sample (overrides the workload)')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--sampler', action='store_true', help='with --mb: the kernel draws the rows itself (no idx)')
    ap.add_argument('--force-missing', action='store_true', help='kernels with the imputation path even for clean data')
    a = ap.parse_args()
    I, n, J, K, L0, L1 = SH[a.workload]
    n = a.n or n
    dev = torch.device('cuda', 0)
    rng = np.random.default_rng(0)
    g = torch.Generator(device=dev).manual_seed(0)
    y = torch.randn((I * n, J), device=dev, generator=g, dtype=torch.float32) * 2.5
    if a.nan > 0:
        y[torch.rand((I * n, J), device=dev, generator=g, dtype=torch.float32) < a.nan] = float('nan')
    dirl = lambda al, size: np.log(rng.dirichlet(al, size=size))
    parts = [-np.cumsum(rng.gamma(2, .6, L0)), np.cumsum(rng.gamma(2, .6, L1)), rng.uniform(.5, 1.2, I),
             dirl(np.ones(L0), (I, J)), dirl(np.ones(L1), (I, J)), dirl(np.ones(K), I),
             rng.uniform(.005, .05, I), np.tile([-22.19, -13.47, -1.925], (I, 1)),
             rng.normal(-3, .3, (I, J)), rng.normal(-2, .3, (I, J))]
    G = torch.tensor(np.concatenate([p.ravel() for p in parts]), dtype=torch.float32, device=dev)
    Z = rng.random((J, K)) < .5
    zm = np.zeros(J, dtype=np.uint64)
    for k in range(K):
        zm |= Z[:, k].astype(np.uint64) << np.uint64(k)
    zm = torch.tensor(zm.view(np.int64), device=dev)
    eng = CellEngine.__new__(CellEngine)
    # build the engine around the already-resident matrix (avoid a host round trip)
    CellEngine.__init__(eng, [torch.zeros((1, J))] * I, dev, K, L0, L1, True, 10 ** .5, 1.0, parts[7])
    eng.y_dev = y
    eng.no_missing = a.nan == 0 and not a.force_missing
    eng.N_local = eng.N_global = [n] * I
    eng.row_base = np.arange(I + 1, dtype=np.int64) * n
    if a.mb:
        idx = torch.cat([torch.randperm(n, device=dev, generator=g)[:a.mb].to(torch.int32) for _ in range(I)])
        batch = Batch([a.mb] * I, None, None, step=1, sampler=(7, 3)) if a.sampler else Batch([a.mb] * I, idx, None, step=1)
    else:
        batch = Batch([n] * I, None, None, step=1)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ev = []
    for t in range(3 + a.iters):
        flush.fill_(t & 255)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.fwdbwd(batch, G, zm)
        e.record()
        ev.append((s, e))
    torch.cuda.synchronize()
    ms = [s.elapsed_time(e) for s, e in ev[3:]]
    cells = batch.ncells
    med = statistics.median(ms)
    blk = eng.grad_block.cpu().numpy()
    if not np.isfinite(blk).all() or not np.isfinite(eng.ll_lqy.cpu().numpy()).all():
        names = ['dmu0', 'dmu1', 'dsig', 'dle0', 'dle1', 'dlw', 'dZ', 'deps', 'dvm', 'dvls', 'dlq']
        bad = [nm for nm, a, b in zip(names, eng.b_off[:-1], eng.b_off[1:]) if not np.isfinite(blk[a:b]).all()]
        print('NON-FINITE sections:', bad, 'll_lqy', eng.ll_lqy.tolist())
    print('%s lib=%s cells=%d  median %.4f ms  min %.4f ms  %.3f Gcells/s  ll=%.6e' % (
        a.workload, os.path.basename(os.environ.get('CYTOF_B200_LIB', 'default')), cells, med, min(ms),
        cells / med / 1e6, eng.ll_lqy[0].item()))


if __name__ == '__main__':
    main()
