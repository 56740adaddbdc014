"""Steps per second of the reference-facing `cytofpy_b200.model.fit` loop itself (host bookkeeping,
read-backs, snapshots included) on cfg2-like shapes: marginal time per step between two run lengths,
so model construction and graph capture cancel.

    python tools/fitbench.py [minibatch_size]
"""
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cytofpy_b200 as cy  # noqa: E402

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
torch.manual_seed(0)
np.random.seed(0)
s = cy.util.synth_cytof([40000] * 3, 32, [100.] * 5, 5, 3, seed=0)
y = [torch.from_numpy(a) for a in s['y']]
pri = cy.model.default_priors([t[:2000].double() for t in y], K=30, L=[5, 3], y_bounds=[-5., -3.5, -2.])


def run(n, pipeline):
    with contextlib.redirect_stdout(io.StringIO()):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = cy.model.fit(y, max_iter=n, lr=1e-2, print_freq=10 ** 9, eps_conv=0, priors=pri, minibatch_size=mb,
                           verbose=0, seed=1, pipeline=pipeline, flush=False)
        torch.cuda.synchronize()
    return time.perf_counter() - t0, out['elbo'][-1]


run(200, True)
for pipeline in (False, True, False, True):
    a, _ = run(2000, pipeline)
    b, e = run(12000, pipeline)
    print('pipeline %-5s  %.1f us/step (marginal, 10,000 steps)  last elbo %.3f' % (pipeline, (b - a) / 1e4 * 1e6, e))
