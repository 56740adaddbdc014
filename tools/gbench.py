"""Times the stages of one fused step separately (global forward | per-cell kernel + finalise |
global backward + Adam), each as a CUDA graph of REP back-to-back calls on one stream, warm L2:

    python tools/gbench.py [--workload cfg4] [--cells 6000]

The global part does not depend on the number of cells, so a small --cells isolates it."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cytofpy_b200 as cy  # noqa: E402
from cytofpy_b200 import _lib  # noqa: E402
from cytofpy_b200.model.engine import Batch, _ptr  # noqa: E402
from cytofpy_b200.model.fused import FusedStep  # noqa: E402

REP = 50


def timed_graph(fn, dev):
    s = torch.cuda.Stream(dev)
    with torch.cuda.stream(s):
        fn()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(REP):
                fn()
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s)
            g.replay()
            b.record(s)
            b.synchronize()
            best = min(best, a.elapsed_time(b) / REP)
    return best * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='cfg4')
    ap.add_argument('--cells', type=int, default=2000)
    a = ap.parse_args()
    w = dict(bench.WORKLOADS[a.workload])
    w['n'] = a.cells
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    I, n = w['I'], w['n']
    y = [torch.from_numpy(x) for x in bench.make_data(w, seed=1000)]
    priors = cy.model.default_priors([t[:2000].double() for t in y], K=w['K'], L=[w['L0'], w['L1']],
                                     y_bounds=[-5., -3.5, -2.])
    torch.manual_seed(1)
    model = cy.model.Model(y=y, priors=priors, verbose=-1, device=dev, noise='philox', seed=1, **w['kw'])
    fs = FusedStep(model, lr=1e-2, seed=1)
    fs.step([range(n)] * I)
    torch.cuda.synchronize()
    eng, g, lib = fs.eng, fs.g, fs.lib
    batch = Batch([n] * I, None, None, step=2)
    counts = model._counts_global(batch)

    def stream():
        return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def gfwd():
        _lib.check(lib.cytof_global_fwd_f64(C.byref(g), _ptr(fs.theta), _ptr(None), _ptr(fs.stash), _ptr(fs.globals),
                                            _ptr(fs.zmask), _ptr(fs.scalars_in), stream()))

    def cells():
        eng.fwdbwd_packed(batch, fs.globals, fs.zmask, fs.packed, counts)

    def gbwd():
        _lib.check(lib.cytof_global_bwd_adam_f64(C.byref(g), _ptr(fs.theta), _ptr(fs.stash), _ptr(fs.packed),
                                                 _ptr(fs.scalars_in), _ptr(fs.adam_m), _ptr(fs.adam_v), _ptr(fs.grad),
                                                 _ptr(fs.scalars), _ptr(None), _ptr(fs.flag), C.c_int32(0), stream()))

    def whole():
        gfwd(); cells(); gbwd()

    for name, fn in (('global_fwd', gfwd), ('cells+finalize', cells), ('global_bwd+adam', gbwd), ('whole step', whole)):
        print('%-18s %7.2f us' % (name, timed_graph(fn, dev)))


if __name__ == '__main__':
    main()
