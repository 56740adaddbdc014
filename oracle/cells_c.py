"""TEST INFRASTRUCTURE — ctypes wrapper of oracle/cells_oracle.c (fp64, plain C)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, '_build', 'libcells_oracle.so')


def build():
    src = os.path.join(HERE, 'cells_oracle.c')
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-o', SO, src, '-lm'])
    return SO


def cells_fwdbwd_c(y_all, row_base, counts, idx, eps, g, fac, scale, noisy_sd, model_noisy):
    """y_all [rows,J] fp64 (NaN = missing); idx concatenated sample-local rows or None;
    eps [C,J] or None; g as in oracle/cells_closed_form.cells_fwdbwd.  Returns ll, lqy, block."""
    lib = C.CDLL(build())
    I = len(counts)
    J, K = g['Z'].shape
    L0, L1 = len(g['mu0']), len(g['mu1'])
    total = L0 + L1 + I + I * J * L0 + I * J * L1 + I * K + J * K + I + 3 * I * J
    dbl = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    i64 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int64))
    P = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)
    cb = i64(np.concatenate([[0], np.cumsum(counts)]))
    arrs = [cb, i64(row_base), dbl(fac), dbl(y_all), None if idx is None else i64(idx),
            None if eps is None else dbl(eps), dbl(g['mu0']), dbl(g['mu1']), dbl(g['sig']),
            dbl(g['logeta0']), dbl(g['logeta1']), dbl(g['logW']), dbl(g['Z']),
            dbl(g['eps'] if model_noisy else np.zeros(I)), dbl(g['b']), dbl(g['vm']), dbl(g['vls'])]
    block, out = np.zeros(total), np.zeros(2)
    lib.cells_oracle_fwdbwd.restype = C.c_int
    rc = lib.cells_oracle_fwdbwd(C.c_int(I), C.c_int(J), C.c_int(K), C.c_int(L0), C.c_int(L1),
                                 C.c_int(int(model_noisy)), C.c_double(noisy_sd), C.c_double(scale),
                                 *[P(a) for a in arrs], P(block), P(out))
    assert rc == 0, rc
    return float(out[0]), float(out[1]), block
