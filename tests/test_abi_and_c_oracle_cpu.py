"""CPU: (a) the C-ABI library loads and exports every symbol include/cytof_b200.h declares
(no compute calls without a GPU); (b) struct / layout agreement between the header, the
ctypes mirror and the oracle; (c) the plain-C oracle against the numpy oracle on the golden
fixtures; (d) the product fails loudly without a GPU / without the library."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from cytofpy_b200 import _lib
from oracle import cells_c
from oracle import cells_closed_form as ccf
from tests._golden import Case, case_names, rel
from tests.test_oracle_golden import kernel_inputs_from_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'cytof_b200.h')).read()
    declared = set(re.findall(r'\b(cytof_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.cytof_abi_version() == 2
    assert lib.cytof_strerror(0) == b'ok' and b'shape' in lib.cytof_strerror(-2)


def test_desc_struct_matches_header():
    hdr = open(os.path.join(ROOT, 'include', 'cytof_b200.h')).read()
    assert int(re.search(r'#define CYTOF_MAX_SAMPLES (\d+)', hdr).group(1)) == _lib.MAX_SAMPLES
    # 8 int32 + 2 float + 2 uint64 + 33 int64 + 32 int64 + 32 int64 + 32 double + 1 pointer
    # + the in-kernel sampler's 2 uint64 + 32 int64
    assert C.sizeof(_lib.Desc) == 8 * 4 + 2 * 4 + 2 * 8 + 33 * 8 + 3 * 32 * 8 + 8 + 2 * 8 + 32 * 8
    assert _lib.Desc.sampler_seed.offset == _lib.Desc.step_dev.offset + 8
    # cytof_peer_desc: 2 int32 + uint64 + 16 pointers + 1 pointer
    assert C.sizeof(_lib.PeerDesc) == 2 * 4 + 8 + 16 * 8 + 8
    assert _lib.Desc.cell_begin.offset == 56


@pytest.mark.parametrize('shape', [(3, 32, 30, 5, 3), (8, 32, 30, 5, 5), (8, 40, 50, 8, 8), (1, 1, 1, 1, 1)])
def test_layouts_match_oracle(shape):
    I, J, K, L0, L1 = shape
    d = _lib.Desc(I=I, J=J, K=K, L0=L0, L1=L1)
    lay = ccf.block_layout(I, J, K, L0, L1)
    names = ['dmu0', 'dmu1', 'dsig', 'dlogeta0', 'dlogeta1', 'dlogW', 'dZ', 'deps', 'dvm', 'dvls', 'dlqy_dvls']
    assert _lib.layout(d, 'block') == [lay[n][0] for n in names] + [lay['_total']]
    sizes = [L0, L1, I, I * J * L0, I * J * L1, I * K, I, 3 * I, I * J, I * J]
    assert _lib.layout(d, 'globals') == [0] + list(np.cumsum(sizes))
    # sizes quoted in SURVEY.md section 8a (+ the dlqy_dvls section this ABI adds)
    if shape == (3, 32, 30, 5, 3):
        assert lay['_total'] - I * J == 2024


@pytest.mark.parametrize('name', case_names())
def test_c_oracle_matches_numpy_oracle(name):
    c = Case(name)
    st, cfg, idx, nz = c.state(), c.cfg(), c.idx(), c.noise()
    _, g, rows, eps, fac = kernel_inputs_from_port(c, st, nz, cfg, idx)
    ll, lqy, blk, _ = ccf.cells_fwdbwd(rows, eps, g, fac, cfg['scale'], float(cfg['noisy_sd']), cfg['model_noisy'])
    y_all = np.concatenate([t.numpy() for t in c.y])
    row_base = np.concatenate([[0], np.cumsum(c.N)])[:-1]
    ll_c, lqy_c, blk_c = cells_c.cells_fwdbwd_c(y_all, row_base, [len(ix) for ix in idx], np.concatenate(idx),
                                                np.concatenate(eps), g, fac, cfg['scale'],
                                                float(cfg['noisy_sd']), cfg['model_noisy'])
    assert abs(ll_c - ll) <= 1e-11 * abs(ll) and abs(lqy_c - lqy) <= 1e-11 * max(1, abs(lqy))
    assert rel(blk_c, blk) < 1e-10
    assert abs(ll_c - c.out('ll')) <= 1e-10 * abs(c.out('ll'))


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_product_fails_loudly_without_gpu():
    import cytofpy_b200 as cy
    y = cy.util.simdata(N=[20, 10], J=4, a_W=[1., 1.], L0=2, L1=2, seed=0)['data']['y']
    pri = cy.model.default_priors(y, K=3, L=[2, 2], y_bounds=[-5., -3.5, -2.])
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        cy.model.Model(y=y, priors=pri, verbose=-1, device='cpu')
    with pytest.raises((RuntimeError, AssertionError)):
        cy.model.Model(y=y, priors=pri, verbose=-1)        # default device is CUDA


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib._build, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU / PyTorch fallback'):
        _lib.load()
