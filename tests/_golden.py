"""Helpers shared by the tests: load a golden fixture (tests/golden/*.npz, made
by oracle/gen_golden.py from the live reference) into the plain-dict form the
oracle port and the product both understand."""
import glob
import json
import os

import numpy as np
import torch

from oracle import elbo_port as port

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
F64 = torch.float64


def case_names(traj=False):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))
    return [n for n in names if n.startswith('t_') == traj]


class Case:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
        self.z = z
        self.name = name
        self.N = [int(n) for n in z['meta_N']]
        self.I = len(self.N)
        self.J, self.K = int(z['meta_J']), int(z['meta_K'])
        self.L0, self.L1 = int(z['meta_L0']), int(z['meta_L1'])
        self.tau = float(z['meta_tau'])
        self.stick, self.noisy = bool(z['meta_stick']), bool(z['meta_noisy'])
        self.scale = float(z['meta_scale'])
        self.keys = [str(k) for k in z['meta_keys']]
        self.traj = int(z['meta_traj'])
        # prior overrides in the notation of oracle/gen_golden.py ({} = reference default_priors)
        self.prior_spec = json.loads(str(z['meta_priors'])) if 'meta_priors' in z.files else {}
        self.y = [torch.tensor(z['y_%d' % i], dtype=F64) for i in range(self.I)]

    def cfg(self):
        return port.make_cfg(self.y, self.K, self.L0, self.L1, tau=self.tau,
                             use_stick_break=self.stick, model_noisy=self.noisy, scale=self.scale,
                             prior_spec=self.prior_spec)

    def state(self):
        st = {k: torch.tensor(self.z['vp_' + k], dtype=F64, requires_grad=True) for k in self.keys}
        st['vae_mean'] = [torch.tensor(self.z['vae_mean_%d' % i], dtype=F64, requires_grad=True)
                          for i in range(self.I)]
        st['vae_log_sd'] = [torch.tensor(self.z['vae_log_sd_%d' % i], dtype=F64, requires_grad=True)
                            for i in range(self.I)]
        return st

    def idx(self, t=None):
        pre = 'idx_' if t is None else 'idx_%d_' % t
        return [self.z[pre + str(i)] for i in range(self.I)]

    def noise(self, t=None):
        pre = 'noise_' if t is None else 'noise_%d_' % t
        nz = {k: torch.tensor(self.z[pre + k], dtype=F64) for k in self.keys}
        pre_y = 'noise_y_' if t is None else 'noise_y_%d_' % t
        nz['y'] = [torch.tensor(self.z[pre_y + str(i)], dtype=F64) for i in range(self.I)]
        return nz

    def out(self, k):
        return float(self.z['out_' + k])

    def grads(self):
        g = {k: self.z['grad_' + k] for k in self.keys}
        for i in range(self.I):
            g['vae_mean_%d' % i] = self.z['grad_vae_mean_%d' % i]
            g['vae_log_sd_%d' % i] = self.z['grad_vae_log_sd_%d' % i]
        return g


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))
