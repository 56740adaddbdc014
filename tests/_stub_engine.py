"""A CPU stand-in for cytofpy_b200.model.engine.CellEngine that evaluates the
per-cell part with the fp64 oracle (oracle/cells_closed_form.py).  TEST ONLY:
it lets the host-side logic of Model / update / the multi-rank all-reduce be
exercised without a GPU.  The product never uses it."""
import numpy as np
import torch

from oracle import cells_closed_form as ccf

F64 = torch.float64


class OracleEngine:
    def __init__(self, y, device, K, L0, L1, model_noisy, noisy_sd, scale, b,
                 N_global=None, row_global=None, seed=0):
        self.device = torch.device(device)
        self.I, self.J, self.K, self.L0, self.L1 = len(y), int(y[0].shape[1]), K, L0, L1
        self.model_noisy, self.noisy_sd, self.scale = model_noisy, noisy_sd, scale
        self.N_local = [int(t.shape[0]) for t in y]
        self.N_global = list(N_global) if N_global is not None else self.N_local
        self.y = [t.double().numpy() for t in y]
        self.b_dev = torch.as_tensor(b, dtype=F64).reshape(self.I, 3)
        lay = ccf.block_layout(self.I, self.J, K, L0, L1)
        names = ['dmu0', 'dmu1', 'dsig', 'dlogeta0', 'dlogeta1', 'dlogW', 'dZ', 'deps', 'dvm', 'dvls', 'dlqy_dvls']
        self.b_off = [lay[n][0] for n in names] + [lay['_total']]
        sizes = [L0, L1, self.I, self.I * self.J * L0, self.I * self.J * L1, self.I * K, self.I,
                 self.I * 3, self.I * self.J, self.I * self.J]
        self.g_off = [0] + list(np.cumsum(sizes))
        self.launches = 0

    def make_idx(self, idx_list):
        counts = [len(ix) for ix in idx_list]
        cat = np.concatenate([np.asarray(ix, dtype=np.int64).reshape(-1) for ix in idx_list]) \
            if sum(counts) else np.zeros(0, np.int64)
        return counts, torch.from_numpy(cat)

    def fwdbwd(self, batch, globals_f32, zmask, counts_global=None):
        o, I, J, K, L0, L1 = self.g_off, self.I, self.J, self.K, self.L0, self.L1
        G = globals_f32.double().numpy()
        g = {'mu0': G[o[0]:o[1]], 'mu1': G[o[1]:o[2]], 'sig': G[o[2]:o[3]],
             'logeta0': G[o[3]:o[4]].reshape(I, J, L0), 'logeta1': G[o[4]:o[5]].reshape(I, J, L1),
             'logW': G[o[5]:o[6]].reshape(I, K), 'eps': G[o[6]:o[7]], 'b': G[o[7]:o[8]].reshape(I, 3),
             'vm': G[o[8]:o[9]].reshape(I, J), 'vls': G[o[9]:o[10]].reshape(I, J)}
        zm = zmask.numpy().astype(np.uint64)
        g['Z'] = ((zm[:, None] >> np.arange(K, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.float64)
        idx = batch.idx.numpy()
        rows, eps, start = [], [], 0
        for i, c in enumerate(batch.counts):
            rows.append(self.y[i][idx[start:start + c]])
            eps.append(batch.eps[start:start + c].double().numpy() if batch.eps is not None
                       else np.zeros((c, J)))
            start += c
        cg = counts_global if counts_global is not None else batch.counts
        fac = np.array([self.N_global[i] / cg[i] if cg[i] else 0.0 for i in range(I)])
        ll, lqy, blk, _ = ccf.cells_fwdbwd(rows, eps, g, fac, self.scale, self.noisy_sd, self.model_noisy)
        self.launches += 2
        return torch.tensor(blk, dtype=torch.float32), torch.tensor([ll, lqy], dtype=F64)
