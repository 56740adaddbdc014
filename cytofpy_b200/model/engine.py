"""Host side of the C ABI for the per-cell ELBO: owns the device-resident data
matrix, the workspace and output buffers, builds `cytof_desc` and enqueues the
fused kernel on torch's current CUDA stream.  torch is used for device memory
and streams only; all arithmetic over cells happens in libcytof_b200."""
import ctypes as C

import numpy as np
import torch

from .. import _lib

F64 = torch.float64


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Batch:
    """The cells of one step: per-sample counts, (optional) gathered row indices
    on the device and (optional) external noise."""
    __slots__ = ('counts', 'idx', 'eps', 'step', 'sampler')

    def __init__(self, counts, idx=None, eps=None, step=0, sampler=None):
        """`sampler` = (seed, step): the per-cell kernel draws the minibatch rows itself
        (CYTOF_FLAG_SAMPLE_IN_KERNEL) -- the rows cytof_sample_minibatch would write for that seed and step;
        `idx` must then be None."""
        self.counts, self.idx, self.eps, self.step = [int(c) for c in counts], idx, eps, int(step)
        self.sampler = sampler
        if sampler is not None and idx is not None:
            raise ValueError('Batch: either explicit row indices or the in-kernel sampler, not both')

    @property
    def ncells(self):
        return sum(self.counts)


class CellEngine:
    def __init__(self, y, device, K, L0, L1, model_noisy, noisy_sd, scale, b,
                 N_global=None, row_global=None, seed=0):
        """y: list of I (N_i, J) tensors (any float dtype, NaN = missing) — the rows THIS rank
        holds.  N_global: per-sample global cell counts (defaults to the local ones);
        row_global: global id of the first local row of each sample (multi-GPU sharding)."""
        self.lib = _lib.load()
        self.step_dev = None      # device pointer of a step counter (CUDA-graph replay), or None
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('cytofpy_b200: the ELBO hot path needs a CUDA device (got %s); '
                               'there is no CPU fallback' % self.device)
        self.I, self.J = len(y), int(y[0].shape[1])
        self.K, self.L0, self.L1 = int(K), int(L0), int(L1)
        if self.I > _lib.MAX_SAMPLES or self.J > _lib.MAX_J or self.K > _lib.MAX_K or \
                max(self.L0, self.L1) > _lib.MAX_L:
            raise ValueError('cytofpy_b200: sizes beyond compiled limits (I<=32, J<=64, K<=64, L<=8)')
        self.model_noisy, self.noisy_sd, self.scale = bool(model_noisy), float(noisy_sd), float(scale)
        self.N_local = [int(t.shape[0]) for t in y]
        self.N_global = [int(n) for n in (N_global if N_global is not None else self.N_local)]
        self.row_global = [int(r) for r in (row_global if row_global is not None else [0] * self.I)]
        self.row_base = np.concatenate([[0], np.cumsum(self.N_local)]).astype(np.int64)
        self.seed = int(seed)
        # device-resident data: [sum N_i, J] fp32, NaN = missing
        self.y_dev = torch.cat([t.to(torch.float32) for t in y], 0).contiguous().to(self.device)
        # the reference derives m = isnan(y) once (fit.py:66); an all-false mask lets the library
        # use the kernels without the imputation path (cytof_desc.flags)
        self.no_missing = not any(bool(torch.isnan(t).any()) for t in y)
        self.b_dev = torch.as_tensor(b, dtype=F64).reshape(self.I, 3).to(self.device)
        d = self._desc(Batch(self.N_local))
        self.g_off = _lib.layout(d, 'globals')
        self.b_off = _lib.layout(d, 'block')
        self.ws_bytes = int(self.lib.cytof_workspace_bytes(C.byref(d)))
        if self.ws_bytes == 0:
            raise RuntimeError('cytofpy_b200: descriptor rejected by the CUDA library')
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.grad_block = torch.zeros(self.b_off[-1], dtype=torch.float32, device=self.device)
        self.ll_lqy = torch.zeros(2, dtype=F64, device=self.device)
        self.launches = 0

    def _desc(self, batch, noise_mode=None):
        d = _lib.Desc()
        d.I, d.J, d.K, d.L0, d.L1 = self.I, self.J, self.K, self.L0, self.L1
        d.model_noisy = int(self.model_noisy)
        if noise_mode is None:
            noise_mode = _lib.NOISE_EXTERNAL if batch.eps is not None else _lib.NOISE_PHILOX
        d.noise_mode = noise_mode
        d.noisy_sd, d.scale = self.noisy_sd, self.scale
        d.flags = _lib.FLAG_NO_MISSING if self.no_missing else 0
        if batch.sampler is not None:
            d.flags |= _lib.FLAG_SAMPLE_IN_KERNEL
            d.sampler_seed, d.sampler_step = int(batch.sampler[0]) & (2 ** 64 - 1), int(batch.sampler[1])
            for i in range(self.I):
                d.sampler_N[i] = int(self.N_local[i])
        d.seed, d.step = self.seed & (2 ** 64 - 1), batch.step
        d.step_dev = self.step_dev
        cb = 0
        for i in range(self.I):
            d.cell_begin[i] = cb
            cb += batch.counts[i]
            d.row_base[i] = int(self.row_base[i])
            d.row_global[i] = self.row_global[i]
        d.cell_begin[self.I] = cb
        return d

    def set_fac(self, d, batch, counts_global=None):
        """fac_i = N_i / B_i with GLOBAL counts (reference Model.py:257)."""
        cg = counts_global if counts_global is not None else batch.counts
        for i in range(self.I):
            d.fac[i] = self.N_global[i] / cg[i] if (cg[i] > 0 and batch.counts[i] > 0) else 0.0

    def make_idx(self, idx_list):
        """list of per-sample index collections (numpy / list / range / device tensor) ->
        (counts, int32 device tensor or None when every sample is taken whole, in order)."""
        counts = [len(ix) for ix in idx_list]
        if all(isinstance(ix, range) and ix == range(n) for ix, n in zip(idx_list, self.N_local)):
            return counts, None
        # host-side indices are range-checked like the reference's y[idx] (IndexError; negative values
        # wrap once, numpy style): the kernel's row gather itself is unchecked
        host_parts = {}
        for q, (ix, n) in enumerate(zip(idx_list, self.N_local)):
            if torch.is_tensor(ix) and ix.is_cuda:
                continue
            a = np.asarray(ix.cpu() if torch.is_tensor(ix) else ix).reshape(-1)
            if a.size:
                if not np.issubdtype(a.dtype, np.integer):
                    raise IndexError('cytofpy_b200: minibatch indices must be integers')
                a = np.where(a < 0, a + n, a)
                if a.min() < 0 or a.max() >= n:
                    raise IndexError('cytofpy_b200: index out of range for sample %d with %d cells' % (q, n))
            host_parts[q] = a.astype(np.int32)
        if len(host_parts) == len(idx_list):
            host = np.concatenate([host_parts[q] for q in range(len(idx_list))]) if sum(counts) else np.zeros(0, np.int32)
            return counts, torch.from_numpy(host).to(self.device, non_blocking=True)
        parts = []
        for q, ix in enumerate(idx_list):
            if q in host_parts:
                parts.append(torch.from_numpy(host_parts[q]).to(self.device))
            else:      # device tensors (the device sampler's output) are trusted: checking would cost a sync
                parts.append(ix.to(device=self.device, dtype=torch.int32).reshape(-1))
        return counts, torch.cat(parts).contiguous()

    def fwdbwd(self, batch, globals_f32, zmask_i64, counts_global=None):
        """Enqueue the fused kernel; results land in self.grad_block / self.ll_lqy."""
        d = self._desc(batch)
        self.set_fac(d, batch, counts_global)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.cytof_elbo_fwdbwd_f32(
                C.byref(d), _ptr(self.y_dev), _ptr(batch.idx), _ptr(batch.eps), _ptr(globals_f32),
                _ptr(zmask_i64), _ptr(self.grad_block), _ptr(self.ll_lqy), _ptr(self.workspace),
                C.c_size_t(self.ws_bytes), C.c_void_p(stream))
        _lib.check(rc)
        self.launches += 2
        return self.grad_block, self.ll_lqy

    def full_batch(self):
        return Batch(self.N_local)

    def fwdbwd_allreduce(self, batch, globals_f32, zmask_i64, packed_f64, counts_global, peers):
        """fwdbwd_packed with the sum over ranks fused into the finalise kernel (peer-memory
        mailboxes, model/peer.py): `packed_f64` receives the all-reduced vector; no NCCL call."""
        d = self._desc(batch)
        self.set_fac(d, batch, counts_global)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.cytof_elbo_fwdbwd_allreduce(
                C.byref(d), _ptr(self.y_dev), _ptr(batch.idx), _ptr(batch.eps), _ptr(globals_f32),
                _ptr(zmask_i64), _ptr(packed_f64), _ptr(self.workspace), C.c_size_t(self.ws_bytes),
                C.byref(peers.next()), C.c_void_p(stream))
        _lib.check(rc)
        self.launches += 2
        return packed_f64

    def fwdbwd_packed(self, batch, globals_f32, zmask_i64, packed_f64, counts_global=None):
        """Same as fwdbwd, results as one fp64 vector [grad block | ll | log_qy] in `packed_f64`."""
        d = self._desc(batch)
        self.set_fac(d, batch, counts_global)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.cytof_elbo_fwdbwd_packed(
                C.byref(d), _ptr(self.y_dev), _ptr(batch.idx), _ptr(batch.eps), _ptr(globals_f32),
                _ptr(zmask_i64), _ptr(packed_f64), _ptr(self.workspace), C.c_size_t(self.ws_bytes),
                C.c_void_p(stream))
        _lib.check(rc)
        self.launches += 2
        return packed_f64

    def impute(self, batch, globals_f32):
        """Imputed minibatch [C, J] fp32 (reference vae.py:33)."""
        d = self._desc(batch)
        out = torch.empty((batch.ncells, self.J), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.cytof_impute_emit_f32(C.byref(d), _ptr(self.y_dev), _ptr(batch.idx),
                                                _ptr(batch.eps), _ptr(globals_f32), _ptr(out),
                                                C.c_void_p(stream))
        _lib.check(rc)
        self.launches += 1
        return out

    def philox_noise(self, batch):
        """The draws CYTOF_NOISE_PHILOX would use for this batch, [C, J] fp32 (test hook)."""
        d = self._desc(batch, noise_mode=_lib.NOISE_PHILOX)
        out = torch.empty((batch.ncells, self.J), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            rc = self.lib.cytof_noise_emit_f32(C.byref(d), _ptr(batch.idx), _ptr(out), C.c_void_p(stream))
        _lib.check(rc)
        return out
