"""The whole ADVI step as five kernel launches behind the C ABI (global fwd 1, per-cell 2, global bwd 1, Adam 1):

    [cytof_sample_minibatch_multi]  ->  cytof_global_fwd_f64  ->  cytof_elbo_fwdbwd_allreduce / _packed (2 kernels;
        with cells sharded over GPUs the finalise kernel also sums the packed gradient block over the ranks
        through peer memory, or NCCL does)  ->  cytof_global_bwd_adam_f64

replacing one reference `update` (fit.py:32-48: Model.forward, backward, NaN scrub, Adam),
which dispatches ~5,000 ATen ops and syncs the host 4 times.  All variational parameters live
in ONE flat fp64 device buffer `theta`; the Parameters of `model.mp[key].vp` and of
`model.y_vae[i]` are re-pointed to views of it, so every reference-style accessor
(`mod.mp['W'].dist()`, deepcopy snapshots, pickling) keeps working while training runs fused.
"""
import ctypes as C

import numpy as np
import torch
from torch.distributions import Beta, Dirichlet, Gamma, Uniform
from torch.distributions.log_normal import LogNormal

from .. import _lib
from .engine import Batch, _ptr

F64 = torch.float64


def recognise_priors(priors, model_noisy):
    """Hyper-parameters of the prior families the fused kernels implement, or None if `priors`
    holds anything else (callers may put arbitrary distributions in the dict, reference
    sims/cb/cb.py:117-123) — the caller then uses the generic torch path."""
    f = lambda t: float(torch.as_tensor(t).reshape(-1)[0])
    out = {}
    try:
        for k in ('delta0', 'delta1', 'alpha'):
            d = priors[k]
            if not isinstance(d, Gamma):
                return None
            out[k] = (f(d.concentration), f(d.rate))
        d = priors['sig2']
        if isinstance(d, LogNormal):
            out['sig2'], out['sig2_family'] = (f(d.loc), f(d.scale)), 0
        elif isinstance(d, Gamma):
            out['sig2'], out['sig2_family'] = (f(d.concentration), f(d.rate)), 1
        else:
            return None
        for k in ('eta0', 'eta1', 'W'):
            d = priors[k]
            if not isinstance(d, Dirichlet):
                return None
            out[k] = [float(v) for v in d.concentration.reshape(-1)]
        d = priors['H']
        if not (isinstance(d, Uniform) and f(d.low) == 0.0 and f(d.high) == 1.0):
            return None
        if model_noisy:
            d = priors['eps']
            if not isinstance(d, Beta):
                return None
            out['eps'] = (f(d.concentration1), f(d.concentration0))
        else:
            out['eps'] = (1.0, 1.0)
    except (KeyError, AttributeError):
        return None
    return out


class FusedStep:
    IN_KERNEL_SAMPLER_MAX_CELLS = 32768     # see capture()
    def __init__(self, model, lr=1e-1, betas=(0.9, 0.999), adam_eps=1e-8, seed=0, allreduce='peer'):
        hyp = recognise_priors(model.priors, model.model_noisy)
        if hyp is None:
            raise ValueError('priors outside the families the fused global kernels implement')
        self.model, self.eng, self.dev = model, model.engine, model.device
        self.lib = _lib.load()
        I, J, K, L = model.I, model.J, model.K, model.L
        if len(hyp['eta0']) != L[0] or len(hyp['eta1']) != L[1] or len(hyp['W']) != K:
            raise ValueError('Dirichlet prior sizes do not match L / K')
        g = _lib.GlobalDesc()
        g.I, g.J, g.K, g.L0, g.L1 = I, J, K, L[0], L[1]
        g.model_noisy, g.use_stick_break = int(model.model_noisy), int(model.use_stick_break)
        g.noise_mode, g.sig2_family = _lib.NOISE_PHILOX, hyp['sig2_family']
        g.tau = float(model.tau)
        for name in ('delta0', 'delta1', 'sig2', 'alpha', 'eps'):
            getattr(g, name)[0], getattr(g, name)[1] = hyp[name]
        for i, v in enumerate(hyp['eta0']):
            g.eta0_conc[i] = v
        for i, v in enumerate(hyp['eta1']):
            g.eta1_conc[i] = v
        for i, v in enumerate(hyp['W']):
            g.W_conc[i] = v
        g.seed = int(seed) & (2 ** 64 - 1)
        g.grad_scale = -1.0 / float(model.Nsum)
        g.lr, g.beta1, g.beta2, g.adam_eps = float(lr), float(betas[0]), float(betas[1]), float(adam_eps)
        self.g = g
        sizes = (C.c_int64 * 4)()
        _lib.check(self.lib.cytof_global_sizes(C.byref(g), sizes))
        self.R, self.n_scrub, self.P, stash_len = [int(v) for v in sizes]
        dev = self.dev
        z = lambda n, dt=F64: torch.zeros(n, dtype=dt, device=dev)
        self.theta, self.adam_m, self.adam_v, self.grad = z(self.P), z(self.P), z(self.P), z(self.P)
        self.stash = z(stash_len)
        self.packed = z(self.eng.b_off[-1] + 2)
        # launches of the global part (csrc/global_part.cu): the scans run in the last CTA of the element
        # kernels when their operands fit in shared memory (200 KB), else as kernels of their own
        smem = 200 * 1024
        self._fwd_launches = 1 if 16 * self.R <= smem else 2
        self._global_launches = self._fwd_launches + (2 if 8 * (2 * self.R + self.packed.numel()) <= smem else 3)
        self.scalars_in, self.scalars = z(2), z(5)
        self.globals = z(self.eng.g_off[-1], torch.float32)
        o = self.eng.g_off
        self.globals[o[7]:o[8]] = self.eng.b_dev.reshape(-1).to(torch.float32)   # b section: constants
        self.zmask = z(J, torch.int64)
        self.flag = z(1, torch.int32)
        self.step_dev = z(1, torch.int64)
        self.t = 0
        self.launches = 0
        # multi-GPU: 'peer' (default) = one-shot all-reduce over peer memory fused into the finalise
        # kernel; 'nccl' = separate torch.distributed all-reduce (kept for A/B measurements)
        self.graph = None
        self.peers = None
        if model.dist_group is not None and allreduce in ('peer', 'auto'):
            from .peer import PeerMailboxes
            try:
                self.peers = PeerMailboxes(self.eng, model.dist_group)
            except RuntimeError:
                if allreduce == 'peer':
                    raise
                self.peers = None      # 'auto': every rank raised together (peer.py) -> NCCL all-reduce on all ranks
        self._bind()

    def _cells_and_tail(self, batch, cg, update, stream, step_dev=None):
        """per-cell kernel -> packed (summed over ranks) -> backward of the global part -> Adam"""
        m, eng, g = self.model, self.eng, self.g
        if self.peers is not None:      # sum over ranks fused into the finalise kernel (NVLink P2P)
            eng.fwdbwd_allreduce(batch, self.globals, self.zmask, self.packed, cg, self.peers)
        else:
            eng.fwdbwd_packed(batch, self.globals, self.zmask, self.packed, cg)
            if m.dist_group is not None:
                torch.distributed.all_reduce(self.packed, group=m.dist_group)
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.cytof_global_bwd_adam_f64(
                C.byref(g), _ptr(self.theta), _ptr(self.stash), _ptr(self.packed), _ptr(self.scalars_in),
                _ptr(self.adam_m), _ptr(self.adam_v), _ptr(self.grad), _ptr(self.scalars),
                _ptr(step_dev), _ptr(self.flag), C.c_int32(1 if update else 0), stream))
        return 2 + self._global_launches - self._fwd_launches

    def _bind(self):
        """copy the model's current parameters into theta and re-point them to views of it"""
        m, off = self.model, 0
        with torch.no_grad():
            for key in m.mp:
                vp = m.mp[key].vp
                n = vp.numel()
                self.theta[off:off + n] = vp.detach().reshape(-1).to(self.dev)
                vp.data = self.theta[off:off + n].view(vp.shape)
                vp.grad = self.grad[off:off + n].view(vp.shape)
                off += n
            assert off == self.n_scrub, (off, self.n_scrub)
            for vae in m.y_vae:
                for q in (vae.mean, vae.log_sd):
                    n = q.numel()
                    self.theta[off:off + n] = q.detach().reshape(-1).to(self.dev)
                    q.data = self.theta[off:off + n].view(q.shape)
                    q.grad = self.grad[off:off + n].view(q.shape)
                    off += n
            assert off == self.P

    def step(self, idx, global_noise=None, cell_noise=None, update=True):
        """One fused ELBO forward + backward (+ Adam when update).  idx as for Model.forward.
        global_noise (R,) / cell_noise (C, J): externally supplied standard normals (parity
        tests); by default both are drawn in-kernel with Philox keyed by (seed, step).
        Returns the device tensor [elbo, ll, lp, lq, scrubbed] (no host sync)."""
        m, eng, g = self.model, self.eng, self.g
        self.t += 1
        g.step = self.t
        g.adam_step = self.t
        g.noise_mode = _lib.NOISE_EXTERNAL if global_noise is not None else _lib.NOISE_PHILOX
        counts, idx_dev = eng.make_idx(idx)
        batch = Batch(counts, idx_dev, cell_noise, step=self.t)
        stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.cytof_global_fwd_f64(
                C.byref(g), _ptr(self.theta), _ptr(global_noise), _ptr(self.stash), _ptr(self.globals),
                _ptr(self.zmask), _ptr(self.scalars_in), stream))
        self.launches += self._fwd_launches + self._cells_and_tail(batch, m._counts_global(batch), update, stream)
        assert self.peers is None or self.peers.seq == self.t, 'peer sequence and step count must run in lock step'
        return self.scalars

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, minibatch_size=None, repeat=1):
        """Capture one whole training step -- device minibatch sampler, global forward, fused
        per-cell kernel + finalise (+ all-reduce), global backward + Adam -- in ONE CUDA graph.
        All step-dependent inputs (Philox counters of the global and per-cell noise, sampler key,
        Adam step, peer sequence number) are read on the device from `step_dev`, which the last
        kernel of the step bumps, so `graph_step()` is a bare replay (reference fit.py:91-106).
        `repeat` > 1 additionally captures `repeat` consecutive steps as one graph, each followed by a
        copy of its [elbo, ll, lp, lq, scrubbed] into row j of `self.hist`: `graph_steps()` then costs the
        host one launch and one read-back per `repeat` steps."""
        m, eng, g = self.model, self.eng, self.g
        I = eng.I
        counts = [n if (minibatch_size is None or minibatch_size >= n) else int(minibatch_size)
                  for n in eng.N_local]
        full = all(c == n for c, n in zip(counts, eng.N_local))
        if self.t == 0:      # one ordinary step first: module loading / attribute calls stay out of the capture
            from .train import device_minibatch
            self.step([range(n) for n in eng.N_local] if full else
                      device_minibatch(m, minibatch_size, 0, self.g.seed))
        # small minibatches: the per-cell kernel evaluates the sampler's keyed bijection itself
        # (CYTOF_FLAG_SAMPLE_IN_KERNEL, step read from step_dev): no sampler node and no index buffer in the graph
        # (cfg2: 55.2 -> 53.2 us per step).  The cycle-walking bijection diverges inside a warp, so for LARGE gathered
        # batches the separate, fully parallel sampler kernel wins (8 x 100 k cells: kernel 0.269 ms with idx, 0.310 ms
        # sampling in the kernel); the crossover is ~40 k cells.  CYTOF_SAMPLER_KERNEL=1 forces the sampler launch (A/B;
        # the rows are the same either way)
        import os
        in_kernel = (not full and sum(counts) <= self.IN_KERNEL_SAMPLER_MAX_CELLS
                     and os.environ.get('CYTOF_SAMPLER_KERNEL', '0') != '1')
        idx_buf = None if (full or in_kernel) else torch.empty(sum(counts), dtype=torch.int32, device=self.dev)
        batch = Batch(counts, idx_buf, None, step=0, sampler=(self.g.seed, 0) if in_kernel else None)
        cg = m._counts_global(batch)                     # (cached) collective outside the capture
        self.step_dev.fill_(self.t)
        sd = self.step_dev.data_ptr()
        g.step_dev, eng.step_dev, g.noise_mode = sd, sd, _lib.NOISE_PHILOX
        if self.peers is not None:
            assert self.peers.seq == self.t, 'peer sequence and step count must run in lock step'
            self.peers.desc.seq_dev = sd
        n_arr = (C.c_int64 * I)(*[int(n) for n in eng.N_local])
        m_arr = (C.c_int64 * I)(*[int(c) for c in counts])
        seed = C.c_uint64(self.g.seed)

        def body():
            stream = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
            with torch.cuda.device(self.dev):
                if idx_buf is not None:      # every sample's minibatch in one launch
                    _lib.check(self.lib.cytof_sample_minibatch_multi(
                        C.c_int32(I), n_arr, m_arr, seed, C.c_uint64(0), C.c_void_p(sd),
                        C.c_void_p(idx_buf.data_ptr()), stream))
                _lib.check(self.lib.cytof_global_fwd_f64(
                    C.byref(g), _ptr(self.theta), _ptr(None), _ptr(self.stash), _ptr(self.globals),
                    _ptr(self.zmask), _ptr(self.scalars_in), stream))
            tail_launches.append(self._cells_and_tail(batch, cg, True, stream, step_dev=self.step_dev))
            if self.peers is not None:
                self.peers.seq -= 1      # bookkeeping happens in graph_step

        tail_launches = []
        self._graph_launches = 0
        self._graph_idx = idx_buf
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            body()
        self._graph_launches = self._fwd_launches + tail_launches[0] + (1 if idx_buf is not None else 0)
        self.graph_multi, self._repeat = None, int(repeat)
        if self._repeat > 1:
            self.hist = torch.zeros((self._repeat, 5), dtype=F64, device=self.dev)
            self.graph_multi = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_multi):
                for j in range(self._repeat):
                    body()
                    self.hist[j].copy_(self.scalars)
        return self

    def graph_steps(self):
        """Replay the `repeat` captured consecutive steps; returns the device tensor [repeat, 5] of their
        [elbo, ll, lp, lq, scrubbed] rows."""
        self.graph_multi.replay()
        self.t += self._repeat
        if self.peers is not None:
            self.peers.seq += self._repeat
        self.launches += self._repeat * (self._graph_launches + 1)
        return self.hist

    def graph_step(self):
        """Replay the captured step; returns the device tensor [elbo, ll, lp, lq, scrubbed]."""
        self.graph.replay()
        self.t += 1
        if self.peers is not None:
            self.peers.seq += 1
        self.launches += self._graph_launches
        return self.scalars

    def release_graph(self):
        """back to host-driven steps (`step`)"""
        self.graph = None
        self.graph_multi = None
        self.g.step_dev = None
        self.eng.step_dev = None
        if self.peers is not None:
            self.peers.desc.seq_dev = None

    def use_host_matrix(self, y_pinned):
        """Leave the data matrix with the caller: `y_pinned` ([sum N_i, J] fp32, PINNED host memory, NaN =
        missing, same row layout as the resident copy) becomes the matrix the per-cell kernel reads.  A
        pinned allocation is addressable from the device (unified addressing), so the kernel's row gather
        (cp.async, 128-byte rows) pulls exactly the minibatch's rows over PCIe -- no host-side gather, no
        staging copy, and `capture()` / `graph_step()` work unchanged (device sampler included).  Meant
        for minibatch steps; a full-batch step would stream the whole matrix through the kernel at PCIe
        speed (`step_from_host` overlaps that with compute instead).  `use_device_matrix()` undoes it."""
        eng = self.eng
        if not (y_pinned.is_pinned() and y_pinned.dtype == torch.float32 and y_pinned.is_contiguous()
                and tuple(y_pinned.shape) == tuple(eng.y_dev.shape)):
            raise ValueError('use_host_matrix: need a pinned, contiguous fp32 [sum N_i, J] matrix')
        if getattr(self, 'graph', None) is not None:
            self.release_graph()
        if not hasattr(self, '_y_resident'):
            self._y_resident = eng.y_dev
        self._y_host = y_pinned                   # keep the allocation alive
        eng.y_dev = y_pinned

    def use_device_matrix(self):
        """back to the device-resident copy of the matrix"""
        if hasattr(self, '_y_resident'):
            if getattr(self, 'graph', None) is not None:
                self.release_graph()
            self.eng.y_dev = self._y_resident
            del self._y_resident, self._y_host

    def step_from_host(self, y_pinned, update=True, no_missing=False):
        """The same full-batch step with the data matrix in PINNED HOST memory (`y_pinned`:
        [sum N_i, J] fp32, NaN = missing): sample i is uploaded on a side stream while the per-cell
        kernel of sample i-1 runs, so the PCIe copy and the compute overlap; the device holds two copies
        of the matrix and alternates between them, so the upload of the NEXT call overlaps the tail of
        this one (callers that read the returned scalars one step late keep the link saturated).  The
        per-sample packed results are added (one tiny torch op) before the backward of the global part.
        After the call `engine.y_dev` is the buffer just uploaded; a captured graph is released.
        `no_missing=True` is the caller's statement that `y_pinned` holds no NaN (checked once on
        the host by whoever owns the matrix); otherwise the imputation path stays compiled in."""
        m, eng, g = self.model, self.eng, self.g
        eng.no_missing = bool(no_missing)
        self.t += 1
        g.step = self.t
        g.adam_step = self.t
        g.noise_mode = _lib.NOISE_PHILOX
        I = eng.I
        main = torch.cuda.current_stream(self.dev)
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(self.dev)
            self._parts = torch.zeros((I, self.packed.numel()), dtype=F64, device=self.dev)
            self._ev = [torch.cuda.Event() for _ in range(I)]
            # two device copies of the matrix: the upload of step t+1 starts while the kernels of step t
            # still read the other one, so the PCIe link never waits for the tail of a step
            self._ybuf = [eng.y_dev, torch.empty_like(eng.y_dev)]
            self._yfree = [None, None]            # event: the last kernel that read the buffer has finished
            self._ynext = 1 if eng.y_dev.data_ptr() == self._ybuf[0].data_ptr() else 0
        cs = self._copy_stream
        bsel = self._ynext
        self._ynext ^= 1
        ybuf = self._ybuf[bsel]
        if self._yfree[bsel] is not None:
            cs.wait_event(self._yfree[bsel])
        else:
            cs.wait_stream(main)                  # first use: whatever was queued before may still read it
        rb = eng.row_base
        with torch.cuda.stream(cs):
            for i in range(I):
                ybuf[rb[i]:rb[i + 1]].copy_(y_pinned[rb[i]:rb[i + 1]], non_blocking=True)
                self._ev[i].record(cs)
        if eng.y_dev.data_ptr() != ybuf.data_ptr():
            if getattr(self, 'graph', None) is not None:
                self.release_graph()              # the captured step holds the other buffer's address
            self._yfree[bsel ^ 1] = torch.cuda.Event()
            self._yfree[bsel ^ 1].record(main)    # everything queued so far may read the buffer being left
            eng.y_dev = ybuf                      # the resident matrix is now the one just uploaded
        stream = C.c_void_p(main.cuda_stream)
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.cytof_global_fwd_f64(
                C.byref(g), _ptr(self.theta), _ptr(None), _ptr(self.stash), _ptr(self.globals),
                _ptr(self.zmask), _ptr(self.scalars_in), stream))
        full = Batch(eng.N_local, None, None, step=self.t)
        cg = m._counts_global(full) or eng.N_local
        for i in range(I):
            main.wait_event(self._ev[i])
            counts = [0] * I
            counts[i] = eng.N_local[i]
            eng.fwdbwd_packed(Batch(counts, None, None, step=self.t), self.globals, self.zmask,
                              self._parts[i], [cg[j] if j == i else 0 for j in range(I)])
        self._yfree[bsel] = torch.cuda.Event()
        self._yfree[bsel].record(main)            # the per-cell kernels of this step are the last readers
        torch.sum(self._parts, 0, out=self.packed)
        if m.dist_group is not None:
            torch.distributed.all_reduce(self.packed, group=m.dist_group)
        with torch.cuda.device(self.dev):
            _lib.check(self.lib.cytof_global_bwd_adam_f64(
                C.byref(g), _ptr(self.theta), _ptr(self.stash), _ptr(self.packed), _ptr(self.scalars_in),
                _ptr(self.adam_m), _ptr(self.adam_v), _ptr(self.grad), _ptr(self.scalars),
                _ptr(None), _ptr(self.flag), C.c_int32(1 if update else 0), stream))
        if self.peers is not None:
            self.peers.seq += 1     # no exchange through the mailboxes in this step: keep the sequence in lock step with t
        self.launches += 2 * I + self._global_launches
        return self.scalars
