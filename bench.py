#!/usr/bin/env python
"""Benchmark of the ELBO forward+backward hot path (metric of BASELINE.json:
"ELBO fwd+bwd cells/sec (J=32,K=30,L=5) at 1/2/4/8 B200; % of HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4]

A "step" is one full `update` of the reference (fit.py:32-48): draw the global
sample, transforms, per-cell ELBO forward+backward over ONE batch (the fused
CUDA kernel), log prior, entropy, backward to every variational parameter, NaN
scrub and the Adam step — here 5 kernel launches (cytofpy_b200/model/fused.py).  The default workload is BASELINE.json configs[3]
("cfg4": full batch, I=8, J=32, K=30, L=5/5) — the configuration the metric
string names.  `--scaling strong` (default) is the split BASELINE.json states: the
1,000,000 cells of the job are sharded over the N GPUs (rank r holds rows
[r*125000/N, (r+1)*125000/N) of each of the 8 samples; the ~4.3k-double packed
gradient block is summed over the ranks inside the finalise kernel).  With N > 1
the same line carries the weak-scaling figure (1,000,000 cells PER GPU) under
"weak"; `--scaling weak` makes that the headline instead.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: I, cells per sample (per GPU), J, K, L0, L1, minibatch (None = full batch), NaN rate, fit kwargs
    'cfg1': dict(I=3, n=2000, J=32, K=5, L0=3, L1=3, mb=None, nan=0.0, kw={}),
    'cfg2': dict(I=3, n=40000, J=32, K=30, L0=5, L1=3, mb=2000, nan=0.0, kw={}),
    'cfg3': dict(I=3, n=40000, J=32, K=30, L0=5, L1=3, mb=2000, nan=0.3,
                 kw=dict(use_stick_break=False, scale=.1, tau=0.001)),
    'cfg4': dict(I=8, n=125000, J=32, K=30, L0=5, L1=5, mb=None, nan=0.0, kw={}),
    'cfg5': dict(I=8, n=1250000, J=40, K=50, L0=8, L1=8, mb=None, nan=0.0, kw={}),
}
_REAL_STDOUT = None


def emit(obj):
    line = (json.dumps(obj) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


METRIC = 'ELBO fwd+bwd cells/sec (J=32,K=30,L=5)'
UNIT = 'cells/s'
L2_FLUSH_BYTES = 512 << 20


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def profiled_traffic(workload):
    """dram__bytes_read + dram__bytes_write of ONE launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json: written next to the digest it was read from, with the sha1 of the kernel source it was
    taken on).  None when there is no capture for this workload or the kernel source has changed since."""
    import hashlib
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            rec = json.load(f)[workload]
        with open(os.path.join(ROOT, rec['source']), 'rb') as f:
            if hashlib.sha1(f.read()).hexdigest() != rec['source_sha1']:
                return None
        return float(rec['dram_bytes_per_launch'])
    except Exception:
        return None


def make_data(w, seed):
    import cytofpy_b200 as cy
    s = cy.util.synth_cytof([w['n']] * w['I'], w['J'], [100.] * 5, w['L0'], w['L1'], seed=seed)
    if w['nan'] > 0:
        beta = cy.model.solve_beta([-5., -3.5, -2.], [.05, .8, .05]).numpy()
        cy.util.inject_missing(s['y'], beta, rate=w['nan'], seed=seed + 1)
    return s['y']


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons during the timed region (NVML)"""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown',
               0x4: 'sw_power_cap', 0x80: 'hw_power_brake', 0x2: 'applications_clocks_setting'}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        self.ready = threading.Event()       # NVML is initialised before the timed region starts

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[self.index]) if vis and vis.split(',')[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            self.ready.set()
            while not self.stop_flag:
                self.sm.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.0002)
        except Exception as e:  # NVML missing: report it, do not fake numbers
            self.reasons.add('nvml_unavailable:%s' % type(e).__name__)
        self.ready.set()

    def summary(self):
        return {'sm_mhz': statistics.median(self.sm) if self.sm else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.sm)}


def time_cpu_port(w, steps, warmup, cells_per_sample, threads=None):
    """The oracle port of the reference `update` (fp64 torch + autograd + Adam) on host cores,
    on a bounded sample of the workload: same I/J/K/L, `cells_per_sample` cells per sample."""
    from oracle import elbo_port as port
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    ws = dict(w, n=cells_per_sample)
    y = [torch.tensor(a, dtype=torch.float64) for a in make_data(ws, seed=0)]
    kw = w['kw']
    cfg = port.make_cfg(y, w['K'], w['L0'], w['L1'], tau=kw.get('tau', .1),
                        use_stick_break=kw.get('use_stick_break', True), scale=kw.get('scale', 1.0))
    st = port.init_state(w['I'], w['J'], w['K'], w['L0'], w['L1'], use_stick_break=cfg['use_stick_break'])
    opt = torch.optim.Adam(port.state_parameters(st), lr=1e-2)
    mb = w['mb'] if w['mb'] and w['mb'] < cells_per_sample else None
    rng = np.random.RandomState(0)
    times, cells = [], 0
    for t in range(warmup + steps):
        t0 = time.perf_counter()
        idx = [np.arange(cells_per_sample) if mb is None else rng.choice(cells_per_sample, mb, replace=False)
               for _ in range(w['I'])]
        port.update_port(opt, st, y, idx, port.draw_noise(cfg, idx), cfg)
        if t >= warmup:
            times.append(time.perf_counter() - t0)
            cells = sum(len(ix) for ix in idx)
    med = statistics.median(times)
    return cells / med, torch.get_num_threads(), cells, med


REFERENCE_DIR = '/root/reference'


def time_cpu_reference(w, steps, warmup, cells_per_sample, threads=None):
    """The UNMODIFIED reference (`cytofpy.model.update`, imported from /root/reference -- present in the build
    container only) on the same bounded sample as time_cpu_port.  Only shim: validate_args off (reference vae.py:29
    builds Normal(., 0), which torch >= 1.8 rejects by default)."""
    sys.path.insert(0, REFERENCE_DIR)
    torch.distributions.Distribution.set_default_validate_args(False)
    import cytofpy
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    np.random.seed(0)
    ws = dict(w, n=cells_per_sample)
    y = [torch.tensor(a, dtype=torch.float64) for a in make_data(ws, seed=0)]
    kw = w['kw']
    pri = cytofpy.model.default_priors(y, K=w['K'], L=[w['L0'], w['L1']], y_bounds=[-5., -3.5, -2.])
    mod = cytofpy.model.Model(y=y, priors=pri, tau=kw.get('tau', .1), verbose=-1,
                              use_stick_break=kw.get('use_stick_break', True), scale=kw.get('scale', 1.0))
    opt = torch.optim.Adam(mod.parameters(), lr=1e-2)
    mb = w['mb'] if w['mb'] and w['mb'] < cells_per_sample else None
    times, cells = [], 0
    for t in range(warmup + steps):
        t0 = time.perf_counter()
        idx = [np.arange(cells_per_sample) if mb is None else np.random.choice(cells_per_sample, mb, replace=False)
               for _ in range(w['I'])]
        cytofpy.model.update(opt, mod, idx)
        if t >= warmup:
            times.append(time.perf_counter() - t0)
            cells = sum(len(ix) for ix in idx)
    med = statistics.median(times)
    return cells / med, torch.get_num_threads(), cells, med


def time_cpu(w, steps, warmup, cells_per_sample, threads=None):
    """(cells/s, threads, cells per step, seconds per step, kind): the reference itself where it is on the machine,
    else its oracle port"""
    if os.path.isdir(os.path.join(REFERENCE_DIR, 'cytofpy')):
        try:
            return time_cpu_reference(w, steps, warmup, cells_per_sample, threads) + ('reference',)
        except Exception as e:      # e.g. an incompatible torch: fall back to the port, say so on stderr
            sys.stderr.write('bench: reference import failed (%s), timing the oracle port\n' % e)
    return time_cpu_port(w, steps, warmup, cells_per_sample, threads) + ('port',)


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    per = min(w['n'], 4000 if w['mb'] is None else w['n'])
    # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers: override it)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    v, cores, cells, med, kind = time_cpu(w, args.steps, args.warmup, per, threads=threads)
    sample = '%s update (fp64 torch CPU, autograd, Adam) on %d cells/step ' \
             '(%d samples x %d%s) at the workload J/K/L' % (
                 'the reference (cytofpy.model.update from /root/reference)' if kind == 'reference'
                 else 'oracle port of the reference', cells, w['I'], per,
                 '' if w['mb'] is None else ', minibatch %d' % w['mb'])
    emit({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': med * 1e3, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.workload, 'I': w['I'], 'J': w['J'], 'K': w['K'], 'L': [w['L0'], w['L1']]},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    })


def build_job(args, w, rank, world, dev, group, scaling):
    """This rank's shard of the workload and the fused step that trains on it.  strong: the workload's cells (and its
    minibatch) are divided over the ranks; weak: every rank holds the whole workload size."""
    import cytofpy_b200 as cy
    from cytofpy_b200.model.fused import FusedStep
    I = w['I']
    n = w['n'] // world if scaling == 'strong' else w['n']
    mb = None if w['mb'] is None else (max(1, w['mb'] // world) if scaling == 'strong' else w['mb'])
    y_np = [a[:n] for a in make_data(w, seed=1000 + rank)]   # this rank's rows
    y = [torch.from_numpy(np.ascontiguousarray(a)) for a in y_np]
    N_global = [n * world] * I
    pri_y = [t[:2000].double() for t in y]
    priors = cy.model.default_priors(pri_y, K=w['K'], L=[w['L0'], w['L1']], y_bounds=[-5., -3.5, -2.])
    priors['N'] = N_global
    torch.manual_seed(1)
    model = cy.model.Model(y=y, priors=priors, verbose=-1, device=dev, noise='philox', seed=1,
                           N_global=N_global, row_global=[rank * n] * I, dist_group=group, **w['kw'])
    fs = FusedStep(model, lr=1e-2, seed=1, allreduce=args.allreduce)   # the path cytofpy_b200.model.fit drives

    def draw_idx(t):
        return [range(n)] * I if mb is None else cy.model.device_minibatch(model, mb, t, 1)

    def one_step(t):
        # after the first (host-driven) warm-up step the whole step is ONE CUDA-graph replay
        if t == 0:
            return fs.step(draw_idx(t))
        if getattr(fs, 'graph', None) is None:
            fs.capture(mb)
        return fs.graph_step()

    return dict(model=model, fs=fs, y=y, n=n, mb=mb, one_step=one_step, draw_idx=draw_idx,
                cells_step=I * (n if mb is None else mb), scaling=scaling)


def timed_steps(job, args, dev, flush, barrier, local_rank):
    """W warm-up steps, then K steps timed one by one with CUDA events on the launching stream (L2 flushed before
    each); returns (mean ms per step, launches inside the timed region, clock summary)."""
    fs, one_step = job['fs'], job['one_step']
    for t in range(args.warmup):
        one_step(t)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.ready.wait(10.0)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = fs.launches
    for t in range(args.steps):
        flush.fill_(t & 0xff)                                   # evict y from the 126 MB L2
        ev[t][0].record()
        one_step(args.warmup + t)
        ev[t][1].record()
    barrier()
    step_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    launches = fs.launches - launches0      # graph_step counts the sampler kernels itself
    # NVML answers in milliseconds; K short steps can end before its first sample.  Keep running the SAME step
    # (untimed, the same number on every rank) until the region under load has lasted ~50 ms, so that `clocks`
    # always holds samples taken while these kernels run
    est = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(est, op=torch.distributed.ReduceOp.MAX)
    est = max(float(est.item()), 1e-3)
    n_extra = min(20000, max(20, int(50.0 / est) + 1 if args.steps * est < 50.0 else 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(n_extra):
        one_step(args.warmup + args.steps + t)
    e1.record()
    barrier()
    loop_ms = e0.elapsed_time(e1) / n_extra
    sampler.stop_flag = True
    sampler.join(2.0)
    clocks = sampler.summary()
    clocks['note'] = ('NVML samples over the K timed steps and %d further back-to-back replays of the same step '
                      '(K steps last %.1f ms; one NVML round trip takes milliseconds)' % (n_extra, args.steps * est))
    # loop_ms is NOT the contract's number: the same step replayed back to back with no L2 flush, as `fit` runs it
    return step_ms, launches, clocks, (loop_ms, n_extra)


def run_ours(args, w, rank, world, local_rank):
    import cytofpy_b200 as cy
    from cytofpy_b200.model.engine import Batch
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
        group = dist.group.WORLD
    scaling = args.scaling
    job = build_job(args, w, rank, world, dev, group, scaling)
    model, fs, y, n, mb = job['model'], job['fs'], job['y'], job['n'], job['mb']
    I, J = w['I'], w['J']
    one_step, draw_idx = job['one_step'], job['draw_idx']
    cells_step = I * (n if mb is None else mb)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    eng = model.engine

    # ---- device-resident throughput: K timed steps, L2 flushed between them
    def barrier():
        if world > 1:
            torch.distributed.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    step_ms, launches, clocks, (loop_ms, loop_n) = timed_steps(job, args, dev, flush, barrier, local_rank)
    exchange = None
    if fs.peers is not None:      # where the last step's fused exchange spent its time on this rank (CTA 0, globaltimer)
        tl = fs.peers.timeline()
        exchange = {'own_part_stored_us': (tl[1] - tl[0]) / 1e3, 'wait_for_peers_us': (tl[2] - tl[1]) / 1e3,
                    'sum_over_ranks_us': (tl[3] - tl[2]) / 1e3, 'rank': rank}

    fs.release_graph()
    # ---- dominant kernel alone (CUDA events on the launching stream, L2 flushed)
    model._step += 1
    with torch.no_grad():
        reals = model._sample_globals()
        kin = model._kernel_inputs(model.transform(reals))
        g32 = torch.cat([t.reshape(-1) for t in kin[:6]] + [kin[7].reshape(-1), eng.b_dev.reshape(-1),
                                                             kin[8].reshape(-1), kin[9].reshape(-1)]).float()
        bits = (kin[6] > 0.5).to(torch.int64)
        zmask = (bits << torch.arange(w['K'], device=dev, dtype=torch.int64)[None, :]).sum(1).contiguous()
    counts, idx_dev = eng.make_idx(draw_idx(0))
    kb = Batch(counts, idx_dev, None, step=1)
    kev = []
    for t in range(3 + 10):
        flush.fill_(t & 0xff)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.fwdbwd(kb, g32, zmask, None if group is None else [c * world for c in counts])
        b.record()
        kev.append((a, b))
    torch.cuda.synchronize(dev)
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in kev[3:])
    bytes_cell = 4 * J + (4 if idx_dev is not None else 0)
    achieved = cells_step * bytes_cell / (k_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    # ---- end to end through the public API with HOST buffers: every step copies the step's
    # input rows from pinned host memory to the device and reads the ELBO back
    y_host = torch.cat(y, 0)
    with cy.util.on_gpu_node(local_rank) as numa_cpus:      # the pinned copy lands on the GPU's NUMA node
        y_pin = y_host.pin_memory()
    del y_host
    pin_clean = not bool(torch.isnan(y_pin).any())      # host-side mask check, once (fit.py:66)
    h2d = y_pin.numel() * 4 if mb is None else cells_step * 4 * J      # rows that cross PCIe per step
    e_ev = []
    n_e2e = 2 + max(3, args.steps // 2)
    if mb is None:
        # full batch: sample-by-sample upload overlapped with the per-cell kernels; the 40-byte result of
        # every step is copied to pinned host memory and READ ON THE HOST ONE STEP LATE (what a training
        # loop that logs the ELBO does), so the upload of step t+1 is already queued while step t computes
        res_pin = torch.empty((2, 5), dtype=torch.float64).pin_memory()
        res_ev = [torch.cuda.Event(), torch.cuda.Event()]
        seen = []
        for t in range(n_e2e):
            if t == 2:      # timed region starts with an empty pipeline
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            sc = fs.step_from_host(y_pin, no_missing=pin_clean)
            res_pin[t & 1].copy_(sc, non_blocking=True)
            res_ev[t & 1].record()
            if t >= 1:
                res_ev[(t - 1) & 1].synchronize()
                seen.append(res_pin[(t - 1) & 1].tolist())
        res_ev[(n_e2e - 1) & 1].synchronize()
        seen.append(res_pin[(n_e2e - 1) & 1].tolist())
        e1.record()
        e1.synchronize()
        assert len(seen) == n_e2e and all(v[0] == v[0] for v in seen), 'e2e: a step returned NaN'
        e_ev.append((e0, e1))
        e2e_div = n_e2e - 2
    if mb is not None:
        # minibatch: the matrix stays in pinned host memory; the step (one CUDA-graph replay, minibatch drawn
        # by the device sampler) gathers its rows over PCIe inside the per-cell kernel; the result is copied
        # back every step and read on the host one step late
        fs.use_host_matrix(y_pin)
        fs.capture(mb)
        res_pin = torch.empty((2, 5), dtype=torch.float64).pin_memory()
        res_ev = [torch.cuda.Event(), torch.cuda.Event()]
        seen = []
        for t in range(n_e2e):
            if t == 2:
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            sc = fs.graph_step()
            res_pin[t & 1].copy_(sc, non_blocking=True)
            res_ev[t & 1].record()
            if t >= 1:
                res_ev[(t - 1) & 1].synchronize()
                seen.append(res_pin[(t - 1) & 1].tolist())
        res_ev[(n_e2e - 1) & 1].synchronize()
        seen.append(res_pin[(n_e2e - 1) & 1].tolist())
        e1.record()
        e1.synchronize()
        assert len(seen) == n_e2e and all(v[0] == v[0] for v in seen), 'e2e: a step returned NaN'
        e_ev.append((e0, e1))
        e2e_div = n_e2e - 2
        fs.use_device_matrix()
    barrier()
    e2e_ms = e_ev[0][0].elapsed_time(e_ev[0][1]) / e2e_div

    peers_desc = 'single GPU: none' if world == 1 else (
        'fused into the finalise kernel, NVLink peer stores' if fs.peers is not None else 'NCCL')
    # ---- N > 1: the OTHER scaling mode, device-resident steps only, in the same run
    other, other_ms, other_cells = None, 0.0, 0
    if world > 1 and not args.one_mode:
        other = 'weak' if scaling == 'strong' else 'strong'
        fs.release_graph()
        if fs.peers is not None:
            fs.peers.close()
        del model, fs, job, y_pin
        torch.cuda.empty_cache()
        job2 = build_job(args, w, rank, world, dev, group, other)
        other_ms, _, _, _ = timed_steps(job2, args, dev, flush, barrier, local_rank)
        other_cells = job2['cells_step']
        peers2 = job2['fs'].peers is not None
        job2['fs'].release_graph()

    times = torch.tensor([step_ms, k_ms, e2e_ms, other_ms, loop_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(times, op=torch.distributed.ReduceOp.MAX)
    step_ms, k_ms_max, e2e_ms, other_ms, loop_ms = times.tolist()
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            per = min(n, 4000 if mb is None else n)
            v, cores, cells, med, kind = time_cpu(w, 3, 1, per)
            cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind,
                   'sample': '%s update (fp64 torch CPU) on %d cells/step at the workload J/K/L, median of 3 steps' % (
                       'reference' if kind == 'reference' else 'oracle port of the reference', cells)}
        out = {
            'metric': METRIC, 'value': cells_step * world / (step_ms * 1e-3), 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': step_ms,
            'higher_is_better': True, 'scaling': scaling, 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': args.workload, 'I': I, 'cells_total': I * n * world, 'cells_per_gpu': I * n,
                       'cells_per_step_per_gpu': cells_step,
                       'J': J, 'K': w['K'], 'L': [w['L0'], w['L1']], 'minibatch': mb, 'nan_rate': w['nan'],
                       'parallelism': 'cells sharded x%d, all-reduce of gradient block (%s)' % (world, peers_desc),
                       'l2': 'flushed between timed iterations (512 MiB fill)', 'noise': 'in-kernel philox',
                       'step': 'one CUDA-graph replay (device sampler, global fwd, per-cell kernel, finalise, global bwd + Adam)'},
            'e2e': {'value': cells_step * world / (e2e_ms * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 40, 'ms_per_step': e2e_ms,
                    'host_cpus_bound_to_gpu_node': numa_cpus},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': 'elbo_fwdbwd_mma_kernel (+finalize)', 'achieved': achieved,
                         'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         # dram__bytes_read+write of one launch of that kernel on this workload's full size, from the
                         # committed ncu capture (profiles/traffic.json); None if the kernel changed since
                         'traffic': profiled_traffic(args.workload) if (world == 1 or scaling == 'weak') else None,
                         'peak_source': peak_src, 'bytes_per_cell': bytes_cell, 'kernel_ms': k_ms,
                         'kernel_cells_per_s': cells_step / (k_ms * 1e-3)},
            'cpu_baseline': cpu,
        }
        out['back_to_back'] = {'ms_per_step': loop_ms, 'value': cells_step * world / (loop_ms * 1e-3), 'unit': UNIT,
                               'steps': loop_n, 'note': 'extra, not the contract number: the same graph replayed back to '
                               'back WITHOUT the L2 flush (the training loop as fit runs it), one event pair around all '
                               'replays, max over ranks'}
        if exchange is not None:
            out['exchange_last_step'] = exchange
        if other is not None:
            out[other] = {'scaling': other, 'value': other_cells * world / (other_ms * 1e-3), 'unit': UNIT,
                          'ms_per_step': other_ms, 'cells_total': other_cells * world,
                          'cells_per_step_per_gpu': other_cells,
                          'note': 'same run, device-resident steps, L2 flushed, CUDA events, max over ranks'}
        emit(out)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner)
    # are sent to stderr for the whole run; the JSON goes to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg4', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'],
                    help='N>1: strong = the workload is divided over the GPUs (BASELINE.json: "1M cells ... sharded across '
                         '8 B200"), weak = every GPU holds the whole workload size; the other mode is reported under its own key')
    ap.add_argument('--one-mode', action='store_true', help='N>1: skip the measurement of the other scaling mode')
    ap.add_argument('--allreduce', default='auto', choices=['auto', 'peer', 'nccl'],
                    help='N>1: one-shot all-reduce over peer memory fused into the finalise kernel (auto: NCCL if the '
                         'mailboxes cannot be mapped), or NCCL')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else max(args.warmup, 1)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    w = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, w, rank, world)
    else:
        run_ours(args, w, rank, world, local_rank)


if __name__ == '__main__':
    main()
