"""Host-side placement for the upload path (`FusedStep.step_from_host`, `use_host_matrix`): pinned host memory is
placed on the NUMA node of the thread that allocates it, and a GPU reads it at full PCIe speed only from its own node
(4-GPU B200 boxes of this pool: 54 GB/s from the local node, ~35 GB/s across the socket link).  The reference has
no counterpart (it is a CPU program; `sims/cb/cb.py:33` pins it to one thread)."""
import contextlib
import os


def gpu_local_cpus(device_index=0):
    """CPUs NVML reports as local to GPU `device_index` (an index into CUDA_VISIBLE_DEVICES when that is set),
    intersected with the CPUs this thread may run on; empty set if unknown."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        ids = vis.split(',') if vis else []
        phys = int(ids[device_index]) if device_index < len(ids) and ids[device_index].isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        return local & os.sched_getaffinity(0)
    except Exception:
        return set()


@contextlib.contextmanager
def on_gpu_node(device_index=0):
    """`with on_gpu_node(i): buf = t.pin_memory()` -- the calling THREAD runs on the CPUs local to GPU i while the
    block executes (so what it allocates and pins lands on that node) and gets its previous affinity back afterwards;
    other threads (torch's intra-op pool) are not touched.  Yields the number of CPUs bound to, 0 if nothing was
    changed (NVML missing, no affinity information): placement is an optimisation, never an error."""
    before = os.sched_getaffinity(0)
    target = gpu_local_cpus(device_index)
    changed = bool(target) and target != before
    if changed:
        try:
            os.sched_setaffinity(0, target)
        except OSError:
            changed = False
    try:
        yield len(target) if changed else 0
    finally:
        if changed:
            os.sched_setaffinity(0, before)
