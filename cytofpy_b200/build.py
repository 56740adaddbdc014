"""Builds libcytof_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/cytof_b200.h) in-tree with nvcc.  No torch / pybind dependency: the
library is loaded with ctypes (cytofpy_b200/_lib.py)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.path.join(HERE, '_C')
LIB_PATH = os.path.join(OUT_DIR, 'libcytof_b200.so')
SOURCES = ['variants_mma64.cu', 'variants_mma.cu', 'variants_umma.cu', 'elbo_kernel.cu', 'global_part.cu', 'labels.cu', 'abi.cu']
NVCC_FLAGS = ['-std=c++17', '-O3', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC']


def _nvcc():
    for c in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def _stale(path=None):
    path = path or LIB_PATH
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, '..', 'include', 'cytof_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile the CUDA library if missing or older than its sources.  `defines` / `out`
    build a tuning variant under another name (tools/kbench.py)."""
    if out is None and not force and not _stale():
        return LIB_PATH
    os.makedirs(OUT_DIR, exist_ok=True)
    target = out or LIB_PATH
    # one object per translation unit, compiled in parallel (the template-heavy ones take ~1 min each)
    obj_dir = os.path.join(OUT_DIR, 'obj_' + os.path.splitext(os.path.basename(target))[0])
    os.makedirs(obj_dir, exist_ok=True)
    base = [_nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-D' + d for d in defines]

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + '.o')
        r = subprocess.run(base + ['-c', '-o', obj, os.path.join(CSRC, src)], capture_output=True, text=True)
        return src, obj, r

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for src, obj, r in results:
        if r.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n' % src + r.stdout + r.stderr)
        if verbose:
            sys.stderr.write(r.stderr)
    r = subprocess.run([_nvcc(), '-shared', '-o', target] + [obj for _, obj, _ in results], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    return target


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
