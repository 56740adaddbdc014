"""Builds tuning variants of libcytof_b200.so for A/B runs with tools/kbench.py (CYTOF_B200_LIB=...):

    python tools/kvariants.py name:DEF1=V,DEF2=V ...      ->  cytofpy_b200/_C/libv_<name>.so

Only the translation unit that sees the tuning macros is recompiled (variants_mma.cu; variants_mma64.cu for CYTOF_X_*
macros; elbo_kernel.cu as well for CYTOF_MMA_SPLIT); the other objects are those of the default build."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cytofpy_b200 import build as b  # noqa: E402

TUNED = ['variants_mma.cu', 'elbo_kernel.cu']


def one(spec):
    name, _, defs = spec.partition(':')
    defines = [d for d in defs.split(',') if d]
    obj_dir = os.path.join(b.OUT_DIR, 'obj_libv_' + name)
    os.makedirs(obj_dir, exist_ok=True)
    base_dir = os.path.join(b.OUT_DIR, 'obj_libcytof_b200')
    objs = []
    for src in b.SOURCES:
        stem = os.path.splitext(src)[0]
        x64 = any(d.startswith('CYTOF_X_') for d in defines)
        if (src == 'variants_mma.cu' and not x64) or (src == 'variants_mma64.cu' and x64) or \
                (src == 'elbo_kernel.cu' and any('SPLIT' in d for d in defines)):
            obj = os.path.join(obj_dir, stem + '.o')
            r = subprocess.run([b._nvcc()] + b.NVCC_FLAGS + ['-D' + d for d in defines] + ['-c', '-o', obj, os.path.join(b.CSRC, src)],
                               capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(r.stdout + r.stderr)
        else:
            obj = os.path.join(base_dir, stem + '.o')
        objs.append(obj)
    out = os.path.join(b.OUT_DIR, 'libv_%s.so' % name)
    r = subprocess.run([b._nvcc(), '-shared', '-o', out] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)
    return out


if __name__ == '__main__':
    b.build()        # default objects first
    with ThreadPoolExecutor(max_workers=7) as pool:
        for o in pool.map(one, sys.argv[1:]):
            print(o)
