"""GPU: the CTA-wide tcgen05 / TMEM kernel (csrc/elbo_kernel_umma.cuh; all three J x K contractions as M = 128
tcgen05.mma, responsibilities parked in tensor memory, y rows by cp.async.bulk) is an opt-in variant of the
default per-warp kernel (CYTOF_UMMA=1: measured slower, profiles/README.md).  The selection is read once per
process, so the kernel-level parity tests are re-run in a child process with the variable set: golden fixtures
of the live reference, random edge shapes, full-size properties and the full-size fp64 oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_umma_variant_passes_the_kernel_parity_suite():
    env = dict(os.environ, CYTOF_UMMA='1')
    out = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_gpu_parity.py'), '-x', '-q',
                          '-k', 'closed_form or random_shapes or full_size or empty_sample or philox or reproducib'],
                         capture_output=True, text=True, cwd=ROOT, env=env, timeout=1500)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert ' passed' in out.stdout
