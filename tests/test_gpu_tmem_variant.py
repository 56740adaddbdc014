"""GPU: the tcgen05 / TMEM variant of the dZ accumulation (csrc/elbo_kernel_mma.cuh, CYTOF_TMEM_DZ=1:
one tcgen05.mma.kind::tf32 per 8-cell tile, M = N = 64, K = 8, per-warp accumulators in tensor memory,
read back with tcgen05.ld) must pass the same parity tests as the default build.  The variant library
is built by __graft_entry__.build(); the tests run in a subprocess because the library is chosen at
load time (CYTOF_B200_LIB)."""
import os
import subprocess
import sys

import pytest

from cytofpy_b200 import build as b

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parity_suite_with_tmem_dz_variant():
    if not os.path.exists(b.TMEM_LIB_PATH):
        pytest.skip('variant library not built (python __graft_entry__.py)')
    env = dict(os.environ, CYTOF_B200_LIB=b.TMEM_LIB_PATH)
    r = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_parity.py', 'tests/test_gpu_fused_step.py',
                        '-q', '-x', '-k', 'closed_form or fixture or random_shapes or cfg4_shape or reproducible'],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert ' passed' in r.stdout


def test_global_part_fallback_path_and_dependent_launch_option():
    """csrc/global_part.cu runs the scans in the last CTA of the element kernels when their operands fit
    in shared memory and as separate single-CTA kernels otherwise; CYTOF_GLOBAL_UNFUSED=1 forces the
    second path.  CYTOF_PDL=1 launches every kernel of the step with programmatic stream serialisation
    (csrc/common.cuh).  Both must pass the fused-step and step-loop suites."""
    for extra in ({'CYTOF_GLOBAL_UNFUSED': '1'}, {'CYTOF_PDL': '1'}):
        env = dict(os.environ, **extra)
        r = subprocess.run([sys.executable, '-m', 'pytest', 'tests/test_gpu_fused_step.py', 'tests/test_gpu_steploop.py',
                            '-q', '-x'], cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, str(extra) + r.stdout[-3000:] + r.stderr[-2000:]
        assert ' passed' in r.stdout
