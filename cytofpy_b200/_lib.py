"""ctypes binding of libcytof_b200.so (C ABI: include/cytof_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or a call
fails, a RuntimeError is raised — the product never computes the hot path on
the CPU or through plain PyTorch ops."""
import ctypes as C
import os

from . import build as _build

MAX_SAMPLES, MAX_J, MAX_K, MAX_L = 32, 64, 64, 8
NOISE_EXTERNAL, NOISE_PHILOX = 0, 1
FLAG_NO_MISSING = 1
FLAG_SAMPLE_IN_KERNEL = 2


class Desc(C.Structure):
    """mirror of `struct cytof_desc`"""
    _fields_ = [
        ('I', C.c_int32), ('J', C.c_int32), ('K', C.c_int32), ('L0', C.c_int32), ('L1', C.c_int32),
        ('model_noisy', C.c_int32), ('noise_mode', C.c_int32), ('flags', C.c_int32),
        ('noisy_sd', C.c_float), ('scale', C.c_float),
        ('seed', C.c_uint64), ('step', C.c_uint64),
        ('cell_begin', C.c_int64 * (MAX_SAMPLES + 1)),
        ('row_base', C.c_int64 * MAX_SAMPLES),
        ('row_global', C.c_int64 * MAX_SAMPLES),
        ('fac', C.c_double * MAX_SAMPLES),
        ('step_dev', C.c_void_p),
        ('sampler_seed', C.c_uint64), ('sampler_step', C.c_uint64),
        ('sampler_N', C.c_int64 * MAX_SAMPLES),
    ]


class GlobalDesc(C.Structure):
    """mirror of `struct cytof_global_desc`"""
    _fields_ = [
        ('I', C.c_int32), ('J', C.c_int32), ('K', C.c_int32), ('L0', C.c_int32), ('L1', C.c_int32),
        ('model_noisy', C.c_int32), ('use_stick_break', C.c_int32), ('noise_mode', C.c_int32),
        ('sig2_family', C.c_int32), ('reserved0', C.c_int32),
        ('tau', C.c_double),
        ('delta0', C.c_double * 2), ('delta1', C.c_double * 2), ('sig2', C.c_double * 2),
        ('alpha', C.c_double * 2), ('eps', C.c_double * 2),
        ('eta0_conc', C.c_double * MAX_L), ('eta1_conc', C.c_double * MAX_L), ('W_conc', C.c_double * MAX_K),
        ('seed', C.c_uint64), ('step', C.c_uint64),
        ('grad_scale', C.c_double), ('lr', C.c_double), ('beta1', C.c_double), ('beta2', C.c_double),
        ('adam_eps', C.c_double), ('adam_step', C.c_int64),
        ('step_dev', C.c_void_p),
    ]


MAX_PEERS = 16


class PeerDesc(C.Structure):
    """mirror of `struct cytof_peer_desc`"""
    _fields_ = [('world', C.c_int32), ('rank', C.c_int32), ('seq', C.c_uint64),
                ('mailbox', C.c_void_p * MAX_PEERS), ('seq_dev', C.c_void_p)]


_P = C.c_void_p
_PROTOS = {
    'cytof_abi_version': (C.c_int, []),
    'cytof_strerror': (C.c_char_p, [C.c_int]),
    'cytof_last_cuda_error': (C.c_char_p, []),
    'cytof_globals_layout': (C.c_int, [C.POINTER(Desc), C.POINTER(C.c_int64)]),
    'cytof_grad_block_layout': (C.c_int, [C.POINTER(Desc), C.POINTER(C.c_int64)]),
    'cytof_workspace_bytes': (C.c_size_t, [C.POINTER(Desc)]),
    'cytof_elbo_fwdbwd_f32': (C.c_int, [C.POINTER(Desc), _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    'cytof_elbo_fwdbwd_packed': (C.c_int, [C.POINTER(Desc), _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    'cytof_peer_mailbox_bytes': (C.c_size_t, [C.POINTER(Desc), C.c_int32]),
    'cytof_peer_alloc': (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    'cytof_peer_free': (C.c_int, [_P]),
    'cytof_ipc_get_handle': (C.c_int, [_P, C.c_char_p]),
    'cytof_ipc_open_handle': (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    'cytof_ipc_close': (C.c_int, [_P]),
    'cytof_peer_debug': (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    'cytof_elbo_fwdbwd_allreduce': (C.c_int, [C.POINTER(Desc), _P, _P, _P, _P, _P, _P, _P, C.c_size_t,
                                              C.POINTER(PeerDesc), _P]),
    'cytof_global_sizes': (C.c_int, [C.POINTER(GlobalDesc), C.POINTER(C.c_int64)]),
    'cytof_global_fwd_f64': (C.c_int, [C.POINTER(GlobalDesc), _P, _P, _P, _P, _P, _P, _P]),
    'cytof_global_bwd_adam_f64': (C.c_int, [C.POINTER(GlobalDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                            C.c_int32, _P]),
    'cytof_impute_emit_f32': (C.c_int, [C.POINTER(Desc), _P, _P, _P, _P, _P, _P]),
    'cytof_label_sample_f32': (C.c_int, [C.POINTER(Desc), _P, _P, _P, _P, _P, _P]),
    'cytof_noise_emit_f32': (C.c_int, [C.POINTER(Desc), _P, _P, _P]),
    'cytof_adam_step_f64': (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int64, C.c_double, C.c_double,
                                      C.c_double, C.c_double, C.c_double, C.c_int64, _P, _P, _P]),
    'cytof_sample_minibatch': (C.c_int, [C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_int32, _P, _P]),
    'cytof_sample_minibatch_multi': (C.c_int, [C.c_int32, _P, _P, C.c_uint64, C.c_uint64, _P, _P, _P]),
    'cytof_sample_minibatch_dev': (C.c_int, [C.c_int64, C.c_int64, C.c_uint64, _P, C.c_int32, _P, _P]),
}
EXPORTS = tuple(_PROTOS)

_lib = None


def lib_path():
    # CYTOF_B200_LIB selects a tuning build of the same sources (tools/kbench.py); never a fallback
    return os.environ.get('CYTOF_B200_LIB') or _build.LIB_PATH


def load():
    """Load (once) and return the ctypes handle; raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError(
            'cytofpy_b200: %s not found. Build it with `python -m cytofpy_b200.build` '
            '(nvcc, sm_100a). There is no CPU / PyTorch fallback for the ELBO hot path.' % path)
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.cytof_abi_version() != 2:
        raise RuntimeError('cytofpy_b200: ABI version mismatch')
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        lib = load()
        msg = lib.cytof_strerror(rc).decode()
        if rc == -4:
            msg += ' [' + lib.cytof_last_cuda_error().decode() + ']'
        raise RuntimeError('cytofpy_b200 C-ABI call failed (%d): %s' % (rc, msg))


def layout(desc, which):
    lib = load()
    n = 11 if which == 'globals' else 12
    off = (C.c_int64 * n)()
    fn = lib.cytof_globals_layout if which == 'globals' else lib.cytof_grad_block_layout
    check(fn(C.byref(desc), off))
    return [int(v) for v in off]
