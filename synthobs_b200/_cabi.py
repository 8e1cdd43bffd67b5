"""ctypes binding of libsynthobs_b200.so (the C-ABI declared in include/synthobs_b200.h).

This is the only place Python touches the CUDA kernels.  There is no CPU path:
if the library is missing the import fails with the build command, and every
compute entry point raises when no sm_100 device is present.  PyTorch is used by
the callers only for device memory, streams and torch.distributed.
"""

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libsynthobs_b200.so')

if not os.path.exists(LIB_PATH):
    raise ImportError(
        'synthobs_b200: %s not found. Build it first: `python -m synthobs_b200.build` '
        '(or __graft_entry__.build()). There is no CPU fallback.' % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

P = c_void_p
I64 = c_int64
I32 = c_int32
D = c_double

# name -> (restype, argtypes); must list every symbol declared in include/synthobs_b200.h
SIGNATURES = {
    'so_version': (c_char_p, []),
    'so_device_ok': (c_int, []),
    'so_sed_broadband_workspace_bytes': (I64, [I64, I64, I32]),
    'so_sed_broadband': (c_int, [P, P, P, P, P, I64, P, I64, POINTER(c_double), I32, POINTER(c_double), I32, D, P, P, I32,
                                 POINTER(c_double), POINTER(c_double), D, P, P, P, P, I64, P]),
    'so_sed_hist_workspace_bytes': (I64, [I64, I32]),
    'so_sed_hist': (c_int, [P, P, I64, P, I64, I32, P, P, I64, P, I64, P]),
    'so_split_tf32': (c_int, [P, P, P, I64, P]),
    'so_gemm_3xtf32': (c_int, [P, P, I64, P, P, I64, P, I64, I64, I64, I64, P]),
    'so_gemm_f32': (c_int, [P, I64, P, I64, P, I64, I64, I64, I64, P]),
    'so_sed_dust_spectra_workspace_bytes': (I64, [I64, I64, I32]),
    'so_sed_dust_spectra': (c_int, [P, P, P, P, P, I64, P, I64, P, P, I32, I32, P, P, D, P, P, P, P, I64, P]),
    'so_img_ngp_index': (c_int, [P, P, I64, P, I32, D, P, P, P]),
    'so_knn_workspace_bytes': (I64, [I64, I32]),
    'so_knn_kth_distance': (c_int, [P, P, P, I64, I32, D, P, P, I64, P]),
    'so_knn_kth_distance_part': (c_int, [P, P, P, I64, I32, D, I32, I32, P, P, P, P, I64, P]),
    'so_knn_batch_workspace_bytes': (I64, [I64, I64, I32]),
    'so_knn_kth_distance_batch': (c_int, [P, P, P, P, I64, I64, I32, D, P, P, I64, P]),
    'so_knn_kth_distance_ex': (c_int, [P, P, P, P, I64, I64, I32, D, I32, I32, D, P, P, P, P, I64, P]),
    'so_knn_kth_distance_q': (c_int, [P, P, P, I64, I32, D, P, D, P, P, I64, P]),
    'so_share_plan_workspace_bytes': (I64, [I32]),
    'so_share_hist': (c_int, [P, P, I64, I64, D, I32, P, P]),
    'so_share_plan': (c_int, [P, P, I64, D, I32, P, I32, D, D, D, P, P, P, I64, P]),
    'so_share_compact_workspace_bytes': (I64, [I64]),
    'so_share_compact': (c_int, [P, P, P, I64, I32, I64, P, P, P, P, P, P, I64, P]),
    'so_img_plan_workspace_bytes': (I64, [I64]),
    'so_img_plan': (c_int, [P, P, P, P, P, I64, P, I32, D, P, I64, POINTER(c_int64), POINTER(c_double), P]),
    'so_img_deposit_workspace_bytes': (I64, [I64, I64, I32, I32]),
    'so_img_deposit': (c_int, [P, P, I64, I32, I32, I64, P, P, P, P, P, P, I64, P]),
    'so_img_deposit_batch_workspace_bytes': (I64, [I64, I64, I32, I32, I32]),
    'so_img_deposit_batch': (c_int, [P, P, P, I64, I32, I32, I32, I64, P, P, P, P, P, P, I64, P]),
    'so_img_rebin': (c_int, [P, P, I32, I32, I32, P]),
    'so_img_rebin_ld': (c_int, [P, I64, I64, P, I32, I32, I32, P]),
    'so_img_deposit_batch_ld': (c_int, [P, P, P, I64, I32, I32, I32, I64, P, I64, I64, P, P, P, P, P, I64, P]),
    'so_img_deposit_push': (c_int, [P, P, P, I64, I32, I32, I32, I64, P, I64, I64, P, P, P, P, P, I64, P, P]),
    'so_complex_mul_bcast': (c_int, [P, P, P, I64, I32, I32, P]),
    'so_median_workspace_bytes': (I64, [I64, I32]),
    'so_median_f64': (c_int, [P, I64, I32, I64, P, P, I64, P]),
    'so_median_f64_mid': (c_int, [P, I64, I32, I64, P, P, P, I64, P]),
    'so_recentre_chain': (c_int, [P, P, I64, P, I32, POINTER(c_double), POINTER(c_double), I32, POINTER(c_int32),
                                  POINTER(c_void_p), POINTER(c_void_p), P]),
    'so_median_segmented': (c_int, [P, P, I64, P, P]),
    'so_recentre_segmented': (c_int, [P, P, I64, P, D, P]),
    'so_img_sersic_workspace_bytes': (I64, [I32]),
    'so_img_sersic': (c_int, [P, I32, D, D, D, D, D, D, D, P, I32, P, P, I64, P]),
    'so_peer_block_edge': (I32, []),
    'so_peer_alloc': (c_int, [I64, POINTER(c_void_p)]),
    'so_peer_free': (c_int, [P]),
    'so_ipc_get_handle': (c_int, [P, P]),
    'so_ipc_open': (c_int, [P, POINTER(c_void_p)]),
    'so_ipc_close': (c_int, [P]),
    'so_peer_block_mask': (c_int, [P, I32, I32, I64, I64, P, P]),
    'so_peer_sum': (c_int, [POINTER(c_void_p), P, I32, I32, I32, I32, P, P]),
    'so_peak_kernel': (c_int, [I32, I64, I32, P, POINTER(c_double), P]),
    'so_los_workspace_bytes': (I64, [I64, I64]),
    'so_los_column': (c_int, [P, P, P, POINTER(c_int64), P, P, P, P, P, P, POINTER(c_int64), I64, P, I32, D, P, P, I64, P]),
    'so_img_deposit_count': (c_int, [P]),
    'so_prof_enable': (c_int, [I32]),
    'so_prof_report': (I64, [c_char_p, I64]),
}

SO_MAX_PEERS = 16


class PushDesc(ctypes.Structure):
    """so_push_t of include/synthobs_b200.h"""
    _fields_ = [('n_ranks', c_int32), ('rank', c_int32), ('epoch', ctypes.c_uint64),
                ('partial', c_void_p * SO_MAX_PEERS), ('result', c_void_p * SO_MAX_PEERS),
                ('masks', c_void_p * SO_MAX_PEERS), ('flags', c_void_p * SO_MAX_PEERS)]


for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args

_ERR = {-1: 'SO_ERR_BAD_ARG', -2: 'SO_ERR_TOO_LARGE', -3: 'SO_ERR_WORKSPACE', -4: 'SO_ERR_UNSUPPORTED'}


class SynthObsCudaError(RuntimeError):
    pass


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise SynthObsCudaError('%s failed: %s' % (what, _ERR.get(rc, rc)))
    raise SynthObsCudaError('%s failed: CUDA error %d' % (what, rc))


_device_checked = False


def require_device():
    """Fail loudly when there is no B200-class device: no CPU fallback exists."""
    global _device_checked
    if _device_checked:
        return
    import torch
    if not torch.cuda.is_available() or not lib.so_device_ok():
        raise SynthObsCudaError('synthobs_b200 needs a CUDA device of compute capability 10.x (B200); '
                                'none is visible and there is no CPU fallback')
    _device_checked = True


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else c_void_p(t.data_ptr())


def stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def host_doubles(values):
    if values is None:
        return None
    arr = (c_double * len(values))(*[float(v) for v in values])
    return arr


def version():
    return lib.so_version().decode()


def prof_enable(on=True):
    """per-kernel timing of every labelled launch inside the library (bench.py's profiling pass)"""
    check(lib.so_prof_enable(1 if on else 0), 'so_prof_enable')


def prof_report():
    """{label: (launches, total_ms)} collected since prof_enable(True); synchronises"""
    n = lib.so_prof_report(None, 0)
    buf = ctypes.create_string_buffer(int(n) + 16)
    lib.so_prof_report(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        label, cnt, ms = line.split('\t')
        out[label] = (int(cnt), float(ms))
    return out
