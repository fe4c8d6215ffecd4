"""ctypes binding of libcoskad_b200.so -- the ONLY compute backend of this package.

There is no CPU or PyTorch fallback: if the shared library is missing, or no sm_100 device is
visible when a context is created, the calls below raise.  Signatures mirror include/coskad_b200.h.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libcoskad_b200.so')

c_float_p = C.c_void_p     # device pointers travel as integers (tensor.data_ptr())
c_ctx_p = C.c_void_p

# flavours / ops (include/coskad_b200.h)
SCORE_NONE, SCORE_POINCARE, SCORE_POINCARE_NOPROJ, SCORE_EUCLID, SCORE_COSINE, SCORE_POINCARE_HM = range(6)
MAP_EXPMAP0, MAP_PROJECT, MAP_EXPMAP0_PROJECT, MAP_EXPMAP0_HM, MAP_PROJECT_HM, MAP_L2NORMALIZE = range(6)


class LayerParams(C.Structure):
    """coskad_layer_params"""
    _fields_ = [('c_in', C.c_int32), ('c_out', C.c_int32)] + [
        (n, C.c_void_p) for n in ('A', 'T', 'w1', 'b1', 'bn1_w', 'bn1_b', 'bn1_rm', 'bn1_rv',
                                  'w2', 'b2', 'bn2_w', 'bn2_b', 'bn2_rm', 'bn2_rv', 'prelu')]


# name -> (restype, argtypes); every function include/coskad_b200.h declares
SIGNATURES = {
    'coskad_abi_version': (C.c_int, []),
    'coskad_create': (C.c_int, [C.POINTER(c_ctx_p), C.c_int, C.c_int, C.c_int]),
    'coskad_destroy': (C.c_int, [c_ctx_p]),
    'coskad_last_error': (C.c_char_p, [c_ctx_p]),
    'coskad_set_encoder': (C.c_int, [c_ctx_p, C.c_int, C.POINTER(LayerParams), c_float_p, c_float_p, C.c_int, C.c_void_p]),
    'coskad_set_decoder': (C.c_int, [c_ctx_p, c_float_p, c_float_p, C.c_int, C.c_int, C.POINTER(LayerParams), C.c_void_p]),
    'coskad_encode_score_fwd': (C.c_int, [c_ctx_p, C.c_int, c_float_p, c_float_p, C.c_int64, c_float_p, c_float_p, C.c_void_p]),
    'coskad_encode_score_traj_fwd': (C.c_int, [c_ctx_p, C.c_int, c_float_p, C.c_int64, C.c_void_p, C.c_void_p, c_float_p, C.c_int,
                                               c_float_p, C.c_int64, c_float_p, c_float_p, C.c_void_p]),
    'coskad_set_fused_impl': (C.c_int, [c_ctx_p, C.c_int]),
    'coskad_autoencode_score_fwd': (C.c_int, [c_ctx_p, c_float_p, c_float_p, C.c_int64, c_float_p, c_float_p, c_float_p, c_float_p, C.c_void_p]),
    'coskad_geom_map': (C.c_int, [c_ctx_p, C.c_int, c_float_p, C.c_int64, C.c_int, c_float_p, C.c_void_p]),
    'coskad_dist': (C.c_int, [c_ctx_p, C.c_int, c_float_p, c_float_p, C.c_int, C.c_int64, C.c_int, c_float_p, C.c_void_p]),
    'coskad_geom_map_bwd': (C.c_int, [c_ctx_p, C.c_int, c_float_p, c_float_p, C.c_int64, C.c_int, c_float_p, C.c_void_p]),
    'coskad_dist_bwd': (C.c_int, [c_ctx_p, C.c_int, c_float_p, c_float_p, C.c_int, c_float_p, C.c_int64, C.c_int, c_float_p,
                                  c_float_p, C.c_void_p]),
    'coskad_dist0': (C.c_int, [c_ctx_p, c_float_p, C.c_int64, C.c_int, c_float_p, C.c_void_p]),
    'coskad_ps_sample': (C.c_int, [c_ctx_p, c_float_p, c_float_p, c_float_p, C.c_int64, C.c_int, c_float_p, C.c_void_p]),
    'coskad_poincare_score_bwd': (C.c_int, [c_ctx_p, c_float_p, c_float_p, c_float_p, C.c_int64, C.c_int, C.c_int, c_float_p, C.c_void_p]),
    'coskad_center_partial': (C.c_int, [c_ctx_p, C.c_int, c_float_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    'coskad_center_finalize': (C.c_int, [c_ctx_p, C.c_int, C.c_void_p, C.c_int, C.c_float, c_float_p, C.c_void_p]),
    'coskad_mahalanobis': (C.c_int, [c_ctx_p, c_float_p, c_float_p, c_float_p, C.c_int64, C.c_int, c_float_p, C.c_void_p]),
    'coskad_mahalanobis_bwd': (C.c_int, [c_ctx_p, c_float_p, c_float_p, c_float_p, c_float_p, C.c_int64, C.c_int, c_float_p, C.c_void_p]),
    'coskad_cov_partial': (C.c_int, [c_ctx_p, c_float_p, c_float_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    'coskad_frame_aggregate': (C.c_int, [c_ctx_p, c_float_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    'coskad_score_process': (C.c_int, [c_ctx_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    'coskad_train_contract_fwd': (C.c_int, [c_ctx_p] + [c_float_p] * 3 + [C.c_int64] + [c_float_p] * 2 + [C.c_void_p]),
    'coskad_train_contract_bwd': (C.c_int, [c_ctx_p] + [c_float_p] * 6 + [C.c_int64] + [c_float_p] * 3 + [C.c_void_p]),
    'coskad_train_mix_fwd': (C.c_int, [c_ctx_p] + [c_float_p] * 6 + [C.c_int64, C.c_int, C.c_int] + [c_float_p] * 3 + [C.c_void_p]),
    'coskad_train_bn_finalize': (C.c_int, [c_ctx_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float] + [c_float_p] * 5
                                 + [C.c_void_p, C.c_void_p, C.c_void_p]),
    'coskad_train_mix_fwd_bn': (C.c_int, [c_ctx_p] + [c_float_p] * 6 + [C.c_int64, C.c_int, C.c_int] + [c_float_p] * 2
                                + [C.c_float, C.c_float] + [c_float_p] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p]),
    'coskad_train_bn_prelu_bwd_grads': (C.c_int, [c_ctx_p] + [c_float_p] * 9 + [C.c_int64, C.c_int, C.c_void_p] + [c_float_p] * 5
                                        + [C.c_void_p]),
    'coskad_adam_step': (C.c_int, [c_ctx_p] + [c_float_p] * 4 + [C.c_int64, c_float_p, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                   c_float_p, C.c_void_p]),
    'coskad_train_bn_param_grads': (C.c_int, [c_ctx_p, C.c_void_p, C.c_int] + [c_float_p] * 5 + [C.c_void_p]),
    'coskad_train_bn_prelu_fwd': (C.c_int, [c_ctx_p] + [c_float_p] * 8 + [C.c_int64, C.c_int, c_float_p, C.c_void_p]),
    'coskad_train_bn_prelu_bwd': (C.c_int, [c_ctx_p] + [c_float_p] * 9 + [C.c_int64, C.c_int, C.c_void_p, c_float_p, c_float_p, C.c_void_p]),
    'coskad_train_mix_bwd': (C.c_int, [c_ctx_p] + [c_float_p] * 6 + [C.c_int64, C.c_int, C.c_int] + [c_float_p] * 6 + [C.c_void_p]),
    'coskad_train_mix_bwd_tc': (C.c_int, [c_ctx_p] + [c_float_p] * 9 + [C.c_void_p] + [c_float_p] * 4 + [C.c_int64, C.c_int, C.c_int]
                                + [c_float_p] * 8 + [C.c_void_p]),
    'coskad_set_train_impl': (C.c_int, [c_ctx_p, C.c_int]),
    'coskad_train_linear': (C.c_int, [c_ctx_p, C.c_int, c_float_p, c_float_p, c_float_p, C.c_int, c_float_p, C.c_int64, C.c_int, C.c_int, c_float_p, C.c_void_p]),
    'coskad_train_col_sum': (C.c_int, [c_ctx_p, c_float_p, C.c_int64, C.c_int, c_float_p, C.c_void_p]),
    'coskad_measure_fp32_peak': (C.c_int, [c_ctx_p, C.POINTER(C.c_double), C.c_void_p]),
    'coskad_measure_tf32_peak': (C.c_int, [c_ctx_p, C.POINTER(C.c_double), C.c_void_p]),
    'coskad_launch_count': (C.c_int64, [c_ctx_p]),
    'coskad_debug_fused_stage': (C.c_int, [c_ctx_p, C.c_int, c_float_p, C.c_int64, C.c_int, c_float_p, C.c_void_p]),
    'coskad_debug_fused_floats': (C.c_int, []),
    'coskad_debug_tc_mix': (C.c_int, [c_ctx_p, c_float_p, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, c_float_p, C.c_void_p]),
    'coskad_fused_tile_windows': (C.c_int, []),
}

_lib: Optional[C.CDLL] = None


class CoskadError(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the in-tree library and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CoskadError(f'{LIB_PATH} not found: build it with `python -m coskad_b200._build` '
                          '(there is no CPU/PyTorch fallback)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(t) -> Optional[int]:
    """device pointer of a torch tensor (None -> NULL)"""
    if t is None:
        return None
    return t.data_ptr()


class Context:
    """RAII wrapper of coskad_ctx, one per (device, thread)."""

    def __init__(self, device: int = 0, n_frames: int = 12, n_joints: int = 17):
        self.lib = load()
        h = c_ctx_p()
        rc = self.lib.coskad_create(C.byref(h), int(device), int(n_frames), int(n_joints))
        if rc != 0:
            msg = self.lib.coskad_last_error(None)
            raise CoskadError(f'coskad_create failed ({rc}): {msg.decode() if msg else ""}')
        self.h = h
        self.device = int(device)

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.lib.coskad_last_error(self.h)
            raise CoskadError(f'{what} failed ({rc}): {msg.decode() if msg else ""}')

    def launch_count(self) -> int:
        return int(self.lib.coskad_launch_count(self.h))

    def close(self) -> None:
        if getattr(self, 'h', None):
            self.lib.coskad_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}


def context(device: int = 0) -> Context:
    """process-wide default context of a device"""
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx


def stream_ptr(device=None) -> int:
    import torch
    return torch.cuda.current_stream(device).cuda_stream
