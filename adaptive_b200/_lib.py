"""ctypes binding of ``libadaptive_sm100.so`` (C ABI declared in ``include/adaptive_b200.h``).

There is no CPU fallback and no alternative backend: if the shared library is missing or a
call fails, a ``RuntimeError`` is raised with the library's own error message.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadaptive_sm100.so")

P = c_void_p  # every device pointer travels as void*


class AADims(Structure):
    _fields_ = [(n, c_int32) for n in ("B", "T", "k", "a", "H", "E", "Vc", "precision")]


PREC_FP32, PREC_BF16, PREC_TF32X3 = 0, 1, 2


WEIGHT_FIELDS = ("embed", "w_ih", "w_hh", "b_ih", "b_hh", "sen_wx", "sen_wh",
                 "att_wv", "att_wg", "att_ws", "att_wh", "mlp_w", "mlp_b")

# reference Decoder.state_dict() key -> aa_weights field
KEY_TO_FIELD = {
    "embed.weight": "embed",
    "LSTM.weight_ih_l0": "w_ih",
    "LSTM.weight_hh_l0": "w_hh",
    "LSTM.bias_ih_l0": "b_ih",
    "LSTM.bias_hh_l0": "b_hh",
    "adaptive.sentinel.affine_x.weight": "sen_wx",
    "adaptive.sentinel.affine_h.weight": "sen_wh",
    "adaptive.atten.affine_v.weight": "att_wv",
    "adaptive.atten.affine_g.weight": "att_wg",
    "adaptive.atten.affine_s.weight": "att_ws",
    "adaptive.atten.affine_h.weight": "att_wh",
    "adaptive.mlp.weight": "mlp_w",
    "adaptive.mlp.bias": "mlp_b",
}


class AAWeights(Structure):
    _fields_ = [(n, P) for n in WEIGHT_FIELDS]


class AAWeightGrads(Structure):
    _fields_ = [(n, P) for n in WEIGHT_FIELDS]


class AAEncDims(Structure):
    _fields_ = [(n, c_int32) for n in ("B", "C", "hw", "H", "E", "precision")]


ENC_FIELDS = ("wa", "ba", "wb", "bb", "wh0", "bh0", "wc0", "bc0")
# AttentiveCNN.state_dict() key -> aa_enc_weights field
ENC_KEY_TO_FIELD = {"affine_a.weight": "wa", "affine_a.bias": "ba", "affine_b.weight": "wb", "affine_b.bias": "bb",
                    "affine_h0.weight": "wh0", "affine_h0.bias": "bh0", "affine_c0.weight": "wc0", "affine_c0.bias": "bc0"}


class AAEncWeights(Structure):
    _fields_ = [(n, P) for n in ENC_FIELDS]


class AAEncWeightGrads(Structure):
    _fields_ = [(n, P) for n in ENC_FIELDS]


# name -> (restype, argtypes); mirrors include/adaptive_b200.h one to one
_D, _W, _G = POINTER(AADims), POINTER(AAWeights), POINTER(AAWeightGrads)
_ED, _EW, _EG = POINTER(AAEncDims), POINTER(AAEncWeights), POINTER(AAEncWeightGrads)
SIGNATURES = {
    "aa_version": (c_int, []),
    "aa_last_error": (c_char_p, []),
    "aa_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "aa_launch_count": (ctypes.c_longlong, []),
    "aa_profile_enable": (c_int, [c_int]),
    "aa_profile_reset": (c_int, []),
    "aa_profile_count": (c_int, []),
    "aa_profile_get": (c_int, [c_int, ctypes.c_char_p, c_int, POINTER(ctypes.c_double), POINTER(c_int)]),
    "aa_debug_set_trace_buffer": (c_int, [P]),
    "aa_debug_set_decode_atten_simple": (c_int, [c_int]),
    "aa_debug_set_persist_trace": (c_int, [P]),
    "aa_debug_set_atten_sequential": (c_int, [c_int]),
    "aa_debug_set_decode_argmax_refine": (c_int, [c_int]),
    "aa_debug_refine_pairs": (ctypes.c_longlong, [c_int]),
    "aa_debug_refine_units": (c_int, [ctypes.c_void_p, c_int]),
    "aa_debug_set_gemm_splitk": (c_int, [c_int]),
    "aa_debug_set_gemm_pair": (c_int, [c_int]),
    "aa_debug_set_bptt_ksplit": (c_int, [c_int]),
    "aa_debug_set_lstm_cluster": (c_int, [c_int, c_int]),
    "aa_linear_forward": (c_int, [c_int, c_int, c_int, P, c_int64, P, c_int64, P, P, c_int64, P]),
    "aa_gemm": (c_int, [c_int, c_int, c_int, c_int, P, c_int64, c_int, P, c_int64, c_int, P, c_int64, ctypes.c_float, P, P,
                        c_int64, P]),
    "aa_split_tf32": (c_int, [P, c_int64, c_int64, c_int, P, c_int, P]),
    "aa_gemm_split3": (c_int, [c_int, c_int, c_int, P, P, P, P, c_int64, P]),
    "aa_precompute_P": (c_int, [_D, P, P, P, P]),
    "aa_sentinel_forward": (c_int, [_D, P, P, P, P, P, P, P, P]),
    "aa_atten_workspace_bytes": (c_size_t, [_D]),
    "aa_atten_forward": (c_int, [_D, P, P, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "aa_adaptive_workspace_bytes": (c_size_t, [_D]),
    "aa_adaptive_forward": (c_int, [_D, _W, P, P, P, P, P, P, P, P, c_size_t, P]),
    "aa_decoder_saved_bytes": (c_size_t, [_D]),
    "aa_decoder_bwd_scratch_bytes": (c_size_t, [_D]),
    "aa_decoder_forward": (c_int, [_D, _W, P, P, P, P, P, P, P, P, P, P, P, c_size_t, P]),
    "aa_decoder_backward": (c_int, [_D, _W, P, P, P, P, P, P, P, P, c_size_t, P, P, P, P, P, _G, P, P, P, P, P,
                                    c_size_t, P]),
    "aa_decoder_backward_hooked": (c_int, [_D, _W, P, P, P, P, P, P, P, P, c_size_t, P, P, P, P, P, _G, P, P, P, P, P,
                                           c_size_t, P, POINTER(c_void_p), c_void_p, c_void_p]),
    "aa_decoder_forward_packed": (c_int, [_D, _W, P, P, P, P, P, P, c_int64, P, P, P, P, P, P, c_size_t, P]),
    "aa_decoder_backward_packed": (c_int, [_D, _W, P, P, P, P, P, P, P, P, c_size_t, P, c_int64, P, P, P, P, P, _G, P, P, P, P, P,
                                           c_size_t, P, POINTER(c_void_p), c_void_p, c_void_p, P]),
    "aa_allreduce_flag_bytes": (c_size_t, []),
    "aa_allreduce_sum_f32": (c_int, [POINTER(c_void_p), P, ctypes.c_longlong, c_int, c_int, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, P]),
    "aa_allreduce_sum_bf16": (c_int, [POINTER(c_void_p), P, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, ctypes.c_longlong,
                                      ctypes.c_longlong, c_int, c_int, P]),
    "aa_decoder_forward_loss": (c_int, [_D, _W, P, P, P, P, P, P, c_int64, P, c_int64, P, P, P, P, P, P, c_size_t, P]),
    "aa_decoder_loss_grad_scale": (c_int, [_D, P, c_size_t, c_int64, P, P]),
    "aa_clip_adam_step": (c_int, [P, c_int, c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                  ctypes.c_float, c_int, P, P, P]),
    "aa_pack_rows": (c_int, [P, c_int64, P, c_int64, P, P]),
    "aa_unpack_rows": (c_int, [P, c_int64, P, c_int64, c_int64, P, P]),
    "aa_cross_entropy": (c_int, [P, c_int64, c_int64, P, P, P, P]),
    "aa_scale_unless_one": (c_int, [P, P, c_int64, P, P]),
    "aa_cross_entropy_mirror": (c_int, [P, c_int64, c_int64, P, c_int64, P, P, P, POINTER(c_int), P]),
    "aa_copy_multi": (c_int, [c_int, P, P, P, P]),
    "aa_cross_entropy_denom": (c_int, [P, c_int64, c_int64, P, c_int64, P, P, P]),
    "aa_encoder_saved_bytes": (c_size_t, [_ED]),
    "aa_encoder_bwd_scratch_bytes": (c_size_t, [_ED, c_int]),
    "aa_encoder_forward": (c_int, [_ED, _EW, P, P, P, P, P, P, c_size_t, P]),
    "aa_encoder_backward": (c_int, [_ED, _EW, P, c_size_t, P, P, P, P, P, P, P, P, _EG, P, P, c_size_t, P]),
    "aa_decode_workspace_bytes": (c_size_t, [_D, c_int]),
    "aa_greedy_decode": (c_int, [_D, _W, P, P, P, P, c_int, P, P, P, P, P, c_size_t, P]),
    "aa_decode_persistent_workspace_bytes": (c_size_t, [_D]),
    "aa_decode_persistent_supported": (c_int, [_D]),
    "aa_decode_persistent": (c_int, [_D, _W, P, P, P, P, c_int, P, P, P, c_int, P, P, c_size_t, P]),
    "aa_beam_decode": (c_int, [_D, _W, P, P, P, P, c_int, c_int, P, P, P, P, P, c_size_t, P]),
}

_lib = None


def load():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "adaptive_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C adaptive_b200/csrc`. There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale -> loud
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().aa_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def version() -> int:
    return load().aa_version()


def launch_count() -> int:
    """Kernels launched by libadaptive_sm100 in this process so far."""
    return int(load().aa_launch_count())


def profile_enable(on: bool = True):
    check(load().aa_profile_enable(1 if on else 0), "aa_profile_enable")


def profile_reset():
    check(load().aa_profile_reset(), "aa_profile_reset")


def profile_report():
    """{kernel tag: (total_ms, launches)} for the launches recorded since the last reset."""
    lib = load()
    out = {}
    for i in range(lib.aa_profile_count()):
        name = ctypes.create_string_buffer(64)
        ms, n = ctypes.c_double(0), c_int(0)
        check(lib.aa_profile_get(i, name, 64, ctypes.byref(ms), ctypes.byref(n)), "aa_profile_get")
        out[name.value.decode()] = (ms.value, n.value)
    return out
