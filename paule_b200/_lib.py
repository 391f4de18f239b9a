"""ctypes binding of the C ABI in include/paule_b200.h (libpaule_b200.so, built in-tree by build.sh).

The shared library is the product; this module only loads it, declares argument types and turns
non-zero return codes into exceptions.  There is no fallback: if the library is missing or no
sm_100 device is present the public API raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# PAULE_B200_LIB selects another build of the same library (A/B timing of kernel variants); default: the in-tree build
LIB_PATH = os.environ.get("PAULE_B200_LIB") or os.path.join(_HERE, "lib", "libpaule_b200.so")

i64, i32, f32, vp, sz = C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_size_t


class LstmLayer(C.Structure):
    """struct paule_lstm_layer"""
    _fields_ = [("w_ih", vp), ("w_hh", vp), ("w_ih_t", vp), ("w_hh_t", vp), ("bias", vp), ("packed", vp),
                ("packed_ih", vp), ("packed_ih_t", vp), ("input_size", i64)]


class Plan(C.Structure):
    """struct paule_plan"""
    _fields_ = [
        ("B", i64), ("T", i64), ("H", i64), ("C", i64), ("Cm", i64), ("S", i64),
        ("objective", i32), ("math", i32), ("smiling", i32), ("log_slot_count", i32),
        ("log_semantics", i32), ("reserved0", i32),
        ("fwd", LstmLayer), ("post_w", vp), ("post_w_t", vp), ("post_b", vp), ("post_packed", vp),
        ("emb0", LstmLayer), ("emb1", LstmLayer), ("head_w", vp), ("head_w_t", vp), ("head_b", vp),
        ("cp", vp), ("adam_m", vp), ("adam_v", vp), ("step_count", vp),
        ("target_mel", vp), ("target_sv", vp), ("past_cp", vp), ("past_T", i64),
        ("lr", f32), ("beta1", f32), ("beta2", f32), ("eps", f32), ("clamp", f32),
        ("loss_log", vp), ("pred_mel", vp), ("pred_sv", vp), ("grad_out", vp),
        ("workspace", vp), ("workspace_bytes", sz),
        ("word_frames", vp),
        ("cls_w", vp), ("cls_b", vp), ("extra_terms", vp), ("extra_grad", vp), ("aux_log", vp),
        ("bwd_fused_packed", vp),
        ("post_t_packed", vp),
    ]


# name -> (restype, argtypes); must list EVERY symbol include/paule_b200.h declares (tests check this)
SIGNATURES = {
    "paule_version": (C.c_int, []),
    "paule_error_string": (C.c_char_p, [C.c_int]),
    "paule_last_cuda_error": (C.c_char_p, []),
    "paule_device_check": (C.c_int, []),
    "paule_linear_f32": (C.c_int, [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, C.c_int, vp]),
    "paule_gemm_tn_f32": (C.c_int, [vp, vp, vp, i64, i64, i64, i64, i64, C.c_int, vp]),
    "paule_colsum_f32": (C.c_int, [vp, vp, i64, i64, i64, C.c_int, vp]),
    "paule_transpose_btc": (C.c_int, [vp, vp, i64, i64, i64, vp]),
    "paule_lstm_seq_fwd_f32": (C.c_int, [vp, vp, vp, vp, i64, i64, i64, vp]),
    "paule_lstm_seq_bwd_f32": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp, i64, i64, i64, vp]),
    "paule_plan_loss_scratch_floats": (sz, [i64, i64]),
    "paule_plan_loss_f32": (C.c_int, [vp] * 10 + [i64] * 6 + [C.c_int, vp]),
    "paule_step_tick": (C.c_int, [vp, vp]),
    "paule_adam_clamp_f32": (C.c_int, [vp, vp, vp, vp, vp, vp, f32, f32, f32, f32, f32, C.c_int, vp, i64, i64, i64,
                                       i64, vp]),
    "paule_melconv_res_f32": (C.c_int, [vp, vp, vp, vp, i64, i64, i64, vp]),
    "paule_vel_acc_f32": (C.c_int, [vp, vp, i64, i64, i64, vp]),
    "paule_upsample_smooth_f32": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, vp, vp, i64, i64, i64, vp]),
    "paule_tc_packed_lstm_bytes": (sz, [i64, i64]),
    "paule_tc_pack_lstm": (C.c_int, [vp, vp, vp, i64, i64, vp]),
    "paule_tc_gemm_packed_bytes": (sz, [i64, i64]),
    "paule_tc_gemm_pack": (C.c_int, [vp, vp, i64, i64, vp]),
    "paule_tc_gemm_img": (C.c_int, [vp, vp, vp, vp, i64, i64, i64, i64, C.c_int, vp]),
    "paule_tc_gemm_packed_bytes_k64": (sz, [i64]),
    "paule_tc_gemm_pack_k64": (C.c_int, [vp, vp, i64, i64, vp]),
    "paule_tc_a_image_bytes": (sz, [i64, i64]),
    "paule_tc_a_image": (C.c_int, [vp, vp, i64, i64, i64, vp]),
    "paule_tc_gemm_img_k64": (C.c_int, [vp, vp, vp, vp, i64, i64, i64, C.c_int, vp]),
    "paule_tc_rnn_xchg_bytes": (sz, [i64]),
    "paule_tc_rnn_pass_plan": (i32, [i64, i32, vp, vp, i32]),
    "paule_tc_img_seq_bytes": (sz, [i64, i64, i64]),
    "paule_tc_lstm_seq_fwd": (C.c_int, [vp, vp, vp, vp, vp, vp, i64, i64, C.c_int, vp]),
    "paule_tc_x_image_bytes": (sz, [i64, i64]),
    "paule_tc_x_image": (C.c_int, [vp, vp, i64, i64, i64, vp]),
    "paule_tc_lstm_seq_fwd_x": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, C.c_int, vp]),
    "paule_tc_lstm_seq_bwd_img": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp, vp, i64, i64, C.c_int, vp]),
    "paule_tc_lstm_seq_bwd": (C.c_int, [vp, vp, vp, vp, C.c_int, vp, vp, vp, i64, i64, C.c_int, vp]),
    "paule_plan_status_offset": (sz, [i64, i64, i64, i64, i64, i64, C.c_int]),
    "paule_plan_workspace_bytes": (sz, [i64, i64, i64, i64, i64, i64, C.c_int]),
    "paule_plan_grad_lstm_offset": (sz, [i64, i64, i64, i64, i64, i64, C.c_int]),
    "paule_plan_embed": (C.c_int, [C.POINTER(Plan), vp, vp, vp]),
    "paule_plan_step_launches": (i64, [C.POINTER(Plan)]),
    "paule_plan_forward": (C.c_int, [C.POINTER(Plan), vp]),
    "paule_plan_step": (C.c_int, [C.POINTER(Plan), vp]),
}

_lib: Optional[C.CDLL] = None


class PauleB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libpaule_b200.so (once).  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PauleB200Error(
            f"{LIB_PATH} is missing: build it with ./build.sh (nvcc, sm_100a). paule_b200 has no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str = "") -> None:
    if code == 0:
        return
    lib = load()
    msg = lib.paule_error_string(code).decode()
    if code == 2:
        msg += ": " + lib.paule_last_cuda_error().decode()
    raise PauleB200Error(f"{what or 'paule_b200'} failed ({code}): {msg}")


def require_device() -> None:
    """Fail loudly when there is no B200: the product path never silently runs elsewhere."""
    check(load().paule_device_check(), "paule_device_check")
