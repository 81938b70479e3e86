"""ctypes binding of libustrun_sm100.so (the C ABI declared in include/ustrun.h).

There is no CPU or PyTorch fallback: if the shared library is missing the import fails loudly, and
every entry point raises ``RuntimeError`` on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libustrun_sm100.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
SIMT, TCGEN05 = 0, 1
MAX_PARTS = 1280
OPT_CHUNK = 4096


class UstrunError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C ust-run_b200/csrc` (there is no CPU/PyTorch fallback for the UST-RUN kernels)")

lib = C.CDLL(LIB_PATH)

p, i32, i64, f32, f64 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
ip = C.POINTER(C.c_int)

_SIGS = {
    "ustrun_abi_version": [],
    "ustrun_device_supported": [],
    "ustrun_nchw_to_nhwc": [p, p, i32, i32, i32, i32, i32, i32, p],
    "ustrun_nhwc_to_nchw": [p, i32, i32, p, i32, i32, i32, i32, p],
    "ustrun_pack_conv_weight": [p, p, p, i32, i32, i32, i32, p],
    "ustrun_pack_convT_weight": [p, p, p, i32, i32, i32, p],
    "ustrun_pack_weights_multi": [p, p, p, i32, i32, p],
    "ustrun_conv_fwd": [i32, p, i32, p, p, p, i32, i32, i32, i32, i32, i32, i32, i32, i32, p, ip, p],
    "ustrun_conv_wgrad": [i32, p, i32, p, i32, p, i32, i32, i32, i32, i32, i32, i32, i32, p, i64, p],
    "ustrun_convT2x2_fwd": [i32, p, i32, p, p, p, i32, i32, i32, i32, i32, i32, i32, p],
    "ustrun_convT2x2_dgrad": [i32, p, i32, p, p, i32, i32, i32, i32, i32, i32, i32, p],
    "ustrun_convT2x2_wgrad": [i32, p, i32, p, i32, p, i32, i32, i32, i32, i32, i32, i32, p, i64, p],
    "ustrun_channel_sum": [p, i32, i32, i64, i32, p, i32, p, p],
    "ustrun_bn_reduce_partials": [p, i32, i32, p, p],
    "ustrun_bn_finalize": [p, i32, i32, f64, p, p, p, p, p, p, f32, f32, i32, p, p, p, p, p, p],
    "ustrun_bn_running_update": [p, i32, i32, p],
    "ustrun_bn_act_fwd": [p, i32, p, p, i32, p, i32, p, i32, i32, i32, i32, i32, i32, p],
    "ustrun_bn_bwd_reduce": [p, i32, p, i32, p, p, p, p, i32, i32, i64, i32, p, ip, p],
    "ustrun_bn_bwd_finalize": [p, i32, i32, f64, p, p, p, p, i32, f32, p, p],
    "ustrun_bn_bwd_apply": [p, i32, p, i32, p, p, p, p, p, i32, p, i32, i32, i64, i32, p],
    "ustrun_maxpool_bwd": [p, i32, p, i32, p, i32, p, i32, i32, i32, i32, i32, i32, p],
    "ustrun_upsample2x_fwd": [p, i32, p, i32, i32, i32, i32, i32, i32, i32, p],
    "ustrun_upsample2x_bwd": [p, i32, p, i32, i32, i32, i32, i32, i32, i32, p],
    "ustrun_pseudo_label_softmax": [p] * 8 + [f32, i32, i32, i32, i32] + [p] * 9 + [p],
    "ustrun_pseudo_label_sigmoid": [p] * 8 + [f32, f32, i32, i32, i32, i32] + [p] * 9 + [p],
    "ustrun_mix_to_nhwc": [p, p, p, p, p, i32, i32, i32, i32, i32, i32, p],
    "ustrun_mix_any_to_nhwc": [p, p, p, p, p, p, p, i32, i32, i32, i32, i32, i32, p],
    "ustrun_normalize_u8_to_nchw": [p, p, i32, i32, i32, i32, p],
    "ustrun_ce_dice_softmax_fwd": [p, p, p, i32, i32, i32, i32, f32, f32, p, p, p, p, p],
    "ustrun_ce_dice_softmax_partials": [p, p, p, i32, i32, i32, i32, p, ip, p],
    "ustrun_ce_dice_softmax_finalize": [p, i32, i32, f64, f32, f32, p, p, p, p],
    "ustrun_reduce_rows": [p, i32, i32, p, p],
    "ustrun_bce_dice_sigmoid_partials": [p, p, p, i32, i32, i32, i32, p, ip, p],
    "ustrun_bce_dice_sigmoid_finalize": [p, i32, f64, f32, f32, p, p, p],
    "ustrun_ce_dice_softmax_bwd": [p, p, p, i32, i32, i32, i32, p, p, f32, p, i32, p],
    "ustrun_bce_dice_sigmoid_fwd": [p, p, p, i32, i32, i32, i32, f32, f32, p, p, p, p],
    "ustrun_bce_dice_sigmoid_bwd": [p, p, p, i32, i32, i32, i32, p, p, f32, p, i32, p],
    "ustrun_bn_finalize_peer": [p, i32, i32, f64, p, p, p, p, p, p, f32, f32, p, p, p, p, p, p, i32, i32, C.c_uint, p, p, p],
    "ustrun_bn_bwd_finalize_peer": [p, i32, i32, f64, p, p, p, p, i32, p, p, i32, i32, C.c_uint, p, p, p],
    "ustrun_sgd_ema_multi": [p, p, p, i32, f32, f32, f32, f32, f32, i32, i32, p],
    "ustrun_sgd_ema_multi_dev": [p, p, p, i32, p, f32, f32, i32, i32, p],
    "ustrun_fft_amp_mix": [p, p, p, f64, p, i32, i32, i32, i32, p, i64, p],
    "ustrun_hardness": [p, p, i32, i32, i32, i32, i32, p, p, p, p, p],
    "ustrun_bank_update": [p, i32, p, p, p, p, p, p, p, p, p, p, p, p, p, i32, f64, i64, i64, p, p],
    "ustrun_bank_choice": [p, i32, i32, p, p, p, p, p],
    "ustrun_lq_select": [p, p, p, p, p, p, p, i64, i64, p],
    "ustrun_cover_box": [p, p, p, p, i32, i32, p, p, p],
    "ustrun_encode_labels": [p, i32, i32, i32, i32, p, p],
    "ustrun_predict": [p, i32, i32, i32, i32, i32, p, p],
    "ustrun_seg_metrics": [p, p, i32, i32, i32, i32, p, p, p],
}
for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = i32
lib.ustrun_last_error_string.restype = C.c_char_p
lib.ustrun_last_error_string.argtypes = []
lib.ustrun_conv_wgrad_workspace_bytes.restype = i64
lib.ustrun_conv_wgrad_workspace_bytes.argtypes = [i32, i32, i32, i32, i32, i32, i32]

lib.ustrun_tc_plan_query.restype = i32
lib.ustrun_tc_plan_query.argtypes = [i32, i32, i32, i32, i32, i32, i32, C.POINTER(C.c_int)]

lib.ustrun_peer_buffer_bytes.restype = i64
lib.ustrun_peer_buffer_bytes.argtypes = []
lib.ustrun_fft_amp_mix_workspace_bytes.restype = i64
lib.ustrun_fft_amp_mix_workspace_bytes.argtypes = [i32, i32, i32, i32, f64]

EXPORTS = sorted(list(_SIGS) + ["ustrun_last_error_string", "ustrun_conv_wgrad_workspace_bytes", "ustrun_tc_plan_query", "ustrun_peer_buffer_bytes",
                                 "ustrun_fft_amp_mix_workspace_bytes"])


def last_error() -> str:
    return (lib.ustrun_last_error_string() or b"").decode()


_FN = {_name: getattr(lib, _name) for _name in _SIGS}        # bound once: the step makes ~1300 calls


def call(name: str, *args) -> None:
    rc = _FN[name](*args)
    if rc != 0:
        raise UstrunError(f"{name} failed (code {rc}): {last_error()}")


def require_device() -> None:
    """Fail loudly unless a compute-capability-10.x GPU is the current CUDA device."""
    rc = lib.ustrun_device_supported()
    if rc != 1:
        raise UstrunError("libustrun_sm100 needs an sm_100 (B200) CUDA device and has no CPU path: " + (last_error() or f"rc={rc}"))
