"""ctypes binding of librtsds_b200.so (include/rtsds_b200.h).

The library is the ONLY compute path: if it is missing or fails to load the
import raises — there is no CPU / eager-PyTorch fallback (BASELINE.json
north_star).  `build()` compiles it in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "librtsds_b200.so"

F32, BF16, F16 = 0, 1, 2
ABI_VERSION = 3
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2


class ConvDesc(C.Structure):
    """RtsdsConvDesc (include/rtsds_b200.h)."""

    _fields_ = [
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("cin", C.c_int), ("in_ld", C.c_int),
        ("cout", C.c_int), ("out_ld", C.c_int), ("res_ld", C.c_int),
        ("kh", C.c_int), ("kw", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("dil", C.c_int),
        ("oh", C.c_int), ("ow", C.c_int),
        ("act", C.c_int), ("slope", C.c_float),
        ("in_dtype", C.c_int), ("out_dtype", C.c_int), ("split_k", C.c_int),
    ]


class PackJob(C.Structure):
    """RtsdsPackJob (include/rtsds_b200.h)."""

    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p), ("cout", C.c_int), ("cin", C.c_int), ("cin_pad", C.c_int),
                ("taps", C.c_int), ("cout_pad", C.c_int), ("kind", C.c_int), ("ck", C.c_int)]


class UnpackJob(C.Structure):
    """RtsdsUnpackJob (include/rtsds_b200.h)."""

    _fields_ = [("dw_packed", C.c_void_p), ("grad", C.c_void_p), ("cout", C.c_int), ("cin", C.c_int), ("cin_src", C.c_int),
                ("taps", C.c_int), ("accumulate", C.c_int)]


class OptJob(C.Structure):
    """RtsdsOptJob (include/rtsds_b200.h)."""

    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("numel", C.c_int64),
                ("group", C.c_int), ("taps", C.c_int), ("cout", C.c_int), ("cin", C.c_int), ("cout_pad", C.c_int),
                ("cin_pad_fwd", C.c_int), ("cin_pad_dgrad", C.c_int), ("ck", C.c_int), ("out_fwd", C.c_void_p),
                ("out_dgrad", C.c_void_p)]


class ArmSide(C.Structure):
    """RtsdsArmSide (include/rtsds_b200.h)."""

    _fields_ = [("src", C.c_void_p), ("pooled", C.c_void_p), ("w", C.c_void_p), ("b", C.c_void_p), ("gamma", C.c_void_p),
                ("beta", C.c_void_p), ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("eps", C.c_float),
                ("out_scale", C.c_float), ("h", C.c_int), ("w_in", C.c_int), ("c", C.c_int), ("dst_coff", C.c_int),
                ("mul_pooled", C.c_int), ("pooled_parts", C.c_int)]


class OptHyper(C.Structure):
    """RtsdsOptHyper (include/rtsds_b200.h)."""

    _fields_ = [("kind", C.c_int), ("first_step", C.c_int), ("lr", C.c_float * 8), ("weight_decay", C.c_float * 8),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("inv_bias_correction1", C.c_float),
                ("inv_bias_correction2_sqrt", C.c_float), ("momentum", C.c_float)]


_P, _I, _L, _F, _D, _Z = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
_CD = C.POINTER(ConvDesc)

# name -> (restype, argtypes); must list every function include/rtsds_b200.h declares
SIGNATURES = {
    "rtsds_abi_version": (_I, []),
    "rtsds_last_error_string": (C.c_char_p, []),
    "rtsds_check_device": (_I, []),
    "rtsds_launch_count": (_L, []),
    "rtsds_set_deterministic": (None, [_I]),
    "rtsds_get_deterministic": (_I, []),
    "rtsds_confusion_hist": (_I, [_P, _P, _L, _I, _P, _P, _P]),
    "rtsds_argmax_hist": (_I, [_P, _P, _I, _I, _L, _P, _P, _P]),
    "rtsds_argmax_hist_u8": (_I, [_P, _P, _I, _I, _L, _P, _P, _P]),
    "rtsds_adaptive_avgpool_nchw_fwd": (_I, [_P, _L, _I, _I, _I, _I, _P, _P]),
    "rtsds_adaptive_avgpool_nchw_bwd": (_I, [_P, _L, _I, _I, _I, _I, _P, _P]),
    "rtsds_stem_s2d_pack_ex": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _P, _P]),
    "rtsds_stem_s2d_conv_fwd_dt": (_I, [_P, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P, _I, _I, _P]),
    "rtsds_maxpool3x3s2_fwd_ld": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_image_u8_to_f32": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "rtsds_label_resize_clamp": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _L, _L, _P, _P]),
    "rtsds_conv_cout_pad": (_I, [_I]),
    "rtsds_conv2d_tc_tune": (None, [_I, _I]),
    "rtsds_debug_conv_trace": (None, [_P]),
    "rtsds_debug_arm_trace": (None, [_P]),
    "rtsds_conv2d_tc_fwd": (_I, [_CD, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "rtsds_conv2d_tc_gap_parts": (_I, [_CD]),
    "rtsds_conv2d_tc_fwd_gap": (_I, [_CD, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "rtsds_conv2d_tc_workspace_bytes": (_Z, [_CD]),
    "rtsds_conv2d_simt_fwd": (_I, [_CD, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rtsds_pack_conv_weight": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_pack_conv_weights_batch": (_I, [_P, _I, _I, _P]),
    "rtsds_unpack_conv_wgrads_batch": (_I, [_P, _I, _P]),
    "rtsds_tapn_weights": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "rtsds_tapn_weight_grad": (_I, [_P, _I, _I, _I, _P, _P]),
    "rtsds_tapn_gather": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P]),
    "rtsds_arm_gate_resize": (_I, [C.POINTER(ArmSide), C.POINTER(ArmSide), _I, _I, _I, _I, _P, _I, _P]),
    "rtsds_ffm_head_resize": (_I, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "rtsds_tapn_gather_parts": (_I, [_I, _I, _I]),
    "rtsds_tapn_scatter": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P]),
    "rtsds_conv2d_tc_dgrad_workspace_bytes": (_Z, [_CD]),
    "rtsds_conv2d_tc_dgrad": (_I, [_CD, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "rtsds_conv2d_tc_wgrad": (_I, [_CD, _P, _P, _P, _P]),
    "rtsds_conv2d_simt_dgrad": (_I, [_CD, _P, _P, _P, _P, _I, _P]),
    "rtsds_conv2d_simt_wgrad": (_I, [_CD, _P, _P, _P, _P]),
    "rtsds_pack_conv_weight_dgrad": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_unpack_conv_wgrad": (_I, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_stem_conv_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _F, _I, _P, _I, _P, _P]),
    "rtsds_stem_pack_weights": (_I, [_P, _P, _I, _P, _P]),
    "rtsds_stem_pair_tc_fwd": (_I, [_P, _I, _I, _I, _P, _P, _P, _I, _P, _P, _I, _P, _P, _P]),
    "rtsds_stem_pair_tc_fwd_pool": (_I, [_P, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P]),
    "rtsds_stem_pair_tc_wgrad": (_I, [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "rtsds_stem_s2d_pack": (_I, [_P, _I, _I, _I, _P, _P]),
    "rtsds_stem_s2d_weight": (_I, [_P, _I, _I, _I, _P, _P]),
    "rtsds_stem_s2d_weight_grad": (_I, [_P, _I, _I, _I, _P, _P]),
    "rtsds_stem_s2d_conv_fwd": (_I, [_P, _I, _I, _I, _P, _I, _P, _P, _I, _P, _P, _I, _I, _P]),
    "rtsds_stem_s2d_conv_wgrad": (_I, [_P, _I, _I, _I, _P, _I, _I, _P, _P]),
    "rtsds_nchw_to_nhwc": (_I, [_P, _I, _I, _L, _I, _P, _I, _I, _P]),
    "rtsds_nhwc_to_nchw": (_I, [_P, _I, _I, _I, _I, _I, _L, _P, _P]),
    "rtsds_maxpool3x3s2_fwd": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_maxpool3x3s2_fwd_idx": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    "rtsds_bn_fold": (_I, [_P, _P, _P, _P, _P, _F, _I, _P, _P, _P]),
    "rtsds_bn_finalize": (_I, [_P, _D, _P, _P, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P]),
    "rtsds_scale_shift_act": (_I, [_P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _F, _I, _I, _P, _P]),
    "rtsds_bn_finalize_apply": (_I, [_P, _D, _P, _P, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _F, _I, _I, _P, _P]),
    "rtsds_bn_bwd_reduce": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _L, _I, _I, _I, _P, _P]),
    "rtsds_bn_bwd_apply": (_I, [_P, _I, _P, _I, _P, _I, _P, _P, _P, _P, _L, _I, _I, _I, _P, _I, _I, _P, _I, _P, _P, _P]),
    "rtsds_bn_bwd_reduce_rawmask": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _L, _I, _I, _P, _P]),
    "rtsds_bn_bwd_apply_rawmask": (_I, [_P, _I, _P, _I, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P, _I, _I, _P, _I, _P, _P, _P]),
    "rtsds_channel_sum": (_I, [_P, _I, _L, _I, _I, _P, _P]),
    "rtsds_maxpool3x3s2_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_maxpool3x3s2_bwd_idx": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_stem_conv_wgrad": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_resize_bwd_nhwc": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P]),
    "rtsds_gate_bwd_finish": (_I, [_P, _P, _P, _F, _I, _L, _I, _I, _P, _P]),
    "rtsds_arm_gate_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "rtsds_ffm_head_bwd": (_I, [_P, _I, _P, _I, _P, _P, _I, _L, _I, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "rtsds_resize_to_nchw_bwd": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    "rtsds_global_avgpool": (_I, [_P, _I, _L, _I, _I, _I, _P, _P]),
    "rtsds_arm_gate": (_I, [_P, _P, _P, _P, _P, _P, _P, _F, _F, _I, _I, _I, _P, _P, _P, _P, _P]),
    "rtsds_gate_resize_nhwc": (_I, [_P, _I, _I, _I, _I, _I, _P, _F, _I, _I, _P, _I, _I, _I, _P]),
    "rtsds_optim_job_blocks": (_I, [C.POINTER(OptJob)]),
    "rtsds_optim_step": (_I, [_P, _P, _I, _I, C.POINTER(OptHyper), _I, _P]),
    "rtsds_scale_packed_channels": (_I, [_P, _I, _L, _I, _I, _I, _F, _P]),
    "rtsds_ffm_head": (_I, [_P, _I, _I, _P, _I, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P]),
    "rtsds_resize_to_nchw": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_resize_ce_argmax_fwd": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _L, _P, _P, _P]),
    "rtsds_resize_ce_bwd": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _L, _P, _P, _P]),
    "rtsds_ce_argmax_nchw_fwd": (_I, [_P, _I, _I, _L, _P, _L, _P, _P, _P]),
    "rtsds_resize_ce_fused_supported": (_I, [_I, _I, _I, _I, _I]),
    "rtsds_resize_ce_fused": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _L, _P, _P, _P, _P]),
    "rtsds_scale_by_device_scalar": (_I, [_P, _L, _P, _P]),
    "rtsds_pack_conv_weight_cpad": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_unpack_conv_wgrad_cpad": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_s2d_out_size": (_I, [_I]),
    "rtsds_s2d_fwd": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_s2d_bwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "rtsds_s2d_weight": (_I, [_P, _I, _I, _P, _P]),
    "rtsds_s2d_weight_grad": (_I, [_P, _I, _I, _P, _P]),
    "rtsds_act_bwd": (_I, [_P, _I, _P, _I, _L, _I, _I, _F, _I, _P, _I, _P, _P]),
    "rtsds_disc_cls_fwd": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "rtsds_disc_cls_bwd": (_I, [_P, _F, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _F, _P, _I, _P, _P, _P, _P]),
    "rtsds_bce_logits": (_I, [_P, _I, _F, _F, _P, _P, _P]),
}

_lib = None


class RtsdsError(RuntimeError):
    pass


def build(verbose: bool = False, force: bool = False):
    from .build import build as _build

    return _build(verbose=verbose, force=force)


class _DryLib:
    """RTSDS_DRYRUN=1: record the launch sequence without touching a GPU.  It computes
    NOTHING (outputs stay uninitialised) — it exists so the host-side planning logic can
    be unit-tested on a CPU-only box, and is never selected implicitly."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        if name not in SIGNATURES:
            raise AttributeError(name)

        def fn(*args):
            self.calls.append(name)
            if name == "rtsds_conv_cout_pad":
                c = args[0]
                return 32 if c <= 32 else 64 if c <= 64 else (c + 127) // 128 * 128
            if name == "rtsds_conv2d_tc_workspace_bytes":
                return 1024
            if name == "rtsds_last_error_string":
                return b"dry run"
            if name == "rtsds_launch_count":
                return len(self.calls)
            return ABI_VERSION if name == "rtsds_abi_version" else 0

        return fn


def dry_run() -> bool:
    return os.environ.get("RTSDS_DRYRUN") == "1"


def lib():
    """Load (once) and return the shared library; raise loudly if unavailable."""
    global _lib
    if dry_run():
        if not isinstance(_lib, _DryLib):
            _lib = _DryLib()
        return _lib
    if isinstance(_lib, _DryLib):
        _lib = None
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if os.environ.get("RTSDS_NO_AUTOBUILD"):
            raise RtsdsError(f"{LIB_PATH} is missing; run `python -m rtsds_b200.build` (no CPU fallback exists)")
        build()                      # file-locked and linked atomically (rtsds_b200/build.py): safe under torchrun
    elif not os.environ.get("RTSDS_NO_AUTOBUILD"):
        # stale library: csrc/ or include/ changed since it was linked -> rebuild (needs nvcc; otherwise keep what is there)
        try:
            from .build import BUILD, source_digest

            stamp = BUILD / "lib.digest"
            if stamp.exists() and stamp.read_text() != source_digest():
                build()
        except Exception:  # noqa: BLE001 - no nvcc / read-only tree: the ABI version check below still guards the load
            pass
    try:
        handle = C.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover
        raise RtsdsError(f"cannot load {LIB_PATH}: {e} (no CPU fallback exists)") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as e:
            raise RtsdsError(f"{LIB_PATH} does not export {name}; rebuild with `python -m rtsds_b200.build --force`") from e
        fn.restype = res
        fn.argtypes = args
    if handle.rtsds_abi_version() != ABI_VERSION:
        raise RtsdsError("librtsds_b200.so ABI version mismatch")
    _lib = handle
    return handle


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().rtsds_last_error_string().decode(errors="replace")
        raise RtsdsError(f"{what or 'rtsds'} failed (code {rc}): {msg}")
