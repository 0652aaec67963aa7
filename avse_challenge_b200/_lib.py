"""ctypes binding of ``libmtn_b200.so`` (the C ABI declared in ``include/mtn_b200.h``).

There is no fallback: if the library is missing, or an entry point fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmtn_b200.so")

EPI_STORE, EPI_INPROJ, EPI_MASK, EPI_RELU, EPI_XPROJ, EPI_RESADD = 0, 1, 2, 3, 4, 5

# ABI the ctypes structs below were written for (include/mtn_b200.h: MTN_ABI_VERSION).  A stale or experiment-only build of
# the library (it is git-ignored and rebuilt on mtimes) must not be handed structs of another layout.
EXPECTED_ABI = 7

EXPORTS = [
    "mtn_encoder_cln_fwd", "mtn_gemm_fwd", "mtn_gemm_rowsum_parts", "mtn_add_rmsnorm_fwd", "mtn_add_rmsnorm_out_fwd", "mtn_add_norm_fwd", "mtn_conv_silu_fwd", "mtn_conv_silu_halo_fwd", "mtn_conv_silu_dir_fwd", "mtn_decoder_stream_fwd",
    "mtn_scan_fwd", "mtn_fold_states_fwd", "mtn_fold_states_packed_fwd",
    "mtn_gn_partials_bytes", "mtn_gn_stats_fwd", "mtn_gn_apply_fwd", "mtn_gn_apply_norm_fwd", "mtn_dp_num_chunks", "mtn_dp_segment_fwd",
    "mtn_dp_overadd_prelu_fwd", "mtn_bias_planes_fwd", "mtn_gate_planes_fwd",
    "mtn_decoder_fwd", "mtn_cln_fwd", "mtn_softmax_mask_fwd", "mtn_split_planes", "mtn_si_snr_pit_fwd", "mtn_si_snr_workspace_bytes", "mtn_si_snr_pit_n_fwd", "mtn_si_snr_workspace_bytes_n", "mtn_last_error_string", "mtn_abi_version",
    "mtn_sizeof_gemm_args", "mtn_sizeof_scan_args", "mtn_sizeof_gn_apply_args",
    "mtn_stream_push_fwd", "mtn_sizeof_stream_push_args", "mtn_stream_push_smem_bytes", "mtn_conv_xproj_fwd",
]


class GemmArgs(Structure):
    _fields_ = [
        ("a", c_void_p), ("w", c_void_p), ("out", c_void_p), ("aux", c_void_p),
        ("M", c_int), ("N", c_int), ("K", c_int), ("a_rows", c_int),
        ("lda", c_int), ("ldo", c_int), ("ld_aux", c_int),
        ("planes", c_int), ("groups", c_int), ("out_group_stride", c_int),
        ("epilogue", c_int), ("epi_param", c_int), ("out_bf16", c_int), ("max_ctas", c_int),
        ("out2", c_void_p), ("rowsum", c_void_p), ("ldo2", c_int), ("a2_rows", c_int),
        ("rowsq", c_void_p), ("rowsq_scale", c_float), ("rowsq_eps", c_float), ("rowsq_parts", c_int),
    ]


class ScanArgs(Structure):
    _fields_ = [
        ("u", c_void_p), ("dbl", c_void_p), ("z", c_void_p), ("w_dt", c_void_p), ("dt_bias", c_void_p),
        ("A2", c_void_p), ("Dskip", c_void_p), ("y", c_void_p), ("h_in", c_void_p), ("h_out", c_void_p),
        ("batch", c_int), ("L", c_int), ("di", c_int), ("R", c_int), ("n_dbl", c_int), ("ld_dbl", c_int),
        ("ldz", c_int), ("z_col0", c_int), ("planes", c_int), ("z_bf16", c_int), ("dir_mask", c_int),
        ("sum_delta", c_void_p), ("L_last", c_int), ("dtp", c_void_p),
    ]


class GnApplyArgs(Structure):
    _fields_ = [
        ("x", c_void_p), ("partials", c_void_p), ("w", c_void_p), ("bias", c_void_p), ("skip", c_void_p),
        ("out_a", c_void_p), ("out_a2", c_void_p), ("out_t", c_void_p), ("planes", c_void_p),
        ("batch", c_int), ("S", c_int), ("K", c_int), ("C", c_int), ("x_transposed", c_int),
        ("n_planes", c_int), ("plane_rows", c_int), ("eps", c_float), ("blend", c_void_p),
    ]


class StreamPushArgs(Structure):
    _fields_ = [
        ("mix", c_void_p), ("in_tail", c_void_p), ("est", c_void_p), ("halo", c_void_p), ("h", c_void_p), ("ola_tail", c_void_p),
        ("head", c_void_p), ("bot_frag", c_void_p), ("mask_frag", c_void_p), ("layer_vec", c_void_p),
        ("layer_frag", c_void_p),
        ("h_layer_stride", c_size_t), ("layer_vec_stride", c_size_t), ("layer_frag_stride", c_size_t),
        ("B", c_int), ("F", c_int), ("N", c_int), ("D", c_int), ("di", c_int), ("R", c_int), ("n_spk", c_int),
        ("n_layers", c_int), ("ld_mix", c_int), ("first", c_int), ("eps_cln", c_float), ("eps_rms", c_float),
        ("timeline", c_void_p), ("halo_stream_stride", c_size_t), ("halo_layer_stride", c_size_t),
        ("halo_rows", c_int), ("stack_x", c_void_p), ("stack_out", c_void_p), ("dsl", c_int),
    ]


class MtnError(RuntimeError):
    pass


_lib = None


def load():
    """Load the CUDA library; raise loudly when it has not been built (no CPU / eager fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MtnError(
            f"{LIB_PATH} not found: build it with `python -m avse_challenge_b200.build` "
            "(nvcc, sm_100a). There is no fallback implementation.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.mtn_last_error_string.restype = c_char_p
    lib.mtn_last_error_string.argtypes = []
    lib.mtn_abi_version.restype = c_int
    lib.mtn_abi_version.argtypes = []
    abi = int(lib.mtn_abi_version())
    if abi != EXPECTED_ABI:
        raise MtnError(f"{LIB_PATH} reports ABI {abi}, this binding was written for ABI {EXPECTED_ABI}: rebuild it with "
                       "`python -m avse_challenge_b200.build --force` (dev / experiment builds carry ABI + 1000 and load "
                       "only through tools/ with MTN_LIB set)")
    lib.mtn_encoder_cln_fwd.argtypes = [c_void_p, c_int] + [c_void_p] * 5 + [c_int] * 5 + [c_float, c_void_p]
    lib.mtn_gemm_fwd.argtypes = [POINTER(GemmArgs), c_void_p]
    lib.mtn_gemm_rowsum_parts.argtypes = [c_int]
    lib.mtn_add_rmsnorm_fwd.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                        c_void_p]
    lib.mtn_add_rmsnorm_out_fwd.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                            c_float, c_void_p]
    lib.mtn_add_norm_fwd.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_float, c_void_p]
    lib.mtn_conv_silu_fwd.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_void_p]
    lib.mtn_conv_silu_halo_fwd.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                           c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.mtn_conv_silu_dir_fwd.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                          c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.mtn_conv_xproj_fwd.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                       c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.mtn_decoder_stream_fwd.argtypes = [c_void_p] * 5 + [c_int] * 5 + [c_void_p]
    lib.mtn_scan_fwd.argtypes = [POINTER(ScanArgs), c_void_p]
    lib.mtn_fold_states_fwd.argtypes = [c_void_p] * 6 + [c_int] * 5 + [c_void_p]
    lib.mtn_fold_states_packed_fwd.argtypes = [c_void_p] * 5 + [c_int] * 6 + [c_void_p]
    lib.mtn_decoder_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.mtn_cln_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]
    lib.mtn_softmax_mask_fwd.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.mtn_split_planes.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.mtn_si_snr_workspace_bytes.argtypes = [c_int, c_int]
    lib.mtn_si_snr_pit_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]
    lib.mtn_si_snr_workspace_bytes_n.argtypes = [c_int, c_int, c_int]
    lib.mtn_si_snr_pit_n_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p,
                                         c_int, c_void_p]
    lib.mtn_gn_partials_bytes.argtypes = [c_int, c_int, c_int]
    lib.mtn_gn_stats_fwd.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    lib.mtn_gn_apply_fwd.argtypes = [POINTER(GnApplyArgs), c_void_p]
    lib.mtn_gn_apply_norm_fwd.argtypes = [POINTER(GnApplyArgs), c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p]
    lib.mtn_dp_num_chunks.argtypes = [c_int, c_int]
    lib.mtn_dp_segment_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.mtn_dp_overadd_prelu_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                             c_void_p]
    lib.mtn_bias_planes_fwd.argtypes = [c_void_p, c_int, c_void_p, c_float, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.mtn_gate_planes_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.mtn_stream_push_fwd.argtypes = [POINTER(StreamPushArgs), c_void_p]
    lib.mtn_stream_push_smem_bytes.argtypes = [c_int, c_int, c_int]
    lib.mtn_stream_push_smem_bytes.restype = c_size_t
    for name, st in (("mtn_sizeof_gemm_args", GemmArgs), ("mtn_sizeof_scan_args", ScanArgs),
                     ("mtn_sizeof_gn_apply_args", GnApplyArgs), ("mtn_sizeof_stream_push_args", StreamPushArgs)):
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = c_size_t, []
        if int(fn()) != ctypes.sizeof(st):
            raise MtnError(f"{name}() = {int(fn())} but ctypes.sizeof({st.__name__}) = {ctypes.sizeof(st)}: the binding "
                           "and include/mtn_b200.h disagree about the struct layout")
    for name in EXPORTS:
        fn = getattr(lib, name, None)
        if fn is not None and name not in ("mtn_last_error_string", "mtn_abi_version", "mtn_si_snr_workspace_bytes",
                                           "mtn_si_snr_workspace_bytes_n",
                                           "mtn_gn_partials_bytes", "mtn_sizeof_gemm_args", "mtn_sizeof_scan_args",
                                           "mtn_sizeof_gn_apply_args", "mtn_sizeof_stream_push_args",
                                           "mtn_stream_push_smem_bytes"):
            fn.restype = c_int
    lib.mtn_si_snr_workspace_bytes.restype = c_size_t
    lib.mtn_si_snr_workspace_bytes_n.restype = c_size_t
    lib.mtn_gn_partials_bytes.restype = c_size_t
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().mtn_last_error_string().decode("utf-8", "replace")
        raise MtnError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()
