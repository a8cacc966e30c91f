"""ctypes binding of libed2b200.so (the C ABI in include/ed2b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
PyTorch is used only for device memory (`tensor.data_ptr()`) and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libed2b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_COS = 0, 1, 2
MAJOR_K, MAJOR_MN = 0, 1

_vp, _i, _f, _ll, _u64 = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_uint64


class Noise(C.Structure):
    _fields_ = [("injected", _vp), ("seed", _u64), ("stream", _u64), ("step", _vp),
                ("row0", _ll), ("rows_per_group_local", _i), ("rows_per_group_global", _i)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("dtype", _i),
        ("a", _vp), ("lda", _i), ("major_a", _i),
        ("b", _vp), ("ldb", _i), ("major_b", _i),
        ("m", _i), ("n", _i), ("k", _i),
        ("alpha", _f), ("beta", _f), ("act_scale", _f),
        ("act", _i), ("group_rows", _i),
        ("elem_scale", _vp), ("elem_scale_ld", _i), ("elem_scale_dtype", _i),
        ("col_scale", _vp), ("col_scale_ld", _i),
        ("addend", _vp), ("addend_ld", _i), ("addend_dtype", _i),
        ("col_bias", _vp), ("col_bias_ld", _i),
        ("out", _vp), ("out_ld", _i), ("out_dtype", _i),
        ("raw", _vp), ("raw_ld", _i), ("raw_dtype", _i),
        ("row_sum", _vp),
        ("mask_src", _vp), ("mask_ld", _i), ("mask_dtype", _i),
        ("col_sum", _vp), ("col_sum_ld", _i),
        ("block_n", _i), ("cta_group", _i), ("max_ctas", _i), ("no_tma_store", _i),
        ("accumulate", _i), ("lower_only", _i),
    ]


class BeFwd(C.Structure):
    _fields_ = [
        ("dtype", _i), ("batch", _i), ("in_dim", _i), ("out_dim", _i), ("ensemble_size", _i), ("act", _i),
        ("x", _vp), ("x_ld", _i), ("w_lp", _vp), ("w_ld", _i),
        ("alpha", _vp), ("gamma", _vp), ("bias", _vp),
        ("y", _vp), ("y_ld", _i), ("xa", _vp), ("xa_ld", _i), ("z", _vp), ("z_ld", _i),
    ]


class BeBwd(C.Structure):
    _fields_ = [
        ("dtype", _i), ("batch", _i), ("in_dim", _i), ("out_dim", _i), ("ensemble_size", _i), ("act", _i),
        ("gy", _vp), ("gy_ld", _i), ("y", _vp), ("y_ld", _i), ("z", _vp), ("z_ld", _i),
        ("x", _vp), ("x_ld", _i), ("xa", _vp), ("xa_ld", _i), ("w_lp", _vp), ("w_ld", _i),
        ("alpha", _vp), ("gamma", _vp),
        ("dz", _vp), ("dz_ld", _i), ("dxa", _vp), ("dxa_ld", _i), ("dx", _vp), ("dx_ld", _i),
        ("dw", _vp), ("dalpha", _vp), ("dgamma", _vp), ("dbias", _vp),
    ]


class Rank1Fwd(C.Structure):
    _fields_ = [
        ("dtype", _i), ("batch", _i), ("in_dim", _i), ("out_dim", _i), ("ensemble_size", _i), ("act", _i),
        ("x", _vp), ("x_ld", _i), ("w_lp", _vp), ("w_ld", _i),
        ("mu_r", _vp), ("rho_r", _vp), ("eps_r", Noise),
        ("mu_s", _vp), ("rho_s", _vp), ("eps_s", Noise),
        ("bias", _vp),
        ("y", _vp), ("y_ld", _i), ("xr", _vp), ("xr_ld", _i), ("s_buf", _vp), ("s_ld", _i), ("z", _vp), ("z_ld", _i),
        ("additive", _i), ("clip_lo", _f), ("clip_hi", _f),
        ("rho_b", _vp), ("eps_b", Noise), ("bias_buf", _vp), ("bias_ld", _i),
        ("fast_ws", _vp), ("xr_ready", _i),
        ("next_mu_r", _vp), ("next_rho_r", _vp), ("next_eps_r", Noise), ("next_xr", _vp), ("next_xr_ld", _i),
        ("gemm_ready_event", _vp),
    ]


class Rank1Bwd(C.Structure):
    _fields_ = [
        ("dtype", _i), ("batch", _i), ("in_dim", _i), ("out_dim", _i), ("ensemble_size", _i), ("act", _i),
        ("gy", _vp), ("gy_ld", _i), ("y", _vp), ("y_ld", _i), ("z", _vp), ("z_ld", _i),
        ("x", _vp), ("x_ld", _i), ("xr", _vp), ("xr_ld", _i), ("s_buf", _vp), ("s_ld", _i),
        ("w_lp", _vp), ("w_ld", _i),
        ("mu_r", _vp), ("rho_r", _vp), ("eps_r", Noise),
        ("mu_s", _vp), ("rho_s", _vp), ("eps_s", Noise),
        ("dz", _vp), ("dz_ld", _i), ("dxr", _vp), ("dxr_ld", _i), ("dx", _vp), ("dx_ld", _i),
        ("dw", _vp), ("dmu_r", _vp), ("drho_r", _vp), ("dmu_s", _vp), ("drho_s", _vp), ("dbias", _vp),
        ("additive", _i), ("clip_lo", _f), ("clip_hi", _f),
        ("rho_b", _vp), ("eps_b", Noise), ("drho_b", _vp),
        ("fast_ws", _vp), ("dz_ready", _i),
        ("prev_z", _vp), ("prev_z_ld", _i), ("prev_mu_s", _vp), ("prev_rho_s", _vp), ("prev_eps_s", Noise),
        ("prev_act", _i), ("prev_dz", _vp), ("prev_dz_ld", _i),
        ("prev_dmu_s", _vp), ("prev_drho_s", _vp), ("prev_dbias", _vp),
        ("dw_lp", _vp), ("dw_done_event", _vp),
    ]


class ReparamFwd(C.Structure):
    _fields_ = [
        ("dtype", _i), ("batch", _i), ("in_dim", _i), ("out_dim", _i), ("act", _i),
        ("x", _vp), ("x_ld", _i),
        ("mu", _vp), ("rho", _vp), ("eps", Noise), ("bias", _vp),
        ("w_lp", _vp), ("w_ld", _i), ("y", _vp), ("y_ld", _i),
        ("kl_out", _vp), ("kl_scale", _f), ("prior_mean", _f), ("prior_std", _f),
    ]


class ReparamBwd(C.Structure):
    _fields_ = [
        ("dtype", _i), ("batch", _i), ("in_dim", _i), ("out_dim", _i), ("act", _i),
        ("gy", _vp), ("gy_ld", _i), ("y", _vp), ("y_ld", _i), ("x", _vp), ("x_ld", _i),
        ("w_lp", _vp), ("w_ld", _i),
        ("mu", _vp), ("rho", _vp), ("eps", Noise),
        ("gp", _vp), ("gp_ld", _i), ("dx", _vp), ("dx_ld", _i),
        ("dmu", _vp), ("drho", _vp), ("dbias", _vp),
        ("kl_grad_scale", _f), ("prior_mean", _f), ("prior_std", _f), ("kl_upstream", _vp),
        ("gy_premasked", _i), ("x_is_relu", _i), ("prev_dbias", _vp),
        ("phase", _i), ("dmu_lp", _vp), ("drho_lp", _vp),
    ]


MAX_PEERS = 8
IPC_HANDLE_BYTES = 64


class PeerExchange(C.Structure):
    _fields_ = [
        ("world", _i), ("rank", _i), ("n", _ll), ("chunk", _ll), ("small_n", _ll),
        ("bucket", _vp * MAX_PEERS), ("reduced", _vp * MAX_PEERS), ("small_vec", _vp * MAX_PEERS),
        ("flags", _vp * MAX_PEERS),
    ]


class FlipoutFwd(C.Structure):
    _fields_ = [
        ("dtype", _i), ("batch", _i), ("in_dim", _i), ("out_dim", _i), ("act", _i),
        ("x", _vp), ("x_ld", _i),
        ("mu", _vp), ("rho", _vp), ("eps", Noise),
        ("sign_in", Noise), ("sign_out", Noise), ("bias", _vp),
        ("mu_lp", _vp), ("mu_ld", _i), ("delta_lp", _vp), ("delta_ld", _i),
        ("xs", _vp), ("xs_ld", _i), ("pert", _vp), ("pert_ld", _i), ("y", _vp), ("y_ld", _i),
        ("kl_out", _vp), ("kl_scale", _f), ("prior_mean", _f), ("prior_std", _f),
    ]


class FlipoutBwd(C.Structure):
    _fields_ = [
        ("dtype", _i), ("batch", _i), ("in_dim", _i), ("out_dim", _i), ("act", _i),
        ("gy", _vp), ("gy_ld", _i), ("y", _vp), ("y_ld", _i), ("x", _vp), ("x_ld", _i), ("xs", _vp), ("xs_ld", _i),
        ("mu_lp", _vp), ("mu_ld", _i), ("delta_lp", _vp), ("delta_ld", _i),
        ("mu", _vp), ("rho", _vp), ("eps", Noise), ("sign_in", Noise), ("sign_out", Noise),
        ("gp", _vp), ("gp_ld", _i), ("gt", _vp), ("gt_ld", _i), ("tmp", _vp), ("tmp_ld", _i),
        ("dx", _vp), ("dx_ld", _i), ("dmu", _vp), ("drho", _vp), ("dbias", _vp),
        ("kl_grad_scale", _f), ("prior_mean", _f), ("prior_std", _f), ("kl_upstream", _vp),
    ]


class FlipoutFusedFwd(C.Structure):
    _fields_ = [
        ("batch", _i), ("in_dim", _i), ("out_dim", _i), ("act", _i),
        ("x", _vp), ("x_ld", _i), ("xs", _vp), ("xs_ld", _i), ("xs_ready", _i),
        ("sign_in", Noise), ("sign_out", Noise),
        ("mu_lp", _vp), ("mu_ld", _i), ("delta_lp", _vp), ("delta_ld", _i),
        ("bias", _vp), ("y", _vp), ("y_ld", _i),
        ("next_xs", _vp), ("next_xs_ld", _i), ("next_sign_in", Noise),
    ]


class FlipoutFusedBwd(C.Structure):
    _fields_ = [
        ("batch", _i), ("in_dim", _i), ("out_dim", _i), ("act", _i), ("premade", _i),
        ("gy", _vp), ("gy_ld", _i), ("y", _vp), ("y_ld", _i),
        ("gp", _vp), ("gp_ld", _i), ("gt", _vp), ("gt_ld", _i),
        ("x", _vp), ("x_ld", _i), ("xs", _vp), ("xs_ld", _i),
        ("mu_lp", _vp), ("mu_ld", _i), ("delta_lp", _vp), ("delta_ld", _i),
        ("mu", _vp), ("rho", _vp), ("eps", Noise), ("sign_in", Noise), ("sign_out", Noise),
        ("dmu", _vp), ("drho", _vp), ("dbias", _vp),
        ("kl_grad_scale", _f), ("prior_mean", _f), ("prior_std", _f), ("kl_upstream", _vp),
        ("dx", _vp), ("dx_ld", _i),
        ("prev_relu", _i), ("prev_gp", _vp), ("prev_gp_ld", _i), ("prev_gt", _vp), ("prev_gt_ld", _i),
        ("prev_sign_out", Noise), ("prev_dbias", _vp),
    ]


class ConvGeometry(C.Structure):
    _fields_ = [("batch", _i), ("height", _i), ("width", _i), ("channels", _i),
                ("kh", _i), ("kw", _i), ("sh", _i), ("sw", _i), ("same", _i)]


class ConvFactorFwd(C.Structure):
    _fields_ = [
        ("geom", ConvGeometry), ("dtype", _i), ("out_channels", _i), ("act", _i),
        ("x", _vp), ("w_lp", _vp), ("w_ld", _i), ("f_in", _vp), ("f_out", _vp), ("bias", _vp),
        ("xa", _vp), ("xa_ld", _i), ("y", _vp), ("y_ld", _i), ("z", _vp), ("z_ld", _i),
    ]


class ConvFactorBwd(C.Structure):
    _fields_ = [
        ("geom", ConvGeometry), ("dtype", _i), ("out_channels", _i), ("act", _i),
        ("gy", _vp), ("gy_ld", _i), ("y", _vp), ("y_ld", _i), ("z", _vp), ("z_ld", _i),
        ("x", _vp), ("xa", _vp), ("xa_ld", _i), ("w_lp", _vp), ("w_ld", _i),
        ("f_in", _vp), ("f_out", _vp),
        ("dz", _vp), ("dz_ld", _i), ("dxa", _vp), ("dxa_ld", _i), ("gimg", _vp), ("dx", _vp),
        ("dw", _vp), ("df_in", _vp), ("df_out", _vp), ("dbias", _vp),
    ]


# name -> (restype, argtypes); every symbol include/ed2b200.h declares
SIGNATURES = {
    "ed2_version": (C.c_char_p, []),
    "ed2_last_error": (C.c_char_p, []),
    "ed2_launch_count": (_ll, []),
    "ed2_check_device": (_i, [_i]),
    "ed2_gemm": (_i, [C.POINTER(GemmDesc), _vp]),
    "ed2_philox_normal": (_i, [_vp, _ll, _u64, _u64, _vp]),
    "ed2_philox_normal_fast": (_i, [_vp, _ll, _u64, _u64, _vp]),
    "ed2_philox_uniform": (_i, [_vp, _ll, _u64, _u64, _vp]),
    "ed2_philox_sign": (_i, [_vp, _ll, _u64, _u64, _vp]),
    "ed2_philox_raw": (_i, [_vp, _ll, _u64, _u64, _vp]),
    "ed2_cast": (_i, [_vp, _i, _i, _vp, _i, _i, _ll, _ll, _vp]),
    "ed2_split_bf16": (_i, [_vp, _i, _i, _vp, _i, _ll, _i, _i, _i, _vp]),
    "ed2_sample_normal_weights": (_i, [_vp, _vp, Noise, _ll, _ll, _i, _vp, _i, _i, _vp, _i, _vp, _f, _f, _f, _vp]),
    "ed2_normal_kl_fwd": (_i, [_vp, _vp, _ll, _f, _f, _f, _vp, _vp]),
    "ed2_normal_kl_bwd": (_i, [_vp, _vp, _ll, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "ed2_softmax_xent": (_i, [_vp, _i, _i, _vp, _ll, _i, _f, _f, _vp, _vp, _i, _i, _vp, _i, _vp]),
    "ed2_act_bwd": (_i, [_i, _vp, _i, _vp, _i, _ll, _i, _i, _vp, _i, _vp, _vp]),
    "ed2_be_dense_fwd": (_i, [C.POINTER(BeFwd), _vp]),
    "ed2_be_dense_bwd": (_i, [C.POINTER(BeBwd), _vp]),
    "ed2_rank1_dense_fwd": (_i, [C.POINTER(Rank1Fwd), _vp]),
    "ed2_rank1_dense_bwd": (_i, [C.POINTER(Rank1Bwd), _vp]),
    "ed2_rank1_fusable": (_i, [_i, _i, _i]),
    "ed2_rank1_workspace_floats": (_ll, [_i, _i, _i]),
    "ed2_reparam_dense_fwd": (_i, [C.POINTER(ReparamFwd), _vp]),
    "ed2_reparam_dense_bwd": (_i, [C.POINTER(ReparamBwd), _vp]),
    "ed2_normal_grad_finish": (_i, [_vp, _i, _vp, _vp, Noise, _ll, _f, _vp, _f, _f, _vp, _vp, _i, _vp]),
    "ed2_peer_alloc": (_i, [C.c_size_t, C.POINTER(_vp), _vp]),
    "ed2_peer_open": (_i, [_vp, C.POINTER(_vp)]),
    "ed2_peer_close": (_i, [_vp]),
    "ed2_peer_free": (_i, [_vp]),
    "ed2_peer_signal": (_i, [C.POINTER(PeerExchange), _vp, _vp]),
    "ed2_peer_flag": (_i, [C.POINTER(PeerExchange), _i, _vp]),
    "ed2_peer_wait": (_i, [C.POINTER(PeerExchange), _i, _vp]),
    "ed2_dma_copy": (_i, [_vp, _vp, C.c_size_t, _vp]),
    "ed2_peer_reduce_slice": (_i, [C.POINTER(PeerExchange), _i, _vp]),
    "ed2_peer_gather_f32": (_i, [C.POINTER(PeerExchange), _vp, _vp]),
    "ed2_peer_grad_finish": (_i, [C.POINTER(PeerExchange), _vp, _vp, Noise, _f, _vp, _f, _f, _vp, _vp, _vp, _i, _i, _vp]),
    "ed2_flipout_dense_fwd": (_i, [C.POINTER(FlipoutFwd), _vp]),
    "ed2_flipout_dense_bwd": (_i, [C.POINTER(FlipoutBwd), _vp]),
    "ed2_flipout_fused_fwd": (_i, [C.POINTER(FlipoutFusedFwd), _vp]),
    "ed2_flipout_fused_bwd": (_i, [C.POINTER(FlipoutFusedBwd), _vp]),
    "ed2_flipout_fusable": (_i, [_i, _i, _i]),
    "ed2_mc_softmax_fa_fwd": (_i, [_vp, _vp, _vp, _vp, _f, _f, Noise, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp]),
    "ed2_mc_softmax_fa_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, Noise, _i, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp, _vp]),
    "ed2_lstm_pointwise_fwd": (_i, [_vp, _i, _ll, _vp, _i, _i, _vp, _vp, _ll, _vp]),
    "ed2_lstm_pointwise_bwd": (_i, [_vp, _i, _ll, _vp, _vp, _vp, _ll, _vp, _i, _i, _vp, _ll, _vp, _vp, _vp]),
    "ed2_lstm_gates_fwd": (_i, [_vp, _vp, _i, _ll, _ll, _vp, _i, _i, _vp, _vp, _ll, _vp]),
    "ed2_lstm_gates_bwd": (_i, [_vp, _vp, _i, _ll, _ll, _vp, _vp, _vp, _ll, _vp, _i, _i, _vp, _ll, _vp, _vp]),
    "ed2_spd_inverse_workspace_floats": (_ll, [_i]),
    "ed2_spd_inverse": (_i, [_vp, _i, _vp, _vp]),
    "ed2_conv_out_size": (_i, [C.POINTER(ConvGeometry), C.POINTER(_i), C.POINTER(_i)]),
    "ed2_im2col": (_i, [C.POINTER(ConvGeometry), _vp, _i, _vp, _ll, _vp]),
    "ed2_col2im": (_i, [C.POINTER(ConvGeometry), _vp, _i, _ll, _vp, _i, _vp]),
    "ed2_conv_factor_fwd": (_i, [C.POINTER(ConvFactorFwd), _vp]),
    "ed2_conv_factor_bwd": (_i, [C.POINTER(ConvFactorBwd), _vp]),
    "ed2_expand_table": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ed2_expand_table_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ed2_rank1_factors": (_i, [_vp, _vp, Noise, _i, _i, _i, _f, _f, _vp, _vp]),
    "ed2_rank1_factors_bwd": (_i, [_vp, _vp, _vp, Noise, _i, _i, _i, _f, _f, _vp, _vp, _vp]),
    "ed2_noise_table": (_i, [_i, Noise, _vp, _i, _ll, _ll, _ll, _vp]),
    "ed2_normal_grad_finish2": (_i, [_vp, _vp, _vp, _vp, Noise, _ll, _f, _vp, _f, _f, _vp, _vp, _i, _vp]),
    "ed2_spectral_norm_update": (_i, [_vp, _i, _i, _vp, _vp, _i, _f, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "ed2_input_normalize": (_i, [_vp, _i, _i, _vp, _i, _i, _ll, _i, _vp, _vp, _f, _i, _f, _vp]),
    "ed2_layernorm_bwd": (_i, [_vp, _i, _i, _vp, _i, _i, _ll, _i, _vp, _f, _vp, _i, _i, _vp, _vp, _vp]),
    "ed2_rff_fwd": (_i, [_i, _vp, _i, _vp, _i, _vp, _f, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp]),
    "ed2_rff_bwd": (_i, [_i, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _i, _f, _i, _i, _i, _vp, _i, _vp, _i, _i, _vp]),
    "ed2_layernorm_split": (_i, [_vp, _i, _i, _vp, _i, _ll, _i, _vp, _vp, _f, _vp]),
    "ed2_laplace_precision_update": (_i, [_i, _vp, _i, _i, _i, _f, _ll, _vp, _vp, _vp]),
    "ed2_precision_blend": (_i, [_vp, _vp, _i, _f, _ll, _vp]),
    "ed2_predictive_variance": (_i, [_i, _vp, _i, _vp, _i, _i, _i, _f, _vp, _vp, _i, _f, _i, _vp, _vp]),
    "ed2_mean_field_logits": (_i, [_vp, _vp, _i, _i, _f, _i, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load libed2b200.so (building it in-tree first if the sources are newer and nvcc is present)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            from . import _build

            _build.build()
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: the CUDA extension is required (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here means the header and the library diverged
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    return load().ed2_last_error().decode()


def check(status: int, what: str = "") -> None:
    if status != 0:
        raise RuntimeError(f"libed2b200 {what} failed with status {status}: {last_error()}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dt_of(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}; libed2b200 takes float32 or bfloat16")


def torch_dtype(dt: int) -> torch.dtype:
    return torch.float32 if dt == F32 else torch.bfloat16


def ptr(t) -> int | None:
    if t is None:
        return None
    assert t.is_cuda, "libed2b200 takes device pointers only (no CPU path)"
    return t.data_ptr()


def ld(t: torch.Tensor) -> int:
    """Leading dimension (elements) of a 2-D row-major tensor."""
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D tensor"
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def noise(injected: torch.Tensor | None = None, seed: int = 0, stream: int = 0, step: torch.Tensor | None = None,
          rows: tuple[int, int, int] | None = None) -> Noise:
    """`rows` = (row0, rows_per_group_local, rows_per_group_global): the data-parallel row mapping of a per-example
    noise tensor (ed2_noise in include/ed2b200.h); None = identity."""
    n = Noise()
    n.row0, n.rows_per_group_local, n.rows_per_group_global = rows if rows is not None else (0, 0, 0)
    if injected is not None:
        assert injected.dtype == torch.float32 and injected.is_contiguous() and injected.is_cuda
        n.injected = injected.data_ptr()
    else:
        n.injected = None
    n.seed = seed & 0xFFFFFFFFFFFFFFFF
    n.stream = stream & 0xFFFFFFFFFFFFFFFF
    n.step = None if step is None else step.data_ptr()
    return n


def launch_count() -> int:
    return int(load().ed2_launch_count())


def pad8(n: int) -> int:
    """Row pitch (elements) that keeps 16-byte alignment for bf16 rows (TMA requirement)."""
    return (n + 7) // 8 * 8


def empty_padded(rows: int, cols: int, dtype: torch.dtype, device) -> torch.Tensor:
    """[rows, cols] view over a buffer whose pitch is a multiple of 8 elements."""
    p = pad8(cols)
    buf = torch.empty((rows, p), dtype=dtype, device=device)
    return buf[:, :cols] if p != cols else buf
