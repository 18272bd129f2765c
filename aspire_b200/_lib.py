"""ctypes binding of ``libaspire_b200.so`` (the C-ABI in ``include/aspire_b200.h``).

There is no CPU fallback: if the shared library is missing, or the process has
no sm_100 device, every op raises :class:`AspireB200Error`.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASPIRE_B200_LIB selects another build of the same library (kernel A/B experiments)
LIB_PATH = os.environ.get("ASPIRE_B200_LIB") or os.path.join(_HERE, "lib", "libaspire_b200.so")

ASP_OK = 0
ASP_ERR_ARG = -1
ASP_ERR_DEVICE = -2
ASP_ERR_CUDA = -3

DENSITY_GAUSSIAN = 0
DENSITY_ROSENBROCK = 1
DENSITY_GMM = 2
DENSITY_UNIFORM = 3

CDF_TILE = 2048


class AspireB200Error(RuntimeError):
    """Raised for every failure of the native library (no silent fallback)."""


class Density(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("dims", C.c_int32),
        ("n_comp", C.c_int32),
        ("reserved", C.c_int32),
        ("p0", C.c_void_p),
        ("p1", C.c_void_p),
        ("p2", C.c_void_p),
        ("s0", C.c_double),
        ("s1", C.c_double),
        ("s2", C.c_double),
    ]


DEFAULT_DEVICE = "cuda"  # where objects live when the caller names no device (the kernels run nowhere else)

MAX_PEERS = 8
PEER_BOX_BYTES = 2 * MAX_PEERS * 16 * 8 + 64 + 128


class PeerDesc(C.Structure):
    """``asp_peer_t``: the ranks' mailboxes as addressable from this device."""
    _fields_ = [("box", C.c_void_p * MAX_PEERS), ("rank", C.c_int32), ("world", C.c_int32)]


class TransformDesc(C.Structure):
    """``asp_transform_t`` (include/aspire_b200.h)."""
    _fields_ = [("dims", C.c_int32), ("bounded", C.c_int32), ("affine", C.c_int32), ("reserved", C.c_int32),
                ("eps", C.c_double), ("d_kind", C.c_void_p), ("d_lower", C.c_void_p), ("d_upper", C.c_void_p),
                ("d_mean", C.c_void_p), ("d_std", C.c_void_p)]


BOUNDED_NONE, BOUNDED_LOGIT, BOUNDED_PROBIT = 0, 1, 2


class BetaState(C.Structure):
    _fields_ = [
        ("beta_prev", C.c_double),
        ("beta_min", C.c_double),
        ("beta_max", C.c_double),
        ("target_eff", C.c_double),
        ("tol", C.c_double),
        ("n_total", C.c_double),
        ("beta_try", C.c_double),
        ("s1", C.c_double),
        ("s2", C.c_double),
        ("eff_beta_max", C.c_double),
        ("s1_min", C.c_double),
        ("s2_min", C.c_double),
        ("d_min", C.c_double),
        ("d_max", C.c_double),
        ("phase", C.c_int32),
        ("n_evals", C.c_int32),
        ("min_valid", C.c_int32),
        ("reserved", C.c_int32),
    ]


class PcnParams(C.Structure):
    _fields_ = [
        ("beta", C.c_double),
        ("rho", C.c_double),
        ("d_mu", C.c_void_p),
        ("d_sigma", C.c_void_p),
        ("n_moves", C.c_int32),
        ("noise_mode", C.c_int32),
        ("seed", C.c_uint64),
        ("stream0", C.c_uint32),
        ("reserved", C.c_uint32),
        ("index0", C.c_int64),
        ("d_xi", C.c_void_p),
        ("d_log_u", C.c_void_p),
        ("d_slot", C.c_void_p),
        ("nu", C.c_double),
        ("d_scale", C.c_void_p),
        ("d_gamma", C.c_void_p),
    ]


class Flow(C.Structure):
    _fields_ = [
        ("dims", C.c_int32),
        ("hidden", C.c_int32),
        ("n_layers", C.c_int32),
        ("reserved", C.c_int32),
        ("d_weights", C.c_void_p),
        ("d_bias", C.c_void_p),
        ("d_mean", C.c_void_p),
        ("d_inv_std", C.c_void_p),
        ("d_std", C.c_void_p),
        ("log_det0", C.c_double),
        ("d_error", C.c_void_p),
    ]


class FlowTrain(C.Structure):
    """``asp_flow_train_t`` (include/aspire_b200.h)."""
    _fields_ = [("kdims", C.c_int32), ("hidden", C.c_int32), ("n_layers", C.c_int32), ("batch_cap", C.c_int32),
                ("d_w", C.c_void_p), ("d_m", C.c_void_p), ("d_v", C.c_void_p), ("d_g", C.c_void_p),
                ("d_out_mask", C.c_void_p), ("d_mean", C.c_void_p), ("d_inv_std", C.c_void_p), ("d_ws", C.c_void_p)]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_D = C.c_double

# name -> (restype, argtypes); every symbol include/aspire_b200.h declares
SIGNATURES = {
    "asp_version": (C.c_int, []),
    "asp_last_error": (C.c_char_p, []),
    "asp_device_check": (C.c_int, []),
    "asp_logdensity": (C.c_int, [C.POINTER(Density), _P, _I64, _P, _P]),
    "asp_importance_workspace_bytes": (C.c_size_t, [_I64]),
    "asp_importance_weights": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P, _P]),
    "asp_reduce_workspace_bytes": (C.c_size_t, []),
    "asp_reweight": (C.c_int, [_P, _P, _P, _I64, _P, _P, _P, _P]),
    "asp_ess_sums": (C.c_int, [_P, _I64, _P, C.c_int, _P, _P, _P, _P]),
    "asp_beta_search_max_passes": (C.c_int, [_D, _D]),
    "asp_beta_search_init": (C.c_int, [_P, _D, _D, _D, _D, _P, _P]),
    "asp_beta_search_pass": (C.c_int, [_P, _I64, _P, C.c_int, _P, _P]),
    "asp_beta_search_decide": (C.c_int, [_P, _P]),
    "asp_beta_search_max_passes3": (C.c_int, [_D, _D]),
    "asp_beta_search_pass3": (C.c_int, [_P, _I64, _P, _P, C.c_int, _P, _P]),
    "asp_beta_search_decide3": (C.c_int, [_P, _P, _P]),
    "asp_beta_search_scratch_words": (_I32, []),
    "asp_beta_search_run": (C.c_int, [_P, _I64, _P, _D, _D, _D, _D, _P, _P, _P, _P, _P]),
    "asp_beta_search_root": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P]),
    "asp_beta_search_pass_p2p": (C.c_int, [_P, _I64, _P, C.c_int, _P, _P, _P, _P]),
    "asp_smc_weights": (C.c_int, [_P, _I64, _D, _D, _D, _D, _P, _P, _P, _P, _P]),
    "asp_cdf_workspace_bytes": (C.c_size_t, [_I64]),
    "asp_cdf_tile_stats": (C.c_int, [_P, _I64, _P, _P]),
    "asp_cdf_set_tile_stats": (C.c_int, [_P, _I64, _P, _P]),
    "asp_cdf_plan": (C.c_int, [_I64, _P, _P, _P]),
    "asp_cdf_aggregate": (C.c_int, [_P, _I64, _P, _P]),
    "asp_cdf_resolve": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P]),
    "asp_cdf_write": (C.c_int, [_P, _I64, _P, C.c_int, _P, _P, _P]),
    "asp_transform_forward": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "asp_transform_inverse": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "asp_cdf_resolve_p2p": (C.c_int, [_P, _I64, _P, _P, _P, _P, _P]),
    "asp_cdf_write_p2p": (C.c_int, [_P, _I64, C.c_int, _P, _P, _P, _P, _P]),
    "asp_cdf_exact": (C.c_int, [_P, _I64, C.c_int, _P, _P, _P]),
    "asp_cdf_info": (C.c_int, [_P, C.POINTER(C.c_int64), _P]),
    "asp_search_multinomial": (C.c_int, [_P, _I64, _P, _I64, _P, _P]),
    "asp_search_workspace_bytes": (C.c_size_t, [_I64]),
    "asp_search_multinomial_guided": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _P]),
    "asp_search_systematic_workspace_bytes": (C.c_size_t, [_I64]),
    "asp_search_systematic": (C.c_int, [_P, _I64, _D, _I64, _I64, _I64, _P, _P, _P]),
    "asp_sorted_tiles": (_I64, [_I64]),
    "asp_sorted_spacings": (C.c_int, [C.c_uint64, C.c_uint32, _I64, _I64, _I64, _P, _P, _P]),
    "asp_sorted_tile_scan": (C.c_int, [_P, _I64, _P, _P, _P]),
    "asp_search_sorted_workspace_bytes": (C.c_size_t, [_I64]),
    "asp_search_sorted": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _I64, _I64, _P, _P, _P]),
    "asp_sorted_lower_bound": (C.c_int, [C.c_uint64, C.c_uint32, _P, _I64, _I64, _P, _I32, _P, _P]),
    "asp_gather_rows": (C.c_int, [_P, _I32, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "asp_moments_workspace_bytes": (C.c_size_t, [_I32]),
    "asp_weighted_moments": (C.c_int, [_P, _I32, _P, _I64, _P, _I64, _D, _D, _P, _P, _P]),
    "asp_moments_finish": (C.c_int, [_P, _I32, _D, _P, _P, _P]),
    "asp_pcn_sweep": (
        C.c_int,
        [_P, _P, _P, _P, _P, _I64, _I32, C.POINTER(Density), C.POINTER(Density), C.POINTER(Density),
         C.POINTER(PcnParams), _P, _P, _P, _P, _P, _P],
    ),
    "asp_pcn_propose": (C.c_int, [_P, _I64, _I32, C.POINTER(PcnParams), _I32, _P, _P]),
    "asp_pcn_accept": (
        C.c_int,
        [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, C.POINTER(PcnParams), _I32, _P, _P],
    ),
    "asp_tpcn_scales": (C.c_int, [_P, _I64, _I32, C.POINTER(PcnParams), _I32, _P, _P]),
    "asp_pcn_propose_eval_supported": (_I32, [_I32, _I32, _I32]),
    "asp_pcn_sweep_tpcn_supported": (_I32, [_I32]),
    "asp_pcn_propose_eval": (C.c_int, [_P, _I64, _I32, C.POINTER(Density), C.POINTER(Density), C.POINTER(PcnParams), _I32,
                                       _P, _P, _P, _P, _P]),
    "asp_pcn_accept_scalar": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, C.POINTER(PcnParams), _I32, _P,
                                        _P]),
    "asp_flow_weights_bytes": (C.c_size_t, [_I32, _I32]),
    "asp_flow_logprob_bf16": (C.c_int, [C.POINTER(Flow), _P, _I64, _P, _P, _P]),
    "asp_flow_sample_bf16": (C.c_int, [C.POINTER(Flow), _P, _I64, _P, _P, _P]),
    "asp_student_t_workspace_bytes": (C.c_size_t, [_I32]),
    "asp_student_t_fit": (C.c_int, [_P, _I32, _P, _I64, _I64, _I32, _D, _D, _P, _P, _P]),
    "asp_pcn_sweep_flow_supported": (_I32, [_I32, _I32, _I32]),
    "asp_pcn_sweep_flow_workspace_bytes": (C.c_size_t, [_I32]),
    "asp_pcn_sweep_flow": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, C.POINTER(Density), C.POINTER(Density),
                                     C.POINTER(Flow), C.POINTER(PcnParams), _P, _P, _P, _P, _P, _P, _P]),
    "asp_flow_fit_params": (C.c_size_t, [_I32, _I32]),
    "asp_flow_fit_workspace_floats": (C.c_size_t, [_I32, _I32, _I32]),
    "asp_flow_fit_step": (C.c_int, [C.POINTER(FlowTrain), _P, _I32, C.c_float, C.c_float, C.c_float, C.c_float, _I32,
                                    C.c_float, C.c_float, _P, _P]),
    "asp_flow_fit_loss": (C.c_int, [C.POINTER(FlowTrain), _P, _I32, C.c_float, _P, _P]),
    "asp_select_workspace_bytes": (C.c_size_t, []),
    "asp_select_uniforms": (C.c_int, [C.c_uint64, C.c_uint32, _P, _I64, _D, _D, _I64, _P, _P, _P, _P, _P]),
    "asp_philox_uniform": (C.c_int, [C.c_uint64, C.c_uint32, _I64, _I64, _P, _P]),
    "asp_philox_normal": (C.c_int, [C.c_uint64, C.c_uint32, _I64, _I64, _I32, _P, _P, _P, _P]),
    "asp_pcg64_uniform": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _I64, _I64, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the native library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AspireB200Error(
            f"{LIB_PATH} not found: build it with aspire_b200/csrc/build.sh "
            "(or __graft_entry__.build()).  aspire_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the .so drifted apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().asp_last_error().decode("utf-8", "replace")


# kernels launched per successful C call (for bench.py's gpu_launches)
KERNELS_PER_CALL = {"asp_importance_weights": 3, "asp_cdf_exact": 6, "asp_cdf_tile_stats": 2,
                    "asp_cdf_set_tile_stats": 2, "asp_cdf_info": 0, "asp_device_check": 0,
                    "asp_search_multinomial_guided": 4, "asp_select_uniforms": 3, "asp_search_systematic": 3,
                    "asp_search_systematic_workspace_bytes": 0, "asp_pcn_propose_eval_supported": 0, "asp_beta_search_max_passes3": 0,
                    "asp_sorted_tiles": 0, "asp_search_sorted": 2, "asp_pcn_sweep_tpcn_supported": 0, "asp_pcn_sweep_flow_supported": 0,
                    "asp_pcn_sweep_flow_workspace_bytes": 0, "asp_student_t_workspace_bytes": 0, "asp_student_t_fit": 30, "asp_flow_fit_params": 0, "asp_flow_fit_workspace_floats": 0,
                    "asp_beta_search_root": 6, "asp_beta_search_run": 8, "asp_beta_search_scratch_words": 0}
launch_count = 0


def check(rc: int, what: str = "") -> None:
    global launch_count
    launch_count += KERNELS_PER_CALL.get(what, 1)
    if rc != ASP_OK:
        raise AspireB200Error(f"{what or 'aspire_b200'} failed ({rc}): {last_error()}")
