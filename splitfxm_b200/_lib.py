"""ctypes binding of libsfxm_b200.so (the C ABI in include/sfxm_b200.h).

There is no CPU fallback: importing this module without the built library, or calling into it without a
CUDA device, raises.  Build with ``splitfxm_b200/csrc/build.sh`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os

from .error import SFXM

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFXM_LIB") or os.path.join(_HERE, "csrc", "libsfxm_b200.so")   # SFXM_LIB: kernel-variant builds

MAX_NV, MAX_NB, MAX_PARAMS = 4, 2, 4

MODEL_AD, MODEL_EULER_PRIM, MODEL_EULER_CONS = 0, 1, 2
MODEL_USER0, MAX_USER_MODELS = 16, 16
FV_LAX_FRIEDRICHS, FV_CENTRAL, FV_UPWIND, FD_CENTRAL, FD_RIGHT_BIAS = 0, 1, 2, 3, 4
BC_PERIODIC, BC_OUTFLOW, BC_NEUMANN, BC_DIRICHLET, BC_KEEP, BC_HALO = 0, 1, 2, 3, 4, 5
GX_PERIODIC, GX_MIRROR, GX_KEEP = 0, 1, 2

c_i32, c_i64, c_f64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p


class Problem(ctypes.Structure):
    """struct sfxm_problem (include/sfxm_b200.h)."""
    _fields_ = [
        ("model", c_i32), ("scheme", c_i32),
        ("nv", c_i32), ("nbl", c_i32), ("nbr", c_i32),
        ("n_params", c_i32),
        ("nx", c_i64), ("batch", c_i64),
        ("xmin", c_f64), ("xmax", c_f64),
        ("bc_ltype", c_i32 * MAX_NV), ("bc_rtype", c_i32 * MAX_NV),
        ("bc_lval", (c_f64 * MAX_NB) * MAX_NV), ("bc_rval", (c_f64 * MAX_NB) * MAX_NV),
        ("gx_left", c_i32), ("gx_right", c_i32),
        ("gx_keep_left", c_f64 * MAX_NB), ("gx_keep_right", c_f64 * MAX_NB),
        ("var_base", c_i64 * MAX_NV),
        ("var_stride", c_i32 * MAX_NV), ("var_off", c_i32 * MAX_NV), ("var_seg", c_i32 * MAX_NV),
        ("halo_left", c_vp), ("halo_right", c_vp),
        ("peer_left_dst", c_vp), ("peer_right_dst", c_vp),
    ]


NEWTON_THOMAS, NEWTON_PCR, NEWTON_BAND4, NEWTON_FUSED_STEP = 0, 1, 2, 3


class NewtonPlan(ctypes.Structure):
    """struct sfxm_newton_plan (include/sfxm_b200.h): the buffers of one Newton / Split-Newton run (sfxm_newton_run)."""
    _fields_ = [
        ("prob", ctypes.POINTER(Problem)),
        ("x_int", c_vp), ("params", c_vp),
        ("param_stride", c_i64),
        ("geom", c_vp),
        ("eps", c_f64), ("abs_tol", c_f64), ("rel_tol", c_f64),
        ("nseg", c_i32), ("solver", c_i32),
        ("row_lo", c_i64 * MAX_NV), ("row_hi", c_i64 * MAX_NV),
        ("solve_launches", c_i32 * MAX_NV),
        ("y", c_vp), ("s", c_vp), ("r", c_vp), ("J", c_vp),
        ("work_solve", c_vp), ("work_norms", c_vp), ("norms2", c_vp),
        ("status", c_vp),
        ("it", c_vp), ("crit", c_vp), ("run", c_vp),
        ("seg_ss", c_vp), ("seg_yy", c_vp),
        ("any_dev", c_vp), ("any_host", c_vp),
    ]


# name -> (restype, argtypes); every symbol include/sfxm_b200.h declares
_PP = ctypes.POINTER(Problem)
SIGNATURES = {
    "sfxm_abi_version": (ctypes.c_int, []),
    "sfxm_last_error": (ctypes.c_char_p, []),
    "sfxm_residual": (ctypes.c_int, [_PP, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_f64, c_vp]),
    "sfxm_residual_warp_count": (c_i64, [_PP]),
    "sfxm_residual_warps": (ctypes.c_int, [_PP, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_f64, c_i64, c_i64, c_vp]),
    "sfxm_geometry": (ctypes.c_int, [_PP, c_vp, c_vp, c_vp]),
    "sfxm_fill_domain": (ctypes.c_int, [_PP, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_jacobian": (ctypes.c_int, [_PP, c_vp, c_vp, c_i64, c_vp, c_vp, c_f64, c_vp, c_vp, c_vp]),
    "sfxm_jacobian_segments": (ctypes.c_int, [_PP, c_vp, c_vp, c_i64, c_vp, c_vp, c_f64, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "sfxm_pack_jacobian": (ctypes.c_int, [_PP, c_vp, c_vp, ctypes.POINTER(c_i32), c_vp]),
    "sfxm_solve_workspace_bytes": (c_i64, [_PP]),
    "sfxm_block_tridiag_solve": (ctypes.c_int, [_PP, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_newton_fused_supported": (ctypes.c_int, [_PP]),
    "sfxm_newton_solve_fused": (ctypes.c_int, [_PP, c_vp, c_vp, c_i64, c_vp, c_vp, c_f64, c_vp, c_vp, c_i32, c_vp, c_vp,
                                               c_vp, c_vp, c_vp]),
    "sfxm_newton_step_fused": (ctypes.c_int, [_PP, c_vp, c_vp, c_i64, c_vp, c_vp, c_f64, c_vp, c_vp, c_i32, c_vp, c_vp,
                                              c_vp, c_vp, c_vp, c_vp]),
    "sfxm_pcr_workspace_bytes": (c_i64, [_PP]),
    "sfxm_block_tridiag_solve_pcr": (ctypes.c_int, [_PP, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_band4_workspace_bytes": (c_i64, [_PP]),
    "sfxm_block_band4_solve_pcr": (ctypes.c_int, [_PP, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_spike_workspace_bytes": (c_i64, [_PP]),
    "sfxm_spike_reduce": (ctypes.c_int, [_PP, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_spike_fold": (ctypes.c_int, [_PP, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_dense_solve": (ctypes.c_int, [c_vp, c_i32, c_i64, c_vp, c_vp, c_vp]),
    "sfxm_sumsq_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "sfxm_sumsq": (ctypes.c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "sfxm_update_norms": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i32, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "sfxm_newton_begin": (ctypes.c_int, [c_i64, c_vp, c_vp, c_vp, c_f64, c_vp, c_f64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_newton_end": (ctypes.c_int, [c_i64, c_vp, c_vp, c_vp, c_i32, c_f64, c_f64, c_vp, c_vp, c_vp, c_vp, c_vp, c_f64,
                                       c_vp, c_vp, c_vp]),
    "sfxm_newton_run": (ctypes.c_int, [ctypes.POINTER(NewtonPlan), c_f64, ctypes.POINTER(c_i32), ctypes.POINTER(c_i64), c_vp]),
    "sfxm_axpy_masked": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp]),
    "sfxm_rk_lincomb": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.POINTER(c_f64), c_i32, c_f64, c_vp, c_i64, c_vp]),
    "sfxm_rk_dense_end": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.POINTER(c_f64), c_i32, c_f64, c_vp, c_i64, c_vp]),
    "sfxm_scaled_sumsq_workspace_bytes": (c_i64, [c_i64, c_i64]),
    "sfxm_scaled_diff_sumsq": (ctypes.c_int, [c_vp, c_vp, c_vp, c_f64, c_f64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "sfxm_rk_error_sumsq": (ctypes.c_int, [c_vp, c_i64, ctypes.POINTER(c_f64), c_i32, c_f64, c_vp, c_vp, c_f64, c_f64,
                                           c_i64, c_i64, c_vp, c_vp, c_vp]),
    "sfxm_rk45_begin": (ctypes.c_int, [c_vp, c_vp, c_vp]),
    "sfxm_rk45_control": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "sfxm_rk_lincomb_dev": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.POINTER(c_f64), c_i32, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "sfxm_rk_error_sumsq_dev": (ctypes.c_int, [c_vp, c_i64, ctypes.POINTER(c_f64), c_i32, c_vp, c_vp, c_vp, c_f64, c_f64,
                                               c_i64, c_i64, c_vp, c_vp, c_vp]),
    "sfxm_rk45_commit": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(c_f64), c_i32, c_vp, c_vp, c_vp, c_i64, c_vp]),
    "sfxm_refine_workspace_bytes": (c_i64, [c_i64]),
    "sfxm_refine_mark": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_f64, c_f64, c_f64, c_f64, c_f64, c_f64, c_vp, c_vp, c_vp]),
    "sfxm_refine_flags": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_refine_apply": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_refine_fused_workspace_bytes": (c_i64, [c_i64]),
    "sfxm_refine_fused": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_f64, c_f64, c_f64, c_f64, c_f64, c_f64, c_vp, c_vp, c_vp,
                                         c_vp, c_vp]),
    "sfxm_refine_fused_flags": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_refine_set_flags": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "sfxm_refine_window_stats": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "sfxm_refine_window_mark": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_i64, c_i64, c_i64, c_i64, c_f64, c_f64, c_f64,
                                               c_f64, c_f64, c_f64, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_refine_window_finalize": (ctypes.c_int, [c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "sfxm_refine_window_apply": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i32, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_refine_window_flags": (ctypes.c_int, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_register_user_model": (ctypes.c_int, [c_i32, c_vp]),
    "sfxm_halo_signal": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "sfxm_halo_wait": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "sfxm_halo_signal_counter": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "sfxm_halo_wait_counter": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_fv_fluxes": (ctypes.c_int, [c_i32, c_i32, c_vp, c_vp, c_f64, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_fv_diffusion_fluxes": (ctypes.c_int, [c_i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_fd_derivative": (ctypes.c_int, [c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_fv_cell_residual": (ctypes.c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_f64, c_vp, c_vp, c_vp]),
    "sfxm_fd_cell_residual": (ctypes.c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "sfxm_selftest_div": (ctypes.c_int, [c_i64, ctypes.c_uint64, c_i32, c_vp, c_vp]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises SFXM if it has not been built - never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SFXM(f"{LIB_PATH} is missing: build it with splitfxm_b200/csrc/build.sh "
                   "(there is no CPU fallback for the hot path)")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)    # user-model plugins resolve sfxm:: symbols against it
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().sfxm_last_error().decode("utf-8", "replace")
        raise SFXM(msg or f"libsfxm_b200 error {status}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    return _torch().cuda.current_stream().cuda_stream


_torch_mod = None


def _torch():
    global _torch_mod
    if _torch_mod is None:
        import torch
        _torch_mod = torch
    return _torch_mod


AMR_MARK_LAUNCHES = 10   # kernels sfxm_refine_mark issues (csrc/sfxm_amr.cu)
AMR_FUSED_LAUNCHES = 3   # kernels of sfxm_refine_fused
LAUNCHES = [0]   # number of kernel launches issued through the C ABI (bench.py reports it)


def count(n=1):
    LAUNCHES[0] += n
