"""ctypes binding of libsapt.so (C-ABI declared in include/sapt.h).

There is NO CPU fallback: if the CUDA library is missing or cannot be loaded this module
raises at import of the first symbol, and every wrapper raises on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SAPT_LIB", os.path.join(_HERE, "libsapt.so"))  # SAPT_LIB: developer override for A/B builds

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM = 0, -1, -2, -3, -4
PATH_AUTO, PATH_F32, PATH_TC = 0, 1, 2
STREAM_Z1, STREAM_Z2, STREAM_LX1, STREAM_KU, STREAM_LX2, STREAM_U, STREAM_SWAP, STREAM_LS_INIT = range(8)
F_POST_TEMPERED, F_TAKE_POST_LIK, F_DENORM_HALF_MIN, F_COMPARE_SURROGATE, F_REDRAW_ON_SURROGATE, F_RECORD_POS, F_LOCAL_SURROGATE, F_SWEEP_PARALLEL = (
    1 << 0, 1 << 1, 1 << 2, 1 << 3, 1 << 4, 1 << 5, 1 << 6, 1 << 7)
F_SWEEP_GRID = 1 << 8

vp = C.c_void_p
i32, i64, u32, f64 = C.c_int32, C.c_int64, C.c_uint32, C.c_double


class ProposeArgs(C.Structure):
    _fields_ = [("base", vp), ("proposal", vp), ("ld", i64), ("R", C.c_int), ("P", C.c_int),
                ("slot_of_row", vp), ("slot_offset", C.c_int), ("rows", vp), ("n_rows", vp),
                ("step_w", f64), ("sigma_squared", f64), ("prior_prop", vp),
                ("seed", u32), ("kind", C.c_int), ("step", C.c_int), ("z_inject", vp),
                ("decide", C.c_int), ("redraw", C.c_int), ("surrogate_ready", C.c_int),
                ("surrogate_prob", f64), ("gate", C.c_int), ("z2_inject", vp), ("ku_inject", vp),
                ("is_sur", vp), ("true_rows", vp), ("n_true", vp), ("sur_rows", vp), ("n_sur", vp), ("step_dev", vp)]


class ChainState(C.Structure):
    _fields_ = [("slot_of_row", vp), ("temperature", vp), ("lik_cur", vp), ("lik_post", vp), ("prior_cur", vp),
                ("mov", vp), ("likl_wo_temp", vp), ("last_rmse_train", vp), ("last_rmse_test", vp),
                ("last_acc_train", vp), ("last_acc_test", vp), ("naccept", vp)]


class History(C.Structure):
    _fields_ = [("n_slots", C.c_int), ("lhood_list", vp), ("rmse_train", vp), ("rmse_test", vp), ("acc_train", vp),
                ("acc_test", vp), ("surg", vp), ("likeh", vp), ("accept_list", vp), ("accepted", vp), ("is_sur", vp),
                ("pos_w", vp)]


class AcceptArgs(C.Structure):
    _fields_ = [("theta", vp), ("proposal", vp), ("ld", i64), ("R", C.c_int), ("P", C.c_int),
                ("st", ChainState), ("hist", History), ("step", C.c_int), ("flags", u32), ("is_sur", vp),
                ("L_train", vp), ("nc_train", vp), ("sq_train", vp), ("L_test", vp), ("nc_test", vp), ("sq_test", vp),
                ("n_train", i64), ("n_test", i64), ("prior_prop", vp), ("sur_s", vp), ("min_y", f64), ("max_y", f64),
                ("seed", u32), ("u_inject", vp), ("train_theta", vp), ("train_L", vp), ("train_count", vp),
                ("train_cap", C.c_int), ("capture", C.c_int), ("surg_first", vp), ("step_dev", vp),
                ("ls_pred", vp), ("ls_avg", vp)]


class LgFirstArgs(C.Structure):
    _fields_ = [("proposal", vp), ("ld", i64), ("R", C.c_int), ("P", C.c_int), ("st", ChainState),
                ("step", C.c_int), ("seed", u32), ("true_rows", vp), ("n_true", vp), ("L_train", vp),
                ("surg_first", vp), ("use_langevin", C.c_int), ("l_prob", f64), ("lx2_inject", vp),
                ("lang_rows", vp), ("n_lang", vp), ("train_theta", vp), ("train_L", vp), ("train_count", vp),
                ("train_cap", C.c_int), ("step_dev", vp)]


class FusedArgs(C.Structure):
    _fields_ = [("theta", vp), ("ld", i64), ("R", C.c_int), ("P", C.c_int), ("step_w", f64), ("sigma_squared", f64), ("seed", u32),
                ("use_surrogate", C.c_int), ("surrogate_ready", C.c_int), ("surrogate_prob", f64), ("gate", C.c_int),
                ("W1", vp), ("b1", vp), ("h1", C.c_int), ("W2", vp), ("b2", vp), ("h2", C.c_int), ("W3", vp), ("b3", vp),
                ("acc", AcceptArgs), ("n_steps", C.c_int)]


class LocalSurrogate(C.Structure):
    _fields_ = [("weights", vp), ("limit", vp), ("surg", vp), ("n_slots", C.c_int), ("slot_of_row", vp), ("R", C.c_int)]


LS_WEIGHTS = 61


class StepArgs(C.Structure):
    _fields_ = [("prop", ProposeArgs), ("acc", AcceptArgs), ("use_surrogate", C.c_int), ("compare_surrogate", C.c_int),
                ("W1", vp), ("b1", vp), ("h1", C.c_int), ("W2", vp), ("b2", vp), ("h2", C.c_int), ("W3", vp), ("b3", vp),
                ("sur_s", vp), ("L_train", vp), ("nc_train", vp), ("sq_train", vp), ("L_test", vp), ("nc_test", vp), ("sq_test", vp),
                ("counters", vp), ("path", C.c_int), ("n_steps", C.c_int)]


class SaptError(RuntimeError):
    pass


_lib = None


def lib():
    """Load libsapt.so once. Fails loudly — the product path has no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.sapt_last_error.restype = C.c_char_p
    L.sapt_launch_count.restype = C.c_longlong
    L.sapt_launch_count_reset.restype = None
    L.sapt_problem_rows.restype = i64
    L.sapt_problem_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, i64, vp, i64, C.c_int]
    L.sapt_problem_create_f32.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, vp, i64, vp, vp, i64, C.c_int]
    L.sapt_problem_destroy.argtypes = [vp]
    L.sapt_problem_num_param.argtypes = [vp]
    L.sapt_problem_rows.argtypes = [vp, C.c_int]
    L.sapt_problem_tc_supported.argtypes = [vp]
    L.sapt_loglik.argtypes = [vp, C.c_int, vp, i64, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp]
    L.sapt_loglik_host.argtypes = [vp, C.c_int, vp, i64, C.c_int, vp, vp, vp, C.c_int]
    L.sapt_evaluate_proposal.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    L.sapt_prior.argtypes = [vp, i64, C.c_int, C.c_int, f64, vp, vp]
    L.sapt_propose.argtypes = [C.POINTER(ProposeArgs), vp]
    L.sapt_langevin.argtypes = [vp, vp, vp, i64, C.c_int, vp, vp, f64, C.c_int, f64, vp]
    L.sapt_surrogate_predict.argtypes = [vp, i64, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp]
    L.sapt_mh_accept.argtypes = [C.POINTER(AcceptArgs), vp]
    L.sapt_lg_first_eval.argtypes = [C.POINTER(LgFirstArgs), vp]
    L.sapt_swap_sweep.argtypes = [C.c_int, vp, vp, vp, u32, u32, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
    L.sapt_swap_sweep_dev.argtypes = [C.c_int, vp, vp, vp, u32, u32, C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
    L.sapt_advance.argtypes = [vp, C.c_int, vp]
    L.sapt_philox_host.argtypes = [C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    L.sapt_philox_fill.argtypes = [u32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    L.sapt_fused_steps.argtypes = [vp, C.POINTER(FusedArgs), vp]
    L.sapt_local_surrogate_train.argtypes = [C.POINTER(LocalSurrogate), C.c_int, C.c_int, C.c_int, vp, u32, f64, C.c_int, vp, vp]
    L.sapt_local_surrogate_predict.argtypes = [C.POINTER(LocalSurrogate), C.c_int, vp, vp, vp, vp, vp, vp]
    L.sapt_step.argtypes = [vp, C.POINTER(StepArgs), vp]
    L.sapt_struct_sizes.argtypes = [C.POINTER(i32)]
    L.sapt_measure_pipes.argtypes = [C.c_int, C.POINTER(f64), C.c_int, vp]
    L.sapt_loglik_work_model.argtypes = [vp, C.c_int, C.POINTER(f64), C.c_int]
    L.sapt_train_rows_pack.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
    L.sapt_train_rows_scatter.argtypes = [vp, i64, C.c_int, vp, C.c_int, vp, vp]
    L.sapt_train_rows_plan.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
    sizes = (i32 * 8)()
    L.sapt_struct_sizes(sizes)
    mine = [C.sizeof(t) for t in (ProposeArgs, ChainState, History, AcceptArgs, LgFirstArgs, FusedArgs, LocalSurrogate, StepArgs)]
    if list(sizes) != mine:
        raise ImportError(f"struct layout mismatch between include/sapt.h and _lib.py: C={list(sizes)} ctypes={mine}")
    if L.sapt_abi_version() != 1:
        raise ImportError("libsapt.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().sapt_last_error().decode(errors="replace")
        exc = ValueError if rc == ERR_INVALID else (NotImplementedError if rc == ERR_UNSUPPORTED else SaptError)
        raise exc(f"{what or 'libsapt'} failed ({rc}): {msg}")


EXPORTS = [
    "sapt_abi_version", "sapt_last_error", "sapt_launch_count", "sapt_launch_count_reset", "sapt_struct_sizes",
    "sapt_problem_create", "sapt_problem_create_f32", "sapt_problem_destroy", "sapt_problem_num_param",
    "sapt_problem_rows", "sapt_problem_tc_supported", "sapt_loglik", "sapt_loglik_host", "sapt_evaluate_proposal",
    "sapt_prior", "sapt_propose", "sapt_fused_steps", "sapt_langevin", "sapt_surrogate_predict", "sapt_mh_accept", "sapt_lg_first_eval",
    "sapt_swap_sweep", "sapt_swap_sweep_dev", "sapt_advance", "sapt_philox_host", "sapt_philox_fill",
    "sapt_local_surrogate_train", "sapt_local_surrogate_predict", "sapt_step", "sapt_measure_pipes", "sapt_loglik_work_model", "sapt_train_rows_pack", "sapt_train_rows_scatter", "sapt_train_rows_plan",
]


PIPE_NAMES = ("sms", "sm_hz", "ex2", "rcp", "ffma", "fmul", "ffma2", "lop3", "mix", "sm_hz_mufu", "f2fp")


def measure_pipes(device: int = 0) -> dict:
    """SM pipe micro-benchmarks (sapt_measure_pipes): lane-operations per second and per clock per SM."""
    out = (f64 * len(PIPE_NAMES))()
    check(lib().sapt_measure_pipes(device, out, len(PIPE_NAMES), None), "sapt_measure_pipes")
    d = dict(zip(PIPE_NAMES, [float(v) for v in out]))
    for k in ("ex2", "rcp", "ffma", "fmul", "ffma2", "lop3", "mix", "f2fp"):
        hz = d["sm_hz_mufu"] if k in ("ex2", "rcp") else d["sm_hz"]
        d[k + "_per_clk_sm"] = d[k] / (d["sms"] * hz)
    return d


def work_model(problem_handle, path: int = PATH_AUTO) -> dict:
    """Per-evaluation work the selected likelihood kernel executes (sapt_loglik_work_model)."""
    out = (f64 * 5)()
    check(lib().sapt_loglik_work_model(problem_handle, path, out, 5), "sapt_loglik_work_model")
    return dict(path=int(out[0]), mufu=float(out[1]), mma_flops=float(out[2]), ffma=float(out[3]), flops_algorithmic=float(out[4]))
