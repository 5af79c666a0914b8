"""ctypes binding of libchk_b200.so (the C ABI declared in include/chk_b200.h).

The shared library is built in-tree (``make -C complexhyperbolickge_b200/csrc`` or
``__graft_entry__.build()``).  There is no CPU fallback and no alternative backend: if the library is
missing, or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libchk_b200.so")
CSRC = os.path.join(_HERE, "csrc")

CHK_ROT, CHK_REF, CHK_ATT = 0, 1, 2
CHK_F32, CHK_F64 = 0, 1
CHK_RANK_FMA, CHK_RANK_MMA = 0, 1

_i, _i64, _p = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p
CHK_MAX_TABLES = 8


class TableDesc(ctypes.Structure):
    """chk_table_desc of include/chk_b200.h."""
    _fields_ = [("param", _p), ("grad", _p), ("state_sum", _p), ("rows", _p), ("m", _i64), ("src_rows", _p),
                ("width", _i64), ("stamp", _p)]


CHK_OPT_NONE, CHK_OPT_ADAGRAD, CHK_OPT_ADAM = 0, 1, 2
CHK_HYPER_LEN = 8
CHK_RED_MAX_COLS, CHK_RED_MAX_GROUPS = 4, 3


class RedCol(ctypes.Structure):
    """chk_red_col of include/chk_b200.h."""
    _fields_ = [("param", _p), ("state0", _p), ("dense_grad", _p), ("width", _i64), ("src", _p * 2), ("lo", _i64 * 2),
                ("hi", _i64 * 2), ("rank_stride", _i64 * 2), ("pair_coef", _p), ("pair_nt", _i64), ("coef_rank_stride", _i64)]


class RedGroup(ctypes.Structure):
    """chk_red_group of include/chk_b200.h."""
    _fields_ = [("ids", _p), ("n_keys", _i64), ("slots_per_rank", _i64), ("world", ctypes.c_int32), ("n_cols", ctypes.c_int32),
                ("single_row", ctypes.c_int32), ("pad_", ctypes.c_int32), ("work", _p), ("cols", RedCol * CHK_RED_MAX_COLS)]


class EvalArgs(ctypes.Structure):
    """chk_eval_args of include/chk_b200.h."""
    _fields_ = [("algo", ctypes.c_int32), ("kind", ctypes.c_int32), ("dtype", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("multi_c", ctypes.c_int32), ("pad_", ctypes.c_int32), ("b", _i64), ("queries", _p),
                ("entity", _p), ("rel", _p), ("rel_diag", _p), ("ctx", _p), ("c_table", _p), ("bh", _p), ("bt", _p),
                ("hn_full", _p), ("n_entities", _i64), ("shard_entity", _p), ("shard_hn", _p), ("shard_bt", _p),
                ("shard_rows", _i64), ("shard_offset", _i64), ("shadow", _p), ("workspace", _p), ("workspace_bytes", _i64),
                ("f_keys", _p), ("f_indptr", _p), ("f_vals", _p), ("f_nkeys", _i64), ("n_rel2", _i64),
                ("scratch", _p), ("scratch_bytes", _i64), ("counts", _p), ("target", _p), ("flags", _p)]


class DenseTab(ctypes.Structure):
    """chk_dense_tab of include/chk_b200.h."""
    _fields_ = [("param", _p), ("grad", _p), ("state0", _p), ("state1", _p), ("n", _i64)]


# name -> (restype, argtypes); mirrors include/chk_b200.h line by line
SIGNATURES = {
    "chk_abi_version": (_i, []),
    "chk_last_error": (ctypes.c_char_p, []),
    "chk_query_fwd": (_i, [_i, _i, _i, _i64, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "chk_group_by_key": (_i, [_p, _i64, _i, _p, _p, _p]),
    "chk_query_fwd_grouped": (_i, [_i, _i, _i, _i64, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "chk_query_bwd": (_i, [_i, _i, _i, _i64, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "chk_score_gather_fwd": (_i, [_i, _i, _i64, _i64, _p, _i64, _i64, _p, _p, _i64, _p, _i64, _i64, _p, _p, _p]),
    "chk_score_gather_bwd": (_i, [_i, _i, _i64, _i64, _p, _i64, _i64, _p, _p, _i64, _p, _p, _p, _p]),
    "chk_score_gather_bwd_scatter": (_i, [_i, _i, _i64, _i64, _p, _i64, _i64, _p, _p, _p, _p, _p, _p]),
    "chk_nsloss": (_i, [_i, _i64, _i64, _p, _p, _p, _p]),
    "chk_sparse_adagrad": (_i, [_i, _p, _p, _p, _p, _i64, _i64, ctypes.c_double, ctypes.c_double, _p, _p, _p]),
    "chk_step_counter_bump": (_i, [_p, _p]),
    "chk_claim_gather_rows": (_i, [_i, _p, _p, _i64, _i64, _p, _p, _p, _p]),
    "chk_multi_scatter_add": (_i, [_i, ctypes.POINTER(TableDesc), _i, _p]),
    "chk_multi_sparse_adagrad": (_i, [_i, ctypes.POINTER(TableDesc), _i, ctypes.c_double, ctypes.c_double, _p, _p]),
    "chk_scatter_add_rows": (_i, [_i, _p, _p, _p, _i64, _i64, _p]),
    "chk_train_prep": (_i, [_p, _i64, _i64, _i64, _i, _p, _p, ctypes.c_uint64, _p, ctypes.c_uint32, _p, _p, _p, _p]),
    "chk_score_gather_train": (_i, [_i, _i, _i64, _i64, _p, _i64, _i64, _p, _p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "chk_group_workspace_bytes": (_i64, [_i64, _i64]),
    "chk_group_build": (_i, [_p, _i64, _i64, _i64, _i64, _p, _p]),
    "chk_score_gather_train_peer": (_i, [_i, _i, _i64, _i64, _p, _i64, _i64, _p, _p, _i64, _i, _p, _p, _i64, _i64, _p, _p, _p, _p, _p, _p, _p,
                                         _p, _p]),
    "chk_peer_gather_rows": (_i, [_i, _p, _i64, _p, _i64, _i64, _p, _p]),
    "chk_reduce_apply": (_i, [_i, _i, ctypes.POINTER(RedGroup), _i, _p, _i, _p, _i64, _p, _p, _p]),
    "chk_step_finish": (_i, [_i, ctypes.POINTER(_p), _i, _p, _i64, _p, _p, _p]),
    "chk_dense_apply": (_i, [_i, _i, ctypes.POINTER(DenseTab), _i, _p, _p, _p]),
    "chk_rowsum_groups": (_i, [_i, _p, _i64, _i64, _i64, _p, _p]),
    "chk_dp_fused_apply": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "chk_dp_all_gather": (_i, [_i, _i, _p, _i64, _p, _p, _i, _p, _p]),
    "chk_reg_factors": (_i, [_i, _i, ctypes.c_double, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _p, _p, _i64, _p, _i64, _p, _i64,
                             _p, _i64, _p, _p]),
    "chk_row_hnorm": (_i, [_i, _i, _i64, _p, _p, _p]),
    "chk_score_all": (_i, [_i, _i, _i64, _p, _p, _p, _p, _p, _p, _i64, _p, _p]),
    "chk_target_scores": (_i, [_i, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _p]),
    "chk_rank_counts": (_i, [_i, _i, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _i64, _p, _p, _i64, _p, _p, _i64,
                             _p, _p]),
    "chk_entity_shadow_bytes": (_i64, [_i, _i64]),
    "chk_entity_shadow_build": (_i, [_i, _i, _i64, _p, _p, _p, _p, _p]),
    "chk_rank_mma_workspace_bytes": (_i64, [_i, _i64]),
    "chk_rank_mma_reset": (_i, [_p, _p]),
    "chk_rank_mma_status": (_i, [_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int), _p]),
    "chk_rank_mma_profile_events": (_i, [_p, _p]),
    "chk_filter_lookup": (_i, [_p, _i64, _i64, _p, _i64, _p, _p, _i64, _i64, _p, _p, _p, _p, _p]),
    "chk_eval_scratch_bytes": (_i64, [_i, _i, _i64]),
    "chk_eval_batch": (_i, [ctypes.POINTER(EvalArgs), _p]),
    "chk_score_all_mma": (_i, [_i, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _i64, _p, _p, _p, _p]),
}

_lib = None


def build_native(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into libchk_b200.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libchk_b200.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or __graft_entry__.build()). "
                "complexhyperbolickge_b200 has no CPU / eager fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {lib().chk_last_error().decode()}")
