"""ctypes binding of the C ABI declared in include/pymra_b200.h.

The shared library is built in-tree by `__graft_entry__.build()` (explicit nvcc, sm_100a).
There is no fallback: if the library is missing, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpymra_b200.so")

MRA_OK = 0
STATUS_NAMES = {0: "MRA_OK", -1: "MRA_ERR_ARG", -2: "MRA_ERR_CUDA", -3: "MRA_ERR_STATE",
                -4: "MRA_ERR_NOT_SPD", -5: "MRA_ERR_NOMEM"}
WARN_NEGATIVE_VARIANCE = 2
COV_EXP, COV_MATERN32, COV_MATERN52, COV_GAUSSIAN, COV_DENSE = 0, 1, 2, 3, 4

_p32 = C.POINTER(C.c_int32)
_p64 = C.POINTER(C.c_int64)
_pd = C.POINTER(C.c_double)


class MraStructure(C.Structure):
    _fields_ = [("n_locs", C.c_int64), ("dim", C.c_int32), ("r", C.c_int32), ("depth", C.c_int32),
                ("n_nodes", C.c_int32), ("node_level", _p32), ("node_parent", _p32), ("node_kind", _p32),
                ("node_row_start", _p64), ("node_row_count", _p64), ("node_child_start", _p32),
                ("node_child_count", _p32), ("node_knot_off", _p64), ("knot_rows", _p64),
                ("n_knot_rows", C.c_int64), ("level_off", _p32), ("perm", _p64)]


# name -> (restype, argtypes); mirrors include/pymra_b200.h one to one
SIGNATURES = {
    "mra_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "mra_destroy": (C.c_int, [C.c_void_p]),
    "mra_last_error": (C.c_char_p, [C.c_void_p]),
    "mra_version": (C.c_char_p, []),
    "mra_build_structure_2d": (C.c_int, [_pd, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.POINTER(C.c_uint32), _p32, C.c_int32, _p32, _p32, _p32, _p32, _p32,
                                         _p64, _p64, _p32, _p32, _p64, _p64, _p32, _p64, _p64, _p32]),
    "mra_build_stream_start": (C.c_int, [_pd, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.POINTER(C.c_uint32), _p32, C.c_int32, _p32, _p32, _p32, _p64, _p64,
                                         _p32, _p32, _p64, _p64, _p32, _p64, _p32, C.POINTER(C.c_void_p)]),
    "mra_build_stream_wait": (C.c_int, [C.c_void_p, C.c_int32]),
    "mra_build_stream_finish": (C.c_int, [C.c_void_p, _p32, _p32, _p64]),
    "mra_set_structure": (C.c_int, [C.c_void_p, C.POINTER(MraStructure)]),
    "mra_expect_shard": (C.c_int, [C.c_void_p]),
    "mra_plan": (C.c_int, [C.c_void_p, _pd, C.c_int, C.POINTER(C.c_size_t)]),
    "mra_bind_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mra_plan_tree": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]),
    "mra_bind_tree": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mra_plan_obs": (C.c_int, [C.c_void_p, _pd, C.POINTER(C.c_size_t)]),
    "mra_bind_obs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mra_upload_data": (C.c_int, [C.c_void_p, _pd, _pd, C.c_void_p]),
    "mra_upload_data_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mra_set_cov": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double]),
    "mra_set_nugget": (C.c_int, [C.c_void_p, C.c_double]),
    "mra_set_cov_dense": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double]),
    "mra_run_likelihood": (C.c_int, [C.c_void_p, C.c_void_p, _pd]),
    "mra_run_predict": (C.c_int, [C.c_void_p, C.c_void_p, _pd, _pd]),
    "mra_run_likelihood_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mra_run_predict_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mra_fetch_likelihood": (C.c_int, [C.c_void_p, C.c_void_p, _pd]),
    "mra_predict_rows": (C.c_int, [C.c_void_p, _p64, C.c_int32, _p32]),
    "mra_run_predict_pack_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "mra_unpermute_tree_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mra_run_graph": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "mra_set_shard": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int8)]),
    "mra_summary_size": (C.c_int, [C.c_void_p, _p64]),
    "mra_run_likelihood_local_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mra_run_likelihood_top_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mra_stream_parts": (C.c_int, [C.c_void_p, _p32]),
    "mra_stream_begin_async": (C.c_int, [C.c_void_p, C.c_void_p, _p64]),
    "mra_stream_part_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _p64]),
    "mra_stream_part_prior_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, _p64]),
    "mra_stream_end_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mra_stream_my_parts": (C.c_int, [C.c_void_p, _p32]),
    "mra_stream_end_local_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "mra_stream_sync_knots": (C.c_int, [C.c_void_p, C.c_void_p, _p64]),
    "mra_set_diagnostics": (C.c_int, [C.c_void_p, C.c_int]),
    "mra_last_warnings": (C.c_int, [C.c_void_p, _p32]),
    "mra_last_launches": (C.c_int, [C.c_void_p, _p64]),
    "mra_last_flops": (C.c_int, [C.c_void_p, _pd, _pd]),
    "mra_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "mra_profile_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "mra_debug_fetch": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_int, _pd, C.c_int64]),
}


class MraError(RuntimeError):
    def __init__(self, status, message):
        RuntimeError.__init__(self, "%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise RuntimeError(
            "pymra_b200: CUDA library %s not found. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = load_library()
    return _LIB


def check(handle, status):
    if status != MRA_OK:
        msg = lib().mra_last_error(handle)
        raise MraError(status, msg.decode() if msg else "")
