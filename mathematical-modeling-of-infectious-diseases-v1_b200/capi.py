"""ctypes binding of the product C ABI (include/sepaihrd_b200.h -> csrc/libsepaihrd_b200.so).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no fallback:
a missing library raises ``RuntimeError`` on first use.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SEPAIHRD_LIB: A/B timing of an experimental build of the same C ABI (tools/ only; tests and bench use the default)
LIB_PATH = os.environ.get("SEPAIHRD_LIB") or os.path.join(_HERE, "csrc", "libsepaihrd_b200.so")
_lib = None

_dp = C.POINTER(C.c_double)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)

# every symbol include/sepaihrd_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "sepaihrd_slot_count": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32]),
    "sepaihrd_slot_for_name": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_char_p]),
    "sepaihrd_create": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "sepaihrd_destroy": (None, [C.c_void_p]),
    "sepaihrd_set_constraint_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "sepaihrd_set_math_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "sepaihrd_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sepaihrd_eval_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sepaihrd_eval_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sepaihrd_simulate_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "sepaihrd_simulate_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "sepaihrd_simulate_from_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "sepaihrd_simulate_from_state_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "sepaihrd_column_quantiles_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "sepaihrd_posterior_predictive": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]),
    "sepaihrd_swarm_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]),
    "sepaihrd_swarm_destroy": (None, [C.c_void_p]),
    "sepaihrd_swarm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "sepaihrd_swarm_evaluate": (C.c_int, [C.c_void_p, _dp, C.POINTER(C.c_int64), C.c_void_p]),
    "sepaihrd_swarm_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]),
    "sepaihrd_swarm_read": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "sepaihrd_swarm_upload_seeds": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "sepaihrd_swarm_init_async": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sepaihrd_swarm_evaluate_async": (C.c_int, [C.c_void_p]),
    "sepaihrd_swarm_record_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), _i32p]),
    "sepaihrd_swarm_adopt_global_best": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32]),
    "sepaihrd_swarm_step_async": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double]),
    "sepaihrd_swarm_read_trace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32]),
    "sepaihrd_swarm_read_global_best": (C.c_int, [C.c_void_p, _dp, C.c_void_p]),
    "sepaihrd_mh_create": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]),
    "sepaihrd_mh_destroy": (None, [C.c_void_p]),
    "sepaihrd_mh_begin": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "sepaihrd_mh_iterate": (C.c_int, [C.c_void_p, C.c_int32]),
    "sepaihrd_mh_propose": (C.c_int, [C.c_void_p]),
    "sepaihrd_mh_evaluate": (C.c_int, [C.c_void_p]),
    "sepaihrd_mh_accept": (C.c_int, [C.c_void_p]),
    "sepaihrd_mh_iteration": (C.c_int32, [C.c_void_p]),
    "sepaihrd_mh_logpost_device": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "sepaihrd_mh_note_gathered": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32]),
    "sepaihrd_mh_read": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "sepaihrd_mh_window_reserve": (C.c_int, [C.c_void_p, C.c_int32]),
    "sepaihrd_mh_window_propose": (C.c_int, [C.c_void_p, C.c_int32]),
    "sepaihrd_mh_window_evaluate": (C.c_int, [C.c_void_p]),
    "sepaihrd_mh_window_commit": (C.c_int, [C.c_void_p, C.c_int64]),
    "sepaihrd_mh_window_record": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "sepaihrd_mh_window_progress": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "sepaihrd_exchange_create": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "sepaihrd_exchange_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sepaihrd_exchange_all_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "sepaihrd_exchange_status": (C.c_int, [C.c_void_p, _i32p]),
    "sepaihrd_exchange_destroy": (None, [C.c_void_p]),
    "sepaihrd_set_ordering": (C.c_int, [C.c_void_p, C.c_int32]),
    "sepaihrd_fit_ordering": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32]),
    "sepaihrd_ordering_state": (C.c_int, [C.c_void_p, _i32p, C.POINTER(C.c_int64)]),
    "sepaihrd_release_scratch": (C.c_int, [C.c_void_p]),
    "sepaihrd_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "sepaihrd_free_pinned": (None, [C.c_void_p]),
    "sepaihrd_synchronize": (C.c_int, [C.c_void_p]),
    "sepaihrd_get_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "sepaihrd_get_merge_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "sepaihrd_measure_fp64_peak": (C.c_int, [C.c_int32, _dp]),
    "sepaihrd_last_error": (C.c_char_p, []),
    "sepaihrd_version": (C.c_char_p, []),
}

RC_NAMES = {0: "OK", 1: "INVALID_ARGUMENT", 2: "NO_DEVICE", 3: "CUDA", 4: "UNSUPPORTED", 5: "OUT_OF_MEMORY"}


class SepaihrdError(RuntimeError):
    def __init__(self, rc: int, msg: str):
        super().__init__(f"sepaihrd_b200: {RC_NAMES.get(rc, rc)}: {msg}")
        self.rc = rc


def load_library():
    """dlopen csrc/libsepaihrd_b200.so and bind every declared symbol (raises if any is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (nvcc build). "
                               "sepaihrd_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise SepaihrdError(rc, (load_library().sepaihrd_last_error() or b"").decode())
