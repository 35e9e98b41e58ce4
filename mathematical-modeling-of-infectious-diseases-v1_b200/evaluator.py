"""Batch evaluator: the Python face of the C ABI.

``BatchEvaluator.eval_batch`` is the batched form of the reference's
``SEPAIHRDObjectiveFunction::calculate`` (src/model/objectives/SEPAIHRDObjectiveFunction.cpp:62-235);
``simulate_batch`` is the batched ``AgeSEPAIHRDSimulator::run``.  Host (numpy) arrays go through
``sepaihrd_eval_batch`` (H2D + kernel + D2H inside the call); CUDA ``torch.Tensor`` arguments go
through the ``*_device`` entry points on the current torch stream with no copies.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import capi
from .problem import Problem, TRAJ_FULL

MATH_FAST, MATH_STRICT, MATH_FAST_GENERAL = 0, 1, 2   # FAST_GENERAL: FAST with the general kernel build (verification)
MATH_FAST_SPLIT = 3                                   # FAST on the warp-pair kernel (csrc/sepaihrd_split.cuh), bit-identical
PPC_PROBS = (0.025, 0.05, 0.5, 0.95, 0.975)     # ResultAggregator.cpp:233
PPC_SERIES = ("daily_hospitalizations", "daily_icu_admissions", "daily_deaths",
              "cumulative_hospitalizations", "cumulative_icu_admissions", "cumulative_deaths")


def _is_tensor(x) -> bool:
    return type(x).__module__.startswith("torch")


class BatchEvaluator:
    def __init__(self, problem: Problem, device: int = -1, math: int = MATH_FAST,
                 constraint_mode: Optional[int] = None):
        self.problem = problem
        self._lib = capi.load_library()
        self._cp = problem.as_c(constraint_mode)
        h = C.c_void_p()
        capi.check(self._lib.sepaihrd_create(C.byref(self._cp), int(device), C.byref(h)))
        self._h = h
        self.device = device
        if math != MATH_FAST:
            self.set_math_mode(math)

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.sepaihrd_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ----------------------------------------------------------------------------
    @property
    def handle(self):
        """The sepaihrd_ctx* (for C-ABI calls that take the context, e.g. the device-resident swarm)."""
        return self._h

    def set_constraint_mode(self, mode: int):
        """SEPAIHRDParameterManager::setConstraintMode."""
        capi.check(self._lib.sepaihrd_set_constraint_mode(self._h, int(mode)))

    def set_math_mode(self, mode: int):
        capi.check(self._lib.sepaihrd_set_math_mode(self._h, int(mode)))

    def set_stream(self, cuda_stream_ptr: int):
        capi.check(self._lib.sepaihrd_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def set_ordering(self, on: bool):
        """Switch the ordering pass in front of large launches (csrc/sepaihrd_order.cu) on (default) or off."""
        capi.check(self._lib.sepaihrd_set_ordering(self._h, 1 if on else 0))

    def fit_ordering(self, params):
        """Fit the ordering model on a representative batch ([B, P] numpy array or CUDA tensor); ~15 ms, synchronous."""
        if _is_tensor(params):
            assert params.is_cuda and params.dim() == 2 and params.is_contiguous()
            self.set_stream(__import__("torch").cuda.current_stream(params.device).cuda_stream)
            capi.check(self._lib.sepaihrd_fit_ordering(self._h, C.c_void_p(params.data_ptr()), params.shape[0], params.shape[1], 1))
        else:
            x = np.ascontiguousarray(params, dtype=np.float64)
            capi.check(self._lib.sepaihrd_fit_ordering(self._h, x.ctypes.data, x.shape[0], x.shape[1], 0))

    def ordering_state(self):
        """(model fitted?, number of fits so far)."""
        a, b = C.c_int32(), C.c_int64()
        capi.check(self._lib.sepaihrd_ordering_state(self._h, C.byref(a), C.byref(b)))
        return bool(a.value), b.value

    def synchronize(self):
        capi.check(self._lib.sepaihrd_synchronize(self._h))

    def counters(self):
        a, b = C.c_int64(), C.c_int64()
        capi.check(self._lib.sepaihrd_get_counters(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def merge_counters(self):
        """(launches that served more than one concurrent eval_batch call, calls served by them)."""
        a, b = C.c_int64(), C.c_int64()
        capi.check(self._lib.sepaihrd_get_merge_counters(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- evaluation -------------------------------------------------------------------------------
    def eval_batch(self, params, return_steps: bool = False):
        """logL for each row of ``params`` ([B, P] float64).  numpy in -> numpy out (synchronous);
        CUDA tensor in -> CUDA tensors out, enqueued on the current torch stream."""
        if _is_tensor(params):
            return self._eval_device(params, return_steps)
        x = np.ascontiguousarray(params, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError("params must be [B, P]")
        B, ld = x.shape
        ll = np.empty(B)
        st = np.zeros(B, dtype=np.uint32)
        steps = np.zeros((B, 2), dtype=np.int32) if return_steps else None
        capi.check(self._lib.sepaihrd_eval_batch(self._h, x.ctypes.data, B, ld, ll.ctypes.data, st.ctypes.data,
                                                 steps.ctypes.data if steps is not None else None))
        return (ll, st, steps) if return_steps else (ll, st)

    def _eval_device(self, params, return_steps):
        import torch
        assert params.is_cuda and params.dtype == torch.float64 and params.dim() == 2 and params.is_contiguous()
        B, ld = params.shape
        ll = torch.empty(B, dtype=torch.float64, device=params.device)
        st = torch.empty(B, dtype=torch.int32, device=params.device)
        steps = torch.empty((B, 2), dtype=torch.int32, device=params.device) if return_steps else None
        self.set_stream(torch.cuda.current_stream(params.device).cuda_stream)
        capi.check(self._lib.sepaihrd_eval_batch_device(self._h, params.data_ptr(), B, ld, ll.data_ptr(), st.data_ptr(),
                                                        steps.data_ptr() if steps is not None else None))
        return (ll, st, steps) if return_steps else (ll, st)

    def eval_into(self, d_params_ptr: int, B: int, ld: int, d_ll_ptr: int, d_status_ptr: int = 0, d_steps_ptr: int = 0):
        """Raw device-pointer form (no allocation): used by bench.py's timed loop."""
        capi.check(self._lib.sepaihrd_eval_batch_device(self._h, d_params_ptr, B, ld, d_ll_ptr,
                                                        d_status_ptr or None, d_steps_ptr or None))

    def eval_host_into(self, params_ptr: int, B: int, ld: int, ll_ptr: int, status_ptr: int = 0, steps_ptr: int = 0):
        """Raw host-pointer form (pinned or pageable): H2D + kernel + D2H + sync inside the C call."""
        capi.check(self._lib.sepaihrd_eval_batch(self._h, params_ptr, B, ld, ll_ptr, status_ptr or None, steps_ptr or None))

    def simulate_batch(self, params, what: int = TRAJ_FULL, stride: int = 1):
        p = self.problem
        W = p.state_size if what == TRAJ_FULL else 3 * p.n_ages
        rows = (p.n_times + stride - 1) // stride
        if _is_tensor(params):
            import torch
            B, ld = params.shape
            out = torch.empty((B, rows, W), dtype=torch.float64, device=params.device)
            st = torch.empty(B, dtype=torch.int32, device=params.device)
            self.set_stream(torch.cuda.current_stream(params.device).cuda_stream)
            capi.check(self._lib.sepaihrd_simulate_batch_device(self._h, params.data_ptr(), B, ld, int(what), int(stride),
                                                                out.data_ptr(), st.data_ptr()))
            return out, st
        x = np.ascontiguousarray(params, dtype=np.float64)
        B, ld = x.shape
        out = np.empty((B, rows, W))
        st = np.zeros(B, dtype=np.uint32)
        capi.check(self._lib.sepaihrd_simulate_batch(self._h, x.ctypes.data, B, ld, int(what), int(stride),
                                                     out.ctypes.data, st.ctypes.data))
        return out, st

    def simulate_from_state(self, params, initial_states, what: int = TRAJ_FULL, stride: int = 1):
        """Batched Simulator::run(initial_state, times): ``initial_states`` is [11n] (shared) or [B, 11n]."""
        p = self.problem
        x = np.ascontiguousarray(params, dtype=np.float64)
        s0 = np.ascontiguousarray(initial_states, dtype=np.float64)
        B, ld = x.shape
        if s0.ndim == 1:
            sstride = 0
        else:
            if s0.shape[0] != B:
                raise ValueError("one initial state per parameter set (or a single shared state)")
            sstride = s0.shape[1]
        if s0.shape[-1] != p.state_size:
            raise ValueError("Initial state size does not match model state size.")
        W = p.state_size if what == TRAJ_FULL else 3 * p.n_ages
        rows = (p.n_times + stride - 1) // stride
        out = np.empty((B, rows, W))
        st = np.zeros(B, dtype=np.uint32)
        capi.check(self._lib.sepaihrd_simulate_from_state(self._h, x.ctypes.data, B, ld, s0.ctypes.data, sstride, int(what),
                                                          int(stride), out.ctypes.data, st.ctypes.data))
        return out, st



    def posterior_predictive(self, params, initial_state, probs=PPC_PROBS):
        """ResultAggregator::aggregatePosteriorPredictives on the device: quantiles [6, T, n_ages, len(probs)] of the six
        PPC_SERIES over the draws (rows of ``params``), all simulated from ``initial_state``; also the number of valid draws."""
        p = self.problem
        x = np.ascontiguousarray(params, dtype=np.float64)
        s0 = np.ascontiguousarray(initial_state, dtype=np.float64)
        if s0.shape != (p.state_size,):
            raise ValueError("Initial state size does not match model state size.")
        pr = np.ascontiguousarray(probs, dtype=np.float64)
        T = int((p.times >= 0).sum())
        out = np.empty((6, T, p.n_ages, len(pr)))
        valid = C.c_int64()
        capi.check(self._lib.sepaihrd_posterior_predictive(self._h, x.ctypes.data, x.shape[0], x.shape[1], s0.ctypes.data, len(pr), pr.ctypes.data,
                                                           out.ctypes.data, C.byref(valid)))
        return out, valid.value

    SELECT_BY_SIZE, SELECT_BLOCK, SELECT_CLUSTER = 0, 1, 2

    def column_quantiles(self, columns, probs, path=0):
        """Exact quantiles (numpy's default interpolation, NaNs left out) of device-resident columns: ``columns`` is a CUDA float64
        tensor [n_cols, B] (contiguous); returns a CUDA tensor [n_cols, len(probs)].  ``path`` picks the kernel (SELECT_*)."""
        import torch
        if columns.dtype != torch.float64 or not columns.is_cuda or not columns.is_contiguous() or columns.dim() != 2:
            raise ValueError("columns: contiguous CUDA float64 tensor [n_cols, B]")
        self.set_stream(torch.cuda.current_stream(columns.device).cuda_stream)
        pr = np.ascontiguousarray(probs, dtype=np.float64)
        out = torch.empty((columns.shape[0], len(pr)), dtype=torch.float64, device=columns.device)
        capi.check(self._lib.sepaihrd_column_quantiles_device(self._h, columns.data_ptr(), columns.shape[1], columns.shape[0], len(pr),
                                                              pr.ctypes.data, out.data_ptr(), int(path)))
        return out


def measure_fp64_peak(device: int = 0) -> float:
    """Measured FP64 pipe peak in DFMA instructions per second (lane-ops): roofline denominator."""
    v = C.c_double()
    capi.check(capi.load_library().sepaihrd_measure_fp64_peak(int(device), C.byref(v)))
    return v.value
