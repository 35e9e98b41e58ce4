"""Device-resident callers of the hot path and their multi-GPU exchange (BASELINE.json configs[2] and [3]).

Everything here is a thin ctypes face over the C ABI (include/sepaihrd_b200.h): the sampler / swarm state lives in HBM
(csrc/sepaihrd_mh.cu, csrc/sepaihrd_swarm.cu), an iteration is a handful of kernel launches on the evaluator's stream, and the
per-iteration collective -- the ranks' log-likelihood blocks of a Metropolis-Hastings run
(reference MetropolisHastingsSampler.cpp:312-330), the per-rank best of a particle swarm (ParticleSwarmOptimizer.cpp:149-156,
417-421) -- is ONE kernel over NVLink peer memory (csrc/sepaihrd_exchange.cu) or, as the library baseline beside it, NCCL's
``all_gather_into_tensor`` on the same device buffers and the same stream.  No host work and no synchronisation per iteration.

One process per GPU (``torchrun``); rendezvous and the 64-byte IPC handles travel over ``torch.distributed``.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Dict, Optional

import numpy as np

from . import capi
from .distributed import shard_range

MH_POSITIONS, MH_LOGPOST, MH_SCALES, MH_ACCEPTED, MH_BEST_LOGPOST, MH_BEST_POSITIONS, MH_ACCEPT_MATRIX, MH_TRACE, MH_PROPOSALS, MH_FAULT = range(10)
HANDLE_BYTES = 64


class _MhSettings(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("burn_in", C.c_int32), ("adapt_scale", C.c_int32), ("record_accepts", C.c_int32),
                ("target_acceptance_rate", C.c_double)]


def std_mt19937_raw(seed: int, count: int) -> np.ndarray:
    """The first ``count`` 32-bit outputs of std::mt19937(seed) (numpy's MT19937 with the classic init_genrand seeding)."""
    bg = np.random.MT19937()
    bg._legacy_seeding(int(seed) & 0xFFFFFFFF)
    return bg.random_raw(int(count)).astype(np.uint32)


class Exchange:
    """All-gather of small device records between the ranks of one node.

    transport "p2p": sepaihrd_exchange_* (CUDA IPC mailboxes, one kernel per all-gather); "nccl": torch.distributed
    all_gather_into_tensor on the evaluator's stream; "single": world size 1 (a device copy).  ``transport=None`` picks p2p
    when the mailboxes can be mapped, else nccl, and says which in ``self.transport`` / ``self.fallback_reason``."""

    def __init__(self, ev, max_doubles: int, rank: int = 0, world: int = 1, transport: Optional[str] = None, torch_device=None):
        self.L = capi.load_library()
        self.ev, self.rank, self.world, self.max_doubles = ev, int(rank), int(world), int(max_doubles)
        self._h = None
        self.fallback_reason = None
        self.torch_device = torch_device
        if self.world == 1:
            transport = "single"
        want = transport or os.environ.get("SEPAIHRD_EXCHANGE", "p2p")
        if want in ("p2p", "single"):
            try:
                self._open_mailboxes()
                self.transport = "single" if self.world == 1 else "p2p"
            except capi.SepaihrdError as exc:
                if transport == "p2p":
                    raise
                self.fallback_reason = str(exc)
                want = "nccl"
        if want == "nccl":
            import torch.distributed as dist
            assert dist.is_initialized(), "the nccl transport needs an initialised process group"
            self.transport = "nccl"

    def _open_mailboxes(self):
        h = C.c_void_p()
        mine = (C.c_ubyte * HANDLE_BYTES)()
        capi.check(self.L.sepaihrd_exchange_create(self.ev.handle, self.world, self.rank, self.max_doubles, C.byref(h), mine))
        self._h = h
        if self.world > 1:
            import torch.distributed as dist
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine))
            blob = (C.c_ubyte * (HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(handles))
            ok = 1
            try:
                capi.check(self.L.sepaihrd_exchange_connect(self._h, blob))
            except capi.SepaihrdError as exc:
                ok, err = 0, exc
            flags = [None] * self.world
            dist.all_gather_object(flags, ok)             # all ranks take the same transport
            if not all(flags):
                self.close()
                raise err if not ok else capi.SepaihrdError(3, "a peer could not map the exchange mailboxes")

    def all_gather(self, src_ptr: int, count: int, dst_ptr: int, src_tensor=None, dst_tensor=None):
        """d_src [count] -> d_dst [world][count], enqueued on the evaluator's stream."""
        if self.transport in ("p2p", "single"):
            capi.check(self.L.sepaihrd_exchange_all_gather(self._h, C.c_void_p(src_ptr), int(count), C.c_void_p(dst_ptr)))
        else:
            import torch.distributed as dist
            dist.all_gather_into_tensor(dst_tensor, src_tensor)

    def status(self) -> int:
        if self._h is None:
            return 0
        v = C.c_int32()
        capi.check(self.L.sepaihrd_exchange_status(self._h, C.byref(v)))
        return v.value

    def close(self):
        if self._h is not None:
            self.L.sepaihrd_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceSwarm:
    """Asynchronous device-resident particle swarm shard (sepaihrd_swarm_*_async)."""

    def __init__(self, ev, swarm_size: int, offset: int, local: int):
        self.L = capi.load_library()
        self.ev, self.P = ev, ev.problem.n_params
        self.swarm_size, self.offset, self.local = int(swarm_size), int(offset), int(local)
        h = C.c_void_p()
        capi.check(self.L.sepaihrd_swarm_create(ev.handle, self.swarm_size, self.offset, self.local, C.byref(h)))
        self._h = h

    def upload_seeds(self, seeds: np.ndarray):
        s = np.ascontiguousarray(seeds, dtype=np.uint32)
        assert s.ndim == 2 and s.shape[1] == self.swarm_size
        capi.check(self.L.sepaihrd_swarm_upload_seeds(self._h, s.ctypes.data, s.shape[0]))

    def init(self, initial=None):
        x = None if initial is None else np.ascontiguousarray(initial, dtype=np.float64)
        capi.check(self.L.sepaihrd_swarm_init_async(self._h, None if x is None else x.ctypes.data))

    def evaluate(self):
        capi.check(self.L.sepaihrd_swarm_evaluate_async(self._h))

    def record_ptr(self):
        p = C.c_void_p(); n = C.c_int32()
        capi.check(self.L.sepaihrd_swarm_record_device(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def adopt(self, records_ptr: int, n_records: int, stride: int, trace_slot: int):
        capi.check(self.L.sepaihrd_swarm_adopt_global_best(self._h, C.c_void_p(records_ptr), int(n_records), int(stride), int(trace_slot)))

    def step(self, iteration: int, omega: float, c1: float, c2: float):
        capi.check(self.L.sepaihrd_swarm_step_async(self._h, int(iteration), float(omega), float(c1), float(c2)))

    def trace(self, n: int) -> np.ndarray:
        out = np.empty(n)
        capi.check(self.L.sepaihrd_swarm_read_trace(self._h, out.ctypes.data, n))
        return out

    def global_best(self):
        v = C.c_double(); pos = np.empty(self.P)
        capi.check(self.L.sepaihrd_swarm_read_global_best(self._h, C.byref(v), pos.ctypes.data))
        return v.value, pos

    def read(self, what: int) -> np.ndarray:
        out = np.empty((self.local, self.P)) if what in (0, 1, 2) else np.empty(self.local)
        if out.size:
            capi.check(self.L.sepaihrd_swarm_read(self._h, int(what), out.ctypes.data))
        return out

    def close(self):
        if self._h is not None:
            self.L.sepaihrd_swarm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceMH:
    """Device-resident Metropolis-Hastings chains (sepaihrd_mh_*): ``local`` of ``n_chains`` chains on the evaluator's GPU."""

    def __init__(self, ev, n_chains: int, offset: int, local: int, iterations: int, burn_in: Optional[int] = None, adapt_scale: bool = True,
                 record_accepts: bool = True, target_acceptance_rate: float = 0.234):
        self.L = capi.load_library()
        self.ev, self.P = ev, ev.problem.n_params
        self.n_chains, self.offset, self.local, self.iterations = int(n_chains), int(offset), int(local), int(iterations)
        st = _MhSettings(int(iterations), int(iterations if burn_in is None else burn_in), int(bool(adapt_scale)), int(bool(record_accepts)),
                         float(target_acceptance_rate))
        h = C.c_void_p()
        capi.check(self.L.sepaihrd_mh_create(ev.handle, self.n_chains, self.offset, self.local, C.byref(st), C.byref(h)))
        self._h = h

    def begin(self, seed: int, initial, chol_lower):
        x = np.ascontiguousarray(initial, dtype=np.float64)
        L = np.asfortranarray(np.asarray(chol_lower, dtype=np.float64))          # column-major, like Eigen
        assert L.shape == (self.P, self.P)
        capi.check(self.L.sepaihrd_mh_begin(self._h, int(seed) & 0xFFFFFFFF, x.ctypes.data, L.ctypes.data))

    def iterate(self, n: int = 1):
        capi.check(self.L.sepaihrd_mh_iterate(self._h, int(n)))

    def propose(self):
        capi.check(self.L.sepaihrd_mh_propose(self._h))

    def evaluate(self):
        capi.check(self.L.sepaihrd_mh_evaluate(self._h))

    def accept(self):
        capi.check(self.L.sepaihrd_mh_accept(self._h))

    # look-ahead windows: K iterations of every chain per likelihood launch (sepaihrd_mh_window_*)
    def window_reserve(self, K: int):
        capi.check(self.L.sepaihrd_mh_window_reserve(self._h, int(K)))

    def window_propose(self, K: int):
        capi.check(self.L.sepaihrd_mh_window_propose(self._h, int(K)))

    def window_evaluate(self):
        capi.check(self.L.sepaihrd_mh_window_evaluate(self._h))

    def window_commit(self, record_stride: int = 0):
        capi.check(self.L.sepaihrd_mh_window_commit(self._h, int(record_stride)))

    def window_record_ptr(self) -> int:
        p = C.c_void_p()
        capi.check(self.L.sepaihrd_mh_window_record(self._h, C.byref(p)))
        return p.value

    def window_progress(self) -> int:
        """Synchronises; the smallest next-iteration index among the local chains (``iterations`` once all are done)."""
        v = C.c_int32()
        capi.check(self.L.sepaihrd_mh_window_progress(self._h, C.byref(v)))
        return v.value

    def run_windows(self, K: int) -> int:
        """All iterations of the local chains in look-ahead windows of K; returns the number of windows (likelihood launches)."""
        n = 0
        while True:
            self.window_propose(K); self.window_evaluate(); self.window_commit()
            n += 1
            if self.window_progress() >= self.iterations:
                return n

    @property
    def iteration(self) -> int:
        return int(self.L.sepaihrd_mh_iteration(self._h))

    def logpost_ptr(self) -> int:
        p = C.c_void_p()
        capi.check(self.L.sepaihrd_mh_logpost_device(self._h, C.byref(p)))
        return p.value

    def note_gathered(self, all_ptr: int, world: int, stride: int, slot: int):
        capi.check(self.L.sepaihrd_mh_note_gathered(self._h, C.c_void_p(all_ptr), int(world), int(stride), int(slot)))

    def read(self, what: int) -> np.ndarray:
        done = self.iteration - 1
        shape, dt = {MH_POSITIONS: ((self.local, self.P), np.float64), MH_LOGPOST: ((self.local,), np.float64),
                     MH_SCALES: ((self.local,), np.float64), MH_ACCEPTED: ((self.local,), np.int64),
                     MH_BEST_LOGPOST: ((self.local,), np.float64), MH_BEST_POSITIONS: ((self.local, self.P), np.float64),
                     MH_ACCEPT_MATRIX: ((done, self.local), np.uint8), MH_TRACE: ((self.iterations + 1,), np.float64),
                     MH_PROPOSALS: ((self.local, self.P), np.float64), MH_FAULT: ((1,), np.uint32)}[what]
        out = np.zeros(shape, dtype=dt)
        if out.size:
            capi.check(self.L.sepaihrd_mh_read(self._h, int(what), out.ctypes.data))
        return out

    def close(self):
        if self._h is not None:
            self.L.sepaihrd_mh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def initial_cholesky(sigmas, regularization_epsilon: float = 1e-6) -> np.ndarray:
    """Lower Cholesky factor of the sampler's start covariance (MetropolisHastingsSampler.cpp:225-240): diag(sigma^2, or 1e-6
    where sigma <= 0) scaled by 2.38^2 / P, plus the ridge -- a diagonal matrix, so the factor is the element-wise root."""
    s = np.asarray(sigmas, dtype=np.float64)
    P = len(s)
    d = np.where(s > 0, s * s, 1e-6) * ((2.38 * 2.38) / float(P)) + regularization_epsilon
    return np.diag(np.sqrt(d))


class _Phases:
    """CUDA-event timing of the phases of an iteration on the evaluator's stream (events cost ~1 us each; no syncs)."""

    def __init__(self, stream, names):
        import torch
        self.torch, self.stream, self.names = torch, stream, list(names)
        self.marks = []

    def mark(self):
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.stream)
        self.marks.append(e)

    def totals(self) -> Dict[str, float]:
        """Marks are laid down as m0 [phase0] m1 [phase1] m2 ... cyclically; returns seconds per phase name."""
        out = {n: 0.0 for n in self.names}
        k = len(self.names)
        for i in range(len(self.marks) - 1):
            out[self.names[i % k]] += self.marks[i].elapsed_time(self.marks[i + 1]) * 1e-3
        return out


def run_pso_resident(ev, swarm_size: int, iterations: int, seed: int, initial=None, rank: int = 0, world: int = 1,
                     transport: Optional[str] = None, settings: Optional[Dict[str, float]] = None, torch_device=None,
                     return_positions: bool = False):
    """STANDARD / GLOBAL_BEST particle swarm (ParticleSwarmOptimizer.cpp:330-425, 576-618, 149-156), particles sharded over
    the ranks, with NOTHING on the host per iteration: update kernel, fused likelihood kernel, personal bests + shard best, the
    all-gather of one 64-double record per rank, the global-best selection -- all enqueued on the evaluator's stream.  Visits
    exactly the positions of the host swarm (host/optimizers.cpp) seeded alike.  Returns the global-best trace and timings."""
    import torch
    st = dict(omega_start=0.9, omega_end=0.4, c1_initial=2.5, c1_final=0.5, c2_initial=0.5, c2_final=2.5)
    st.update(settings or {})
    dev = torch_device if torch_device is not None else torch.device("cuda", torch.cuda.current_device())
    stream = torch.cuda.current_stream(dev)
    ev.set_stream(stream.cuda_stream)
    lo, hi = shard_range(swarm_size, rank, world)
    t0 = time.perf_counter()
    sw = DeviceSwarm(ev, swarm_size, lo, hi - lo)
    # one seed per particle of the WHOLE swarm per draw of the master generator: set 0 initialises, set 1 + it steps (:268-270, :365-371)
    sw.upload_seeds(std_mt19937_raw(seed, (iterations + 1) * swarm_size).reshape(iterations + 1, swarm_size))
    rec_ptr, rec_n = sw.record_ptr()
    ex = Exchange(ev, rec_n, rank, world, transport, dev)
    gathered = torch.empty((world, rec_n), dtype=torch.float64, device=dev)
    rec_view = None
    if ex.transport == "nccl":
        rec_view = _tensor_view(rec_ptr, rec_n, dev)
    t_setup = time.perf_counter() - t0
    ph = _Phases(stream, ["eval", "exchange", "update"])

    def coefficients(it):
        ratio = (it / (iterations - 1)) if iterations > 1 else 0.0
        return (st["omega_start"] + (st["omega_end"] - st["omega_start"]) * ratio,
                st["c1_initial"] + (st["c1_final"] - st["c1_initial"]) * ratio,
                st["c2_initial"] + (st["c2_final"] - st["c2_initial"]) * ratio)

    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    sw.init(initial)
    ph.mark()
    for it in range(iterations + 1):
        sw.evaluate()
        ph.mark()
        ex.all_gather(rec_ptr, rec_n, gathered.data_ptr(), rec_view, gathered)
        sw.adopt(gathered.data_ptr(), world, rec_n, it)
        ph.mark()
        if it < iterations:
            sw.step(it, *coefficients(it))
        ph.mark()
    trace = sw.trace(iterations + 1)                      # the one synchronisation of the run
    t_run = time.perf_counter() - t1
    best_val, best_pos = sw.global_best()
    out = dict(rank=rank, world=world, particles=(lo, hi), trace=trace, best_value=best_val, best_position=best_pos,
               setup_seconds=t_setup, run_seconds=t_run, phase_seconds=ph.totals(), transport=ex.transport,
               fallback_reason=ex.fallback_reason, exchange_status=ex.status(), evaluations=(iterations + 1) * (hi - lo))
    if return_positions:
        out["final_positions"] = sw.read(0)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()                                    # nobody tears its mailbox down while a peer may still write to it
    ex.close(); sw.close()
    return out


def window_length(local: int, requested: Optional[int] = None, rate: float = 0.234) -> int:
    """Iterations per window of the device-resident sampler.  K proposals per chain commit g(K) = (1 - (1 - rate)^K) / rate
    iterations on average; a likelihood launch costs ~0.59 ms while every warp has a scheduler to itself (592 warps of 8 sets:
    4736 sets), ~0.81 ms while two share one (9472 sets), and a further ~0.8 ms per wave of 9472 sets beyond
    (profiles/r02_v17_sets_per_tile_ab.txt); a proposal costs ~5 us of a warp's time, the rest of a window ~0.06 ms.
    K = argmax g(K) / cost(local * K): 11 up to 256 chains per GPU, 9 for 512, 4 for 1024 and for 2048, 2 for 4096."""
    if requested is not None:
        return max(1, min(64, int(requested)))
    best, best_rate, miss = 1, 0.0, 1.0
    for k in range(1, 17):
        miss *= 1.0 - rate
        sets = max(int(local), 1) * k
        launch = 0.59 if sets <= 4736 else 0.81 if sets <= 9472 else 0.81 + 0.8 * -(-(sets - 9472) // 9472)
        r = (1.0 - miss) / rate / (launch + 0.005 * k + 0.06)
        if r > best_rate * 1.02:               # a longer window has to pay by more than the noise
            best, best_rate = k, r
    return best


def run_mh_resident(ev, sigmas, initial, n_chains: int, iterations: int, seed: int, rank: int = 0, world: int = 1,
                    transport: Optional[str] = None, chol_lower=None, torch_device=None, record_accepts: bool = True,
                    adapt_scale: bool = True, lookahead: Optional[int] = 1):
    """``n_chains`` seeded Metropolis-Hastings chains (MetropolisHastingsSampler.cpp:201-412, fixed-kernel phase) sharded over
    the ranks and resident on the devices; per iteration: propose kernel, fused likelihood kernel, accept kernel, all-gather of
    the ranks' log-likelihood blocks, max over all chains into the trace.  Nothing runs on the host inside the loop.  The
    evaluator must be in MCMC_REFLECT mode (constraint_mode 1) like the reference's sampler (:207-210).

    ``lookahead`` = 1: one iteration per likelihood launch (above).  ``lookahead`` = K > 1 or None (sized from the shard,
    ``window_length``): look-ahead windows -- every chain proposes its next K iterations at once, one launch scores them all, every
    chain commits up to its first accepted proposal; the exchange then runs once per WINDOW (the ranks' log-likelihood blocks plus
    each rank's smallest iteration index, from which every rank knows when all chains of the run are done: a small D2H read after
    every second window once the run could be over).  Same decisions and states; ``best_trace`` is then per window and ``windows`` counts them."""
    import torch
    dev = torch_device if torch_device is not None else torch.device("cuda", torch.cuda.current_device())
    stream = torch.cuda.current_stream(dev)
    ev.set_stream(stream.cuda_stream)
    lo, hi = shard_range(n_chains, rank, world)
    block = shard_range(n_chains, 0, world)[1]            # the largest shard: rank 0's
    K = 1 if lookahead == 1 else window_length(block, lookahead)
    windowed = K > 1
    t0 = time.perf_counter()
    mh = DeviceMH(ev, n_chains, lo, hi - lo, iterations, record_accepts=record_accepts, adapt_scale=adapt_scale)
    if windowed:
        mh.window_reserve(K)
    rec_n = block + 1 if windowed else block              # windows: one more double per rank, its smallest iteration index
    ex = Exchange(ev, rec_n, rank, world, transport, dev)
    gathered = torch.zeros((world, rec_n), dtype=torch.float64, device=dev)
    # a rank with a shorter shard still sends `block` doubles: the tail of its buffer is never read (counts are known)
    lp_ptr = mh.logpost_ptr()
    send = torch.zeros(block, dtype=torch.float64, device=dev) if (not windowed and ((hi - lo) < block or ex.transport == "nccl")) else None
    t_setup = time.perf_counter() - t0
    ph = _Phases(stream, ["propose", "eval", "accept", "exchange"])
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    mh.begin(seed, initial, initial_cholesky(sigmas) if chol_lower is None else chol_lower)
    ph.mark()
    windows = 0
    if windowed:
        rec_ptr = rec_view = None
        first_check = max(1, -(-(iterations - 1) // K))
        while True:
            mh.window_propose(K)
            ph.mark()
            mh.window_evaluate()
            ph.mark()
            mh.window_commit(block)
            ph.mark()
            if rec_ptr is None:
                rec_ptr = mh.window_record_ptr()
                rec_view = _tensor_view(rec_ptr, rec_n, dev) if ex.transport == "nccl" else None
            ex.all_gather(rec_ptr, rec_n, gathered.data_ptr(), rec_view, gathered)
            if windows < iterations:
                mh.note_gathered(gathered.data_ptr(), world, rec_n, windows + 1)
            ph.mark()
            windows += 1
            # the host reads the ranks' smallest iteration indices (every rank sees the same values) once the fastest possible
            # run could be over, then after every second window: a window past the end is harmless (nothing left to commit)
            if windows >= first_check and (windows - first_check) % 2 == 0 and float(gathered[:, block].min().item()) >= iterations:
                break
        mh.window_progress()
        trace = mh.read(MH_TRACE)[1:min(windows, iterations) + 1]
    else:
        for it in range(1, iterations):
            mh.propose()
            ph.mark()
            mh.evaluate()
            ph.mark()
            mh.accept()
            ph.mark()
            if send is not None:
                if hi > lo:
                    send[:hi - lo].copy_(_tensor_view(lp_ptr, hi - lo, dev), non_blocking=True)
                ex.all_gather(send.data_ptr(), block, gathered.data_ptr(), send, gathered)
            else:
                ex.all_gather(lp_ptr, block, gathered.data_ptr(), None, gathered)
            mh.note_gathered(gathered.data_ptr(), world, block, it)
            ph.mark()
        trace = mh.read(MH_TRACE)[1:iterations]               # the one synchronisation of the run
    t_run = time.perf_counter() - t1
    if int(mh.read(MH_FAULT)[0]) != 0:
        raise RuntimeError("device-resident sampler: a proposal exhausted its polar attempts (generator fault)")
    out = dict(rank=rank, world=world, chains=(lo, hi), best_trace=trace, x=mh.read(MH_POSITIONS), logpost=mh.read(MH_LOGPOST),
               scale=mh.read(MH_SCALES), accepted=mh.read(MH_ACCEPTED), accepts=mh.read(MH_ACCEPT_MATRIX) if record_accepts else None,
               all_logpost=np.concatenate([gathered[r, :shard_range(n_chains, r, world)[1] - shard_range(n_chains, r, world)[0]].cpu().numpy()
                                           for r in range(world)]),
               setup_seconds=t_setup, run_seconds=t_run, phase_seconds=ph.totals(), transport=ex.transport,
               fallback_reason=ex.fallback_reason, exchange_status=ex.status(),
               evaluations=(windows * K * (hi - lo) + 1) if windowed else ((iterations - 1) * (hi - lo) + 1),
               lookahead=K, windows=windows if windowed else iterations - 1)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    ex.close(); mh.close()
    return out


def _tensor_view(ptr: int, n: int, dev):
    """A float64 torch view of ``n`` doubles of library-owned device memory (no copy, no ownership)."""
    import torch

    class _Iface:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}

    return torch.as_tensor(_Iface(), device=dev)
