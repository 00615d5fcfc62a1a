"""ctypes shim over libsorobn_b200.so (the C ABI in include/sorobn_b200.h).

This is the "thin C-ABI/ctypes shim" between the Python host (`BayesNet.query`) and
the CUDA kernels.  There is deliberately no fallback: if the library has not been
built (`python -m sorobn_b200.csrc.build` / `__graft_entry__.build()`), cannot be
loaded, or no sm_100 GPU is visible, construction raises.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsorobn_b200.so")
_lib = None

SBN_OK = 0
ABI_VERSION = 9


class EngineError(RuntimeError):
    """An error reported by libsorobn_b200; `code` is the SBN_E_* value (-3 = device memory)."""

    def __init__(self, message, code=None):
        super().__init__(message)
        self.code = code


SBN_E_NOMEM = -3


def lib_path() -> str:
    return _LIB_PATH


def load():
    """Load the shared library and declare every entry point of include/sorobn_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise EngineError(
            f"{_LIB_PATH} is missing: build the CUDA library first "
            "(python -m sorobn_b200.csrc.build, or __graft_entry__.build()). "
            "sorobn_b200 has no CPU fallback for exact inference."
        )
    lib = ctypes.CDLL(_LIB_PATH)
    c = ctypes
    vp, i64, i32 = c.c_void_p, c.c_int64, c.c_int
    lib.sbn_abi_version.restype = i32
    lib.sbn_abi_version.argtypes = []
    lib.sbn_last_error.restype = c.c_char_p
    lib.sbn_last_error.argtypes = []
    lib.sbn_device_count.restype = i32
    lib.sbn_device_count.argtypes = [c.POINTER(i32)]
    lib.sbn_program_create.restype = i32
    lib.sbn_program_create.argtypes = [i32, vp, i64, vp, i64, c.POINTER(vp)]
    lib.sbn_program_create_f64.restype = i32
    lib.sbn_program_create_f64.argtypes = [i32, vp, i64, vp, i64, c.POINTER(vp)]
    lib.sbn_program_run_host_f64.restype = i32
    lib.sbn_program_run_host_f64.argtypes = [vp, vp, i64, i64, vp, i64]
    lib.sbn_program_evidence_host.restype = i32
    lib.sbn_program_evidence_host.argtypes = [vp, vp, i64, i64, vp]
    lib.sbn_program_evidence_host_f64.restype = i32
    lib.sbn_program_evidence_host_f64.argtypes = [vp, vp, i64, i64, vp]
    lib.sbn_program_destroy.restype = None
    lib.sbn_program_destroy.argtypes = [vp]
    lib.sbn_program_reserve.restype = i32
    lib.sbn_program_reserve.argtypes = [vp, i64]
    lib.sbn_program_run_host.restype = i32
    lib.sbn_program_run_host.argtypes = [vp, vp, i64, i64, vp, i64]
    lib.sbn_program_run_device.restype = i32
    lib.sbn_program_run_device.argtypes = [vp, vp, i64, i64, vp, i64, vp]
    lib.sbn_program_step_roles.restype = i32
    lib.sbn_program_step_roles.argtypes = [vp, vp, i64]
    lib.sbn_program_profile.restype = i32
    lib.sbn_program_profile.argtypes = [vp, vp, i64, i64, vp, i64, vp, vp, i64]
    lib.sbn_program_info.restype = i32
    lib.sbn_program_info.argtypes = [vp, vp, i64]
    lib.sbn_program_set_graph.restype = i32
    lib.sbn_program_set_graph.argtypes = [vp, i32]
    lib.sbn_program_set_tiled.restype = i32
    lib.sbn_program_set_tiled.argtypes = [vp, i32]
    lib.sbn_gibbs_create.restype = i32
    lib.sbn_gibbs_create.argtypes = [i32, i32, vp, vp, vp, vp, vp, i64, i32, vp, i32, vp, i32, vp, c.POINTER(vp)]
    lib.sbn_gibbs_run_host.restype = i32
    lib.sbn_gibbs_run_host.argtypes = [vp, vp, i64, i64, i64, c.c_uint64, vp, i64]
    lib.sbn_sampler_run_host.restype = i32
    lib.sbn_sampler_run_host.argtypes = [vp, i32, vp, i64, i64, i64, c.c_uint64, vp, i64]
    lib.sbn_gibbs_conditional.restype = i32
    lib.sbn_gibbs_conditional.argtypes = [vp, i32, vp, vp]
    lib.sbn_gibbs_destroy.restype = None
    lib.sbn_gibbs_destroy.argtypes = [vp]
    lib.sbn_host_alloc.restype = i32
    lib.sbn_host_alloc.argtypes = [c.POINTER(vp), i64]
    lib.sbn_host_free.restype = i32
    lib.sbn_host_free.argtypes = [vp]
    if lib.sbn_abi_version() != ABI_VERSION:
        raise EngineError(f"libsorobn_b200.so has ABI {lib.sbn_abi_version()}, Python expects {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


EXPORTS = (
    "sbn_abi_version", "sbn_last_error", "sbn_device_count", "sbn_program_create", "sbn_program_create_f64",
    "sbn_program_run_host_f64", "sbn_program_evidence_host", "sbn_program_evidence_host_f64", "sbn_program_destroy",
    "sbn_program_reserve", "sbn_program_run_host", "sbn_program_run_device", "sbn_program_profile",
    "sbn_program_step_roles",
    "sbn_program_info", "sbn_program_set_graph", "sbn_program_set_tiled", "sbn_gibbs_create", "sbn_gibbs_run_host",
    "sbn_sampler_run_host", "sbn_gibbs_conditional", "sbn_gibbs_destroy", "sbn_host_alloc", "sbn_host_free",
)


def _check(rc: int):
    if rc != SBN_OK:
        raise EngineError(f"libsorobn_b200 error {rc}: {load().sbn_last_error().decode(errors='replace')}", code=rc)


def device_count() -> int:
    n = ctypes.c_int(0)
    rc = load().sbn_device_count(ctypes.byref(n))
    return n.value if rc == SBN_OK else 0


def default_device() -> int:
    return int(os.environ.get("SOROBN_B200_DEVICE", "0"))


class PinnedArray:
    """A numpy array backed by page-locked host memory (cudaHostAlloc) so that the
    engine's host<->device copies are true asynchronous DMA."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        nbytes = max(1, int(np.prod(self.shape)) * self.dtype.itemsize)
        self._ptr = ctypes.c_void_p()
        _check(load().sbn_host_alloc(ctypes.byref(self._ptr), nbytes))
        buf = (ctypes.c_char * nbytes).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            load().sbn_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Program:
    """One compiled (query variables, evidence variables) pair on one GPU."""

    def __init__(self, plan, device: int | None = None, f64: bool = False):
        lib = load()
        self.plan = plan
        self.device = default_device() if device is None else int(device)
        self.Q = int(plan.Q)
        self.n_ev = len(plan.evidence)
        self.f64 = bool(f64)
        self._h = ctypes.c_void_p()
        words = np.ascontiguousarray(plan.words, dtype=np.int32)
        if self.f64:
            blob = np.ascontiguousarray(plan.table_blob64, dtype=np.float64)
            _check(lib.sbn_program_create_f64(self.device, words.ctypes.data, words.size, blob.ctypes.data, blob.size,
                                              ctypes.byref(self._h)))
            return
        blob = np.ascontiguousarray(plan.table_blob, dtype=np.float32)
        _check(lib.sbn_program_create(self.device, words.ctypes.data, words.size, blob.ctypes.data, blob.size,
                                      ctypes.byref(self._h)))
        # developer switches (profiling / A-B runs); the defaults are the fast path
        if os.environ.get("SOROBN_B200_TILED"):
            self.set_tiled(int(os.environ["SOROBN_B200_TILED"]))
        if os.environ.get("SOROBN_B200_GRAPH"):
            self.set_graph(int(os.environ["SOROBN_B200_GRAPH"]))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load().sbn_program_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ control
    def reserve(self, max_rows: int):
        _check(load().sbn_program_reserve(self._h, int(max_rows)))

    def set_graph(self, mode):
        """0 / False: plain launches; 1 / True: CUDA-graph replay (default); 3: graph with the independent
        sub-trees of the elimination as parallel branches."""
        _check(load().sbn_program_set_graph(self._h, int(mode)))

    def set_tiled(self, mode):
        """0: plain kernel; 1: tiled (default); 4: tiled, x-loop schedule only; 5: no slab variant;
        7: on-chip segments on (csrc/sbn_chain.h, opt-in); 6: off again; 9: tensor-map TMA pipeline
        kernel on for the steps it covers (csrc/sbn_tma.h, opt-in); 8: off again; 10: no paired steps; 11: paired
        steps where eligible (default; csrc/sbn_pair.h)."""
        _check(load().sbn_program_set_tiled(self._h, int(mode)))

    def info(self) -> dict:
        buf = (ctypes.c_int64 * 14)()
        _check(load().sbn_program_info(self._h, buf, 14))
        keys = ("Q", "n_ev", "n_steps", "scratch_floats_per_row", "reserved_rows", "launches", "mode",
                "shared_scratch_floats", "segments", "segment_steps", "segment_hbm_bytes_per_row",
                "segment_scratch_floats", "pairs", "pair_bytes_saved_per_row")
        return dict(zip(keys, [int(x) for x in buf]))

    # --------------------------------------------------------------------- runs
    def run(self, codes: np.ndarray, n_rows: int, out: np.ndarray | None = None) -> np.ndarray:
        """Host path: evidence codes uint8 [n_ev, n_rows] in, posterior float32
        [Q, n_rows] out (copies H2D, every step, D2H, synchronises)."""
        n_rows = int(n_rows)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if self.n_ev:
            if codes.shape != (self.n_ev, n_rows):
                raise ValueError(f"evidence codes have shape {codes.shape}, expected {(self.n_ev, n_rows)}")
        dtype = np.float64 if self.f64 else np.float32
        if out is None:
            out = np.empty((self.Q, n_rows), dtype=dtype)
        elif out.shape != (self.Q, n_rows) or out.dtype != dtype or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous {dtype.__name__} [Q, n_rows] array")
        ev_ptr = codes.ctypes.data if self.n_ev else None
        fn = load().sbn_program_run_host_f64 if self.f64 else load().sbn_program_run_host
        _check(fn(self._h, ev_ptr, n_rows, n_rows, out.ctypes.data, n_rows))
        return out

    def evidence(self, codes: np.ndarray, n_rows: int) -> np.ndarray:
        """P(event) per evidence row (the normaliser of the posterior), host path."""
        n_rows = int(n_rows)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if self.n_ev and codes.shape != (self.n_ev, n_rows):
            raise ValueError(f"evidence codes have shape {codes.shape}, expected {(self.n_ev, n_rows)}")
        out = np.empty(n_rows, dtype=np.float64 if self.f64 else np.float32)
        fn = load().sbn_program_evidence_host_f64 if self.f64 else load().sbn_program_evidence_host
        _check(fn(self._h, codes.ctypes.data if self.n_ev else None, n_rows, n_rows, out.ctypes.data))
        return out

    def run_device(self, d_ev: int, ld_ev: int, n_rows: int, d_out: int, ld_out: int, stream: int = 0):
        """Device path: raw device pointers, asynchronous on `stream`."""
        _check(load().sbn_program_run_device(self._h, ctypes.c_void_p(d_ev), int(ld_ev), int(n_rows),
                                             ctypes.c_void_p(d_out), int(ld_out), ctypes.c_void_p(stream)))

    def step_roles(self) -> np.ndarray:
        """Per program step: 0 ran once at creation, 1 own launch, 2 / 3 first / second step of a paired launch,
        4 / 5 expanding product / its consumer as one launch, 6 inside an on-chip segment."""
        roles = np.zeros(len(self.plan.steps), dtype=np.int32)
        _check(load().sbn_program_step_roles(self._h, roles.ctypes.data, roles.size))
        return roles

    def profile(self, d_ev: int, ld_ev: int, n_rows: int, d_out: int, ld_out: int, stream: int = 0) -> np.ndarray:
        n = len(self.plan.steps) + 1
        ms = np.zeros(n, dtype=np.float32)
        _check(load().sbn_program_profile(self._h, ctypes.c_void_p(d_ev), int(ld_ev), int(n_rows),
                                          ctypes.c_void_p(d_out), int(ld_out), ctypes.c_void_p(stream),
                                          ms.ctypes.data, n))
        return ms


class GibbsSampler:
    """Device Gibbs sampler for one (query variables, evidence variables) pair: one chain per
    evidence row (csrc/sbn_gibbs.cuh; the reference: bayes_net.py:665-737)."""

    def __init__(self, net, query_ids, evidence_ids, cycle_ids, device: int | None = None):
        lib = load()
        self.device = default_device() if device is None else int(device)
        n = len(net.names)
        card = np.ascontiguousarray(net.card, dtype=np.int32)
        self._card = [int(c) for c in card]
        par_ptr = np.zeros(n + 1, dtype=np.int32)
        par_idx = []
        offsets = np.zeros(n, dtype=np.int32)
        blob = []
        off = 0
        for v in range(n):
            par_idx.extend(net.parents[v])
            par_ptr[v + 1] = len(par_idx)
            t = np.ascontiguousarray(net.cpt[v], dtype=np.float32).reshape(-1)
            offsets[v] = off
            blob.append(t)
            off += t.size
        self._tables = np.concatenate(blob)
        par_idx = np.ascontiguousarray(par_idx, dtype=np.int32)
        self.query = np.ascontiguousarray(query_ids, dtype=np.int32)
        self.evidence = np.ascontiguousarray(evidence_ids, dtype=np.int32)
        # the cycle is only walked by Gibbs; keep the C side's "non-empty" contract when all is observed
        cycle = np.ascontiguousarray(cycle_ids if len(cycle_ids) else query_ids, dtype=np.int32)
        self.Q = int(np.prod([net.card[q] for q in query_ids]))
        self.n_ev = len(evidence_ids)
        self._h = ctypes.c_void_p()
        _check(lib.sbn_gibbs_create(
            self.device, n, card.ctypes.data, par_ptr.ctypes.data, par_idx.ctypes.data if par_idx.size else None,
            offsets.ctypes.data, self._tables.ctypes.data, self._tables.size, len(self.query), self.query.ctypes.data,
            self.n_ev, self.evidence.ctypes.data if self.n_ev else None, len(cycle), cycle.ctypes.data,
            ctypes.byref(self._h)))

    ALGORITHMS = {"gibbs": 0, "likelihood": 1, "rejection": 2, "gibbs_generic": 3}

    def run(self, codes: np.ndarray, n_chains: int, n_iterations: int, seed: int, algorithm: str = "gibbs") -> np.ndarray:
        """uint8 codes [n_ev, n_rows] -> estimated posterior float32 [Q, n_rows] (one Gibbs
        chain, or n_iterations forward samples, per evidence row)."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if self.n_ev and codes.shape != (self.n_ev, n_chains):
            raise ValueError(f"evidence codes have shape {codes.shape}, expected {(self.n_ev, n_chains)}")
        out = np.empty((self.Q, n_chains), dtype=np.float32)
        _check(load().sbn_sampler_run_host(self._h, self.ALGORITHMS[algorithm], codes.ctypes.data if self.n_ev else None,
                                           n_chains, n_chains, int(n_iterations),
                                           ctypes.c_uint64(int(seed) & (2**64 - 1)), out.ctypes.data, n_chains))
        return out

    def conditional(self, var: int, joint) -> np.ndarray:
        """P(var | Markov blanket) for one joint state (uint8 codes per variable id), as the chain
        evaluates it (bayes_net.py:699-712 precomputes the same table)."""
        joint = np.ascontiguousarray(joint, dtype=np.uint8)
        out = np.zeros(self._card[var], dtype=np.float32)
        _check(load().sbn_gibbs_conditional(self._h, int(var), joint.ctypes.data, out.ctypes.data))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            load().sbn_gibbs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
