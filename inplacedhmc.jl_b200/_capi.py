"""ctypes binding of the C ABI declared in include/bnuts.h.

The default library is the CUDA engine ``csrc/libbnuts.so``; loading fails loudly
if it has not been built (there is no CPU fallback in the product).  Tests pass an
explicit path to drive the CPU oracle, which exports the same ABI.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "csrc", "libbnuts.so")

F64, F32 = 0, 1
X_F64, X_F32, X_BF16 = 0, 1, 2
GRAD_AUTO, GRAD_DETERMINISTIC, GRAD_TENSOR = 0, 1, 2
METRIC_NONE, METRIC_DIAG = 0, 1

ERRORS = {
    -1: "invalid argument", -2: "no model", -3: "CUDA error", -4: "non-finite start",
    -5: "step size search failed", -6: "step size collapsed", -7: "unsupported", -8: "internal error",
    -9: "local optimum search failed",
}


class Config(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("dtype", C.c_int32), ("n_chains", C.c_int32), ("dim", C.c_int32),
        ("max_depth", C.c_int32), ("device", C.c_int32), ("min_delta", C.c_double), ("seed", C.c_uint64),
        ("chain_offset", C.c_int32), ("gradient_path", C.c_int32),
    ]


class DualAveragingParams(C.Structure):
    _fields_ = [("delta", C.c_double), ("gamma", C.c_double), ("kappa", C.c_double), ("t0", C.c_int32),
                ("_pad", C.c_int32)]


class StepsizeSearchParams(C.Structure):
    _fields_ = [("a_min", C.c_double), ("a_max", C.c_double), ("eps0", C.c_double), ("C", C.c_double),
                ("maxiter_crossing", C.c_int32), ("maxiter_bisect", C.c_int32)]


class CounterBlock(C.Structure):
    _fields_ = [("leapfrogs", C.c_int64), ("transitions", C.c_int64), ("lockstep_steps", C.c_int64),
                ("kernel_launches", C.c_int64), ("divergences", C.c_int64), ("gradient_rows", C.c_int64)]


# ≙ TreeStatisticsNUTS (src/NUTS.jl:229-242): 32-byte record
TREE_STATS_DTYPE = np.dtype([("pi", "<f8"), ("acceptance_rate", "<f8"), ("term_left", "<i4"),
                             ("term_right", "<i4"), ("depth", "<i4"), ("steps", "<i4")])
assert TREE_STATS_DTYPE.itemsize == 32

EXPORTS = [
    "bnuts_create", "bnuts_destroy", "bnuts_last_error", "bnuts_model_iid_normal", "bnuts_model_funnel",
    "bnuts_model_gaussian", "bnuts_model_logistic", "bnuts_model_logistic_synthetic", "bnuts_synth_logistic_rows",
    "bnuts_logistic_set_reference", "bnuts_set_positions", "bnuts_get_state",
    "bnuts_set_metric_diag", "bnuts_get_metric_diag", "bnuts_get_metric_diag_w", "bnuts_set_metric_diag_pair", "bnuts_set_metric_dense", "bnuts_get_metric_dense", "bnuts_set_stepsize", "bnuts_get_stepsize", "bnuts_seed", "bnuts_get_rng",
    "bnuts_inject", "bnuts_leapfrog", "bnuts_find_local_optimum", "bnuts_find_initial_stepsize", "bnuts_warmup_stage", "bnuts_sample",
    "bnuts_counters", "bnuts_profile", "bnuts_chain_status", "bnuts_set_allreduce", "bnuts_nccl_unique_id", "bnuts_set_nccl",
    "bnuts_p2p_export", "bnuts_p2p_connect",
]

_P = C.c_void_p
# ≙ bnuts_allreduce_fn: int32 fn(void* ctx, void* buf, int64 count, int32 dtype)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32)


class BnutsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bnuts error {code} ({ERRORS.get(code, '?')}): {msg}")
        self.code = code


def load_library(path=None):
    path = path or DEFAULT_LIB
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build the CUDA engine first (python __graft_entry__.py build). "
            "There is no CPU fallback.")
    lib = C.CDLL(path)
    lib.bnuts_last_error.restype = C.c_char_p
    lib.bnuts_last_error.argtypes = [_P]
    lib.bnuts_create.argtypes = [C.POINTER(Config), C.POINTER(_P)]
    lib.bnuts_destroy.argtypes = [_P]
    for n in ("bnuts_model_iid_normal", "bnuts_model_funnel"):
        getattr(lib, n).argtypes = [_P]
    lib.bnuts_model_gaussian.argtypes = [_P, _P]
    lib.bnuts_model_logistic.argtypes = [_P, _P, C.c_int32, _P, C.c_int64, C.c_double, C.c_int32]
    lib.bnuts_logistic_set_reference.argtypes = [_P, _P]
    lib.bnuts_model_logistic_synthetic.argtypes = [_P, C.c_uint64, C.c_int64, C.c_int64, C.c_double, C.c_int32]
    lib.bnuts_synth_logistic_rows.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int32, _P, _P, _P]
    lib.bnuts_set_allreduce.argtypes = [_P, ALLREDUCE_FN, _P]
    lib.bnuts_nccl_unique_id.argtypes = [_P]
    lib.bnuts_set_nccl.argtypes = [_P, _P, C.c_int32, C.c_int32]
    lib.bnuts_p2p_export.argtypes = [_P, _P]
    lib.bnuts_p2p_connect.argtypes = [_P, _P, C.c_int32, C.c_int32]
    lib.bnuts_set_positions.argtypes = [_P, _P]
    lib.bnuts_get_state.argtypes = [_P, _P, _P, _P]
    lib.bnuts_set_metric_diag.argtypes = [_P, _P]
    lib.bnuts_get_metric_diag.argtypes = [_P, _P]
    lib.bnuts_get_metric_diag_w.argtypes = [_P, _P]
    lib.bnuts_set_metric_diag_pair.argtypes = [_P, _P, _P]
    lib.bnuts_set_metric_dense.argtypes = [_P, _P]
    lib.bnuts_get_metric_dense.argtypes = [_P, _P]
    lib.bnuts_set_stepsize.argtypes = [_P, _P]
    lib.bnuts_get_stepsize.argtypes = [_P, _P]
    lib.bnuts_seed.argtypes = [_P, C.c_uint64, C.c_uint32]
    lib.bnuts_get_rng.argtypes = [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    lib.bnuts_inject.argtypes = [_P, C.c_int32, _P, _P, _P, C.c_int32]
    lib.bnuts_leapfrog.argtypes = [_P, _P, _P, C.c_int32, _P, _P, _P, _P]
    lib.bnuts_find_local_optimum.argtypes = [_P, C.c_double, C.c_int32]
    lib.bnuts_find_initial_stepsize.argtypes = [_P, C.POINTER(StepsizeSearchParams)]
    lib.bnuts_warmup_stage.argtypes = [_P, C.c_int32, C.c_int32, C.POINTER(DualAveragingParams), C.c_double,
                                       _P, C.c_int64, C.c_int64, _P, C.c_int64, _P]
    lib.bnuts_sample.argtypes = [_P, C.c_int32, _P, C.c_int64, C.c_int64, _P, C.c_int64, _P]
    lib.bnuts_counters.argtypes = [_P, C.POINTER(CounterBlock)]
    lib.bnuts_chain_status.argtypes = [_P, _P]
    lib.bnuts_profile.argtypes = [_P, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    for n in EXPORTS:
        if n != "bnuts_last_error":
            getattr(lib, n).restype = C.c_int32
    return lib


def nccl_unique_id(lib=None):
    lib = lib or load_library()
    buf = (C.c_uint8 * 128)()
    rc = lib.bnuts_nccl_unique_id(buf)
    if rc:
        raise BnutsError(rc, "bnuts_nccl_unique_id failed (NCCL not found?)")
    return bytes(buf)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


def synth_logistic_rows(data_seed, row_offset, nrows, dim, lib=None):
    """Host copy of rows [row_offset, row_offset + nrows) of the synthetic design matrix (include/bnuts.h):
    X as bf16 bits (uint16), y, beta_true."""
    lib = lib or load_library()
    X = np.empty((nrows, dim), dtype=np.uint16); y = np.empty(nrows); beta = np.empty(dim)
    rc = lib.bnuts_synth_logistic_rows(data_seed, row_offset, nrows, dim, _ptr(X), _ptr(y), _ptr(beta))
    if rc:
        raise BnutsError(rc, "bnuts_synth_logistic_rows failed")
    return X, y, beta


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


class Engine:
    """One engine = the chains resident on one GPU (or, for the oracle, one process)."""

    def __init__(self, n_chains, dim, dtype=F64, max_depth=10, min_delta=-1000.0, seed=20261018, chain_offset=0,
                 device=0, gradient_path=GRAD_AUTO, lib=None):
        self.lib = lib if lib is not None and not isinstance(lib, str) else load_library(lib)
        self.C, self.D, self.dtype = int(n_chains), int(dim), dtype
        cfg = Config(C.sizeof(Config), dtype, n_chains, dim, max_depth, device, min_delta, seed, chain_offset,
                     gradient_path)
        h = _P()
        rc = self.lib.bnuts_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise BnutsError(rc, (self.lib.bnuts_last_error(None) or b"").decode())
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.bnuts_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, allow=()):
        if rc != 0 and rc not in allow:
            raise BnutsError(rc, (self.lib.bnuts_last_error(self.h) or b"").decode())
        return rc

    # ---- models
    def model_iid_normal(self):
        self._chk(self.lib.bnuts_model_iid_normal(self.h))

    def model_funnel(self):
        self._chk(self.lib.bnuts_model_funnel(self.h))

    def model_gaussian(self, precision):
        P = _f64(precision, (self.D, self.D))
        self._chk(self.lib.bnuts_model_gaussian(self.h, _ptr(P)))

    def model_logistic(self, X, y, prior_precision=1.0, row_blocks=1, x_dtype=None):
        X = np.ascontiguousarray(X)
        if x_dtype is None:
            x_dtype = {np.dtype("float64"): X_F64, np.dtype("float32"): X_F32, np.dtype("uint16"): X_BF16}[X.dtype]
        assert X.ndim == 2 and X.shape[1] == self.D
        y = _f64(y, (X.shape[0],))
        self._chk(self.lib.bnuts_model_logistic(self.h, _ptr(X), x_dtype, _ptr(y), X.shape[0], prior_precision,
                                                row_blocks))

    def model_logistic_synthetic(self, data_seed, row_offset, n_rows, prior_precision=1.0, row_blocks=1):
        """Logistic model on the synthetic rows [row_offset, row_offset + n_rows) (generated on the device for the
        tensor path)."""
        self._chk(self.lib.bnuts_model_logistic_synthetic(self.h, data_seed, row_offset, n_rows, prior_precision, row_blocks))

    def logistic_set_reference(self, beta_ref=None):
        """Tensor-core logistic path: evaluate X·β as X·β_ref + X·(β − β_ref) (two bf16 terms instead of
        three).  β_ref must sit near the posterior mode (checked); None returns to the exact path."""
        b = _f64(beta_ref, (self.D,))
        self._chk(self.lib.bnuts_logistic_set_reference(self.h, _ptr(b)))

    # ---- row-sharded data (every engine of the group holds all chains and one shard of the rows of X)
    def set_allreduce(self, fn):
        """fn(address, count, dtype) sums `count` elements (dtype 0 = Float64, 1 = Float32) at `address` in place
        across the group.  The ctypes thunk is kept alive on the engine."""
        def thunk(_ctx, buf, count, dtype):
            try:
                fn(buf, count, dtype)
                return 0
            except Exception:          # never let an exception cross the C ABI
                import traceback
                traceback.print_exc()
                return -1
        self._allreduce_thunk = ALLREDUCE_FN(thunk)
        self._chk(self.lib.bnuts_set_allreduce(self.h, self._allreduce_thunk, None))

    def set_nccl(self, unique_id, world, rank):
        """unique_id: the 128 bytes of nccl_unique_id(lib) from rank 0 (broadcast them with the host's own plumbing)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._chk(self.lib.bnuts_set_nccl(self.h, buf, world, rank))

    def p2p_export(self):
        """64-byte IPC handle of this engine's receive buffer (peer-memory exchange over NVLink)."""
        buf = (C.c_uint8 * 64)()
        self._chk(self.lib.bnuts_p2p_export(self.h, buf))
        return bytes(buf)

    def p2p_connect(self, handles, rank):
        """handles: the p2p_export() bytes of every rank, in rank order."""
        blob = b"".join(bytes(h) for h in handles)
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._chk(self.lib.bnuts_p2p_connect(self.h, buf, len(handles), rank))

    # ---- state
    def set_positions(self, q=None, allow_nonfinite=False):
        q = _f64(q, (self.C, self.D))
        return self._chk(self.lib.bnuts_set_positions(self.h, _ptr(q)), allow=(-4,) if allow_nonfinite else ())

    def get_state(self):
        q = np.empty((self.C, self.D)); g = np.empty((self.C, self.D)); l = np.empty(self.C)
        self._chk(self.lib.bnuts_get_state(self.h, _ptr(q), _ptr(g), _ptr(l)))
        return q, g, l

    def set_metric_diag(self, minv=None):
        minv = _f64(minv, (self.C, self.D))
        self._chk(self.lib.bnuts_set_metric_diag(self.h, _ptr(minv)))

    def get_metric_diag(self):
        m = np.empty((self.C, self.D))
        self._chk(self.lib.bnuts_get_metric_diag(self.h, _ptr(m)))
        return m

    def get_metric_diag_w(self):
        w = np.empty((self.C, self.D))
        self._chk(self.lib.bnuts_get_metric_diag_w(self.h, _ptr(w)))
        return w

    def set_metric_diag_pair(self, minv, w):
        m = _f64(minv, (self.C, self.D)); w = _f64(w, (self.C, self.D))
        self._chk(self.lib.bnuts_set_metric_diag_pair(self.h, _ptr(m), _ptr(w)))

    def set_metric_dense(self, minv=None):
        """One dense SPD M⁻¹ [D, D] shared by all chains (None: back to the per-chain diagonal metric)."""
        m = _f64(minv, (self.D, self.D))
        self._chk(self.lib.bnuts_set_metric_dense(self.h, _ptr(m)))

    def get_metric_dense(self):
        m = np.empty((self.D, self.D))
        self._chk(self.lib.bnuts_get_metric_dense(self.h, _ptr(m)))
        return m

    def set_stepsize(self, eps):
        eps = _f64(np.broadcast_to(np.asarray(eps, dtype=np.float64), (self.C,)))
        self._chk(self.lib.bnuts_set_stepsize(self.h, _ptr(eps)))

    def get_stepsize(self):
        e = np.empty(self.C)
        self._chk(self.lib.bnuts_get_stepsize(self.h, _ptr(e)))
        return e

    def seed(self, seed, next_transition=0):
        self._chk(self.lib.bnuts_seed(self.h, seed, next_transition))

    def rng(self):
        """(seed, next transition index): the position of the counter-based generator."""
        s, t = C.c_uint64(0), C.c_uint32(0)
        self._chk(self.lib.bnuts_get_rng(self.h, C.byref(s), C.byref(t)))
        return int(s.value), int(t.value)

    # ---- ≙ WarmupState (z, κ, ϵ), src/warmup.jl:47-51, plus the generator position: checkpoint / resume
    def warmup_state(self):
        seed, t = self.rng()
        return {"q": self.get_state()[0], "κ": self.get_metric_diag(), "W": self.get_metric_diag_w(), "ϵ": self.get_stepsize(),
                "seed": seed, "next_transition": t}

    def restore(self, state):
        """Continue from a `warmup_state()` of another engine with the same model, dtype and chain ids: the continued
        run is bit-identical to an uninterrupted one (ℓ and ∇ℓ are re-evaluated at q, deterministically)."""
        if state.get("W") is not None:
            self.set_metric_diag_pair(state["κ"], state["W"])      # ≙ GaussianKineticEnergy(M⁻¹, W): both fields
        else:
            self.set_metric_diag(state["κ"])
        self.set_stepsize(state["ϵ"])
        self.seed(state["seed"], state["next_transition"])
        self.set_positions(state["q"])

    def inject(self, T, dirs=None, p=None, exps=None):
        """≙ sample_tree(rng, ...; p, directions) with a scripted rng (src/NUTS.jl:251-258, :32-34): directions [T, C],
        momenta [T, C, D], merge exponentials [T, C, n] in the order the reference would consume them."""
        if dirs is not None:
            dirs = np.ascontiguousarray(dirs, dtype=np.uint32)
            assert dirs.shape == (T, self.C)
        p = _f64(p, (T, self.C, self.D))
        n_exps = 0
        if exps is not None:
            exps = np.ascontiguousarray(exps, dtype=np.float64)
            assert exps.ndim == 3 and exps.shape[:2] == (T, self.C)
            n_exps = exps.shape[2]
        self._chk(self.lib.bnuts_inject(self.h, T, _ptr(dirs), _ptr(p), _ptr(exps), n_exps))

    def leapfrog(self, p, eps, nsteps=1):
        p = _f64(p, (self.C, self.D))
        eps = _f64(np.broadcast_to(np.asarray(eps, dtype=np.float64), (self.C,)))
        q = np.empty((self.C, self.D)); po = np.empty((self.C, self.D)); g = np.empty((self.C, self.D))
        l = np.empty(self.C)
        self._chk(self.lib.bnuts_leapfrog(self.h, _ptr(p), _ptr(eps), nsteps, _ptr(q), _ptr(po), _ptr(g), _ptr(l)))
        return q, po, g, l

    def find_local_optimum(self, magnitude_penalty=1e-4, iterations=50):
        """≙ warmup!(FindLocalOptimum), src/warmup.jl:152-186 (defaults :143,148)."""
        self._chk(self.lib.bnuts_find_local_optimum(self.h, magnitude_penalty, iterations))

    def find_initial_stepsize(self, a_min=0.25, a_max=0.75, eps0=1.0, C_=2.0, maxiter_crossing=400,
                              maxiter_bisect=400, allow_fail=False):
        P = StepsizeSearchParams(a_min, a_max, eps0, C_, maxiter_crossing, maxiter_bisect)
        return self._chk(self.lib.bnuts_find_initial_stepsize(self.h, C.byref(P)),
                         allow=(-4, -5) if allow_fail else ())

    def warmup_stage(self, N, metric_kind=METRIC_NONE, delta=0.8, gamma=0.05, kappa=0.75, t0=10, lam=-1.0,
                     keep=True, allow_fail=False, fixed_stepsize=False):
        """≙ warmup!(…, TuningNUTS{M}, …), src/warmup.jl:269-314.  fixed_stepsize=True ≙ FixedStepsize (src/stepsize.jl:251-255):
        ϵ is kept, only the metric is tuned."""
        da = DualAveragingParams(delta, gamma, kappa, t0, 0)
        chain = np.empty((self.C, N, self.D)) if keep else None
        stats = np.zeros((self.C, N), dtype=TREE_STATS_DTYPE) if keep else None
        eps = np.empty((self.C, N)) if keep else None
        self._chk(self.lib.bnuts_warmup_stage(self.h, N, metric_kind, None if fixed_stepsize else C.byref(da), lam, _ptr(chain), self.D,
                                              N * self.D, _ptr(stats), N, _ptr(eps)),
                  allow=(-6,) if allow_fail else ())
        return chain, stats, eps

    def sample_device_only(self, N):
        """N transitions with no host output (draws/statistics stay on the device)."""
        self._chk(self.lib.bnuts_sample(self.h, N, None, self.D, N * self.D, None, N, None))

    def sample(self, N, want_index=False, out=None):
        if out is None:
            chain = np.empty((self.C, N, self.D)); stats = np.zeros((self.C, N), dtype=TREE_STATS_DTYPE)
        else:
            chain, stats = out
        sel = np.zeros((self.C, N), dtype=np.int32) if want_index else None
        self._chk(self.lib.bnuts_sample(self.h, N, _ptr(chain), self.D, N * self.D, _ptr(stats), N, _ptr(sel)))
        return (chain, stats, sel) if want_index else (chain, stats)

    def counters(self):
        cb = CounterBlock()
        self._chk(self.lib.bnuts_counters(self.h, C.byref(cb)))
        return {k: getattr(cb, k) for k, _ in CounterBlock._fields_}

    def profile(self, enable=True):
        """Returns (gradient_ms, gradient_launches) accumulated since the last call, then (re)arms."""
        ms = C.c_double(); n = C.c_int64()
        self._chk(self.lib.bnuts_profile(self.h, 1 if enable else 0, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def chain_status(self):
        s = np.zeros(self.C, dtype=np.int32)
        self._chk(self.lib.bnuts_chain_status(self.h, _ptr(s)))
        return s
