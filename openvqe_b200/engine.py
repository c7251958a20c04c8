"""Thin object wrapper over the C ABI (include/vqe_b200.h).

One ``Engine`` = one complex128 state vector resident in HBM plus two scratch
buffers (sigma, work) allocated on first use.  All arithmetic happens in the
CUDA library; this module only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from .lowering import PackedTerms, pack_operator, pack_pool

BUF_PSI, BUF_SIGMA, BUF_WORK, BUF_AUX = 0, 1, 2, 3
GATE_X, GATE_H, GATE_RX, GATE_RY, GATE_RZ, GATE_CNOT = range(6)
GATE_KINDS = {"X": GATE_X, "H": GATE_H, "RX": GATE_RX, "RY": GATE_RY, "RZ": GATE_RZ, "CNOT": GATE_CNOT}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def operator_fingerprint(op):
    """Cheap content mark of a duck-typed operator for the identity-keyed caches: term-list identity and length, first
    and last term.  O(1); catches a replaced or truncated/extended term list and edited end terms."""
    terms = getattr(op, "terms", None)
    if terms is None:
        return None
    n = len(terms)
    if n == 0:
        return (id(terms), 0)
    a, b = terms[0], terms[-1]
    return (id(terms), n, complex(a.coeff), a.op, complex(b.coeff), b.op, getattr(op, "constant_coeff", None))


class PauliSum:
    """Device-resident, X-mask-grouped Pauli sum (Hamiltonian or observable)."""

    def __init__(self, engine, packed: PackedTerms):
        lib = _lib.load()
        self._lib = lib
        self._engine = engine
        self.n_terms = len(packed)
        self.handle = C.c_void_p()
        _lib.check(lib.vqe_paulisum_create(engine.handle, C.byref(self.handle), self.n_terms, _ptr(packed.x),
                                           _ptr(packed.z), _ptr(packed.ny), _ptr(packed.cre), _ptr(packed.cim)))
        self.n_groups = lib.vqe_paulisum_groups(self.handle)
        self.n_passes = lib.vqe_paulisum_passes(self.handle)
        self._finalizer = weakref.finalize(self, lib.vqe_paulisum_destroy, self.handle)


class Engine:
    def __init__(self, n_qubits: int, device: int = 0, n_global: int = 0, rank: int = 0):
        """``n_global`` > 0 creates ONE RANK of a sharded state (see ``openvqe_b200.sharded``): this context
        then holds 2^(n_qubits - n_global) amplitudes and its reductions are per-rank partial sums."""
        lib = _lib.load()
        self._lib = lib
        self.n = int(n_qubits)
        self.device = int(device)
        self.n_global = int(n_global)
        self.rank = int(rank)
        self.n_local = self.n - self.n_global
        self.handle = C.c_void_p()
        if self.n_global:
            _lib.check(lib.vqe_create_shard(C.byref(self.handle), self.n, self.n_global, self.rank, self.device))
        else:
            _lib.check(lib.vqe_create(C.byref(self.handle), self.n, self.device))
        self._finalizer = weakref.finalize(self, lib.vqe_destroy, self.handle)
        self._ps_cache = {}

    # -- state ---------------------------------------------------------------
    def set_basis_state(self, index: int):
        _lib.check(self._lib.vqe_set_basis_state(self.handle, int(index)))

    def set_state(self, vec, buf=BUF_PSI):
        v = np.ascontiguousarray(np.asarray(vec, dtype=np.complex128).reshape(-1))
        if v.shape[0] != 1 << self.n_local:
            raise ValueError("state has %d amplitudes, expected 2^%d" % (v.shape[0], self.n_local))
        _lib.check(self._lib.vqe_set_state(self.handle, buf, _ptr(v)))

    def get_state(self, buf=BUF_PSI):
        out = np.empty(1 << self.n_local, dtype=np.complex128)
        _lib.check(self._lib.vqe_get_state(self.handle, buf, _ptr(out)))
        return out

    def copy_buffer(self, dst, src):
        _lib.check(self._lib.vqe_copy_buffer(self.handle, dst, src))

    # -- state preparation ---------------------------------------------------
    def apply_rotations(self, x, z, ny, angles, buf=BUF_PSI):
        """buf <- prod_k exp(-i angles[k] P_k) buf, k = 0 first (buf = the state unless told otherwise)."""
        x = np.ascontiguousarray(x, dtype=np.uint64)
        z = np.ascontiguousarray(z, dtype=np.uint64)
        ny = np.ascontiguousarray(ny, dtype=np.int32)
        a = np.ascontiguousarray(angles, dtype=np.float64)
        if not (x.shape == z.shape == ny.shape == a.shape):
            raise ValueError("rotation arrays differ in length")
        if buf == BUF_PSI:
            _lib.check(self._lib.vqe_apply_pauli_rotations(self.handle, int(x.shape[0]), _ptr(x), _ptr(z), _ptr(ny), _ptr(a)))
        else:
            _lib.check(self._lib.vqe_apply_pauli_rotations_buf(self.handle, int(buf), int(x.shape[0]), _ptr(x), _ptr(z),
                                                               _ptr(ny), _ptr(a)))

    def apply_gates(self, kinds, q0, q1, angles):
        k = np.ascontiguousarray(kinds, dtype=np.int32)
        a0 = np.ascontiguousarray(q0, dtype=np.int32)
        a1 = np.ascontiguousarray(q1, dtype=np.int32)
        an = np.ascontiguousarray(angles, dtype=np.float64)
        _lib.check(self._lib.vqe_apply_gates(self.handle, int(k.shape[0]), _ptr(k), _ptr(a0), _ptr(a1), _ptr(an)))

    def apply_plane_rotations(self, xmask, offsets, pattern, cosv, sinv):
        """Tabulated plane rotations (see include/vqe_b200.h): op k couples (l, l ^ xmask[k]); for every listed
        a-side pattern the pairs rotate by (cos, sin)."""
        x = np.ascontiguousarray(xmask, dtype=np.uint64)
        o = np.ascontiguousarray(offsets, dtype=np.int32)
        p = np.ascontiguousarray(pattern, dtype=np.uint64)
        c = np.ascontiguousarray(cosv, dtype=np.float64)
        s = np.ascontiguousarray(sinv, dtype=np.float64)
        _lib.check(self._lib.vqe_apply_plane_rotations(self.handle, int(x.shape[0]), _ptr(x), _ptr(o), _ptr(p), _ptr(c), _ptr(s)))

    def scale_state(self, factor: complex, buf=BUF_PSI):
        f = complex(factor)
        _lib.check(self._lib.vqe_scale_state(self.handle, buf, f.real, f.imag))

    def axpby(self, dst, x, alpha: complex = 1.0, beta: complex = 1.0):
        """buffer dst <- alpha * buffer x + beta * buffer dst (rank-local)."""
        a, b = complex(alpha), complex(beta)
        _lib.check(self._lib.vqe_axpby(self.handle, int(dst), int(x), a.real, a.imag, b.real, b.imag))

    def apply_exp(self, packed: PackedTerms, theta: float):
        """psi <- exp(theta * A) psi (exact exponential of the whole generator)."""
        _lib.check(self._lib.vqe_apply_exp_paulisum(self.handle, len(packed), _ptr(packed.x), _ptr(packed.z),
                                                    _ptr(packed.ny), _ptr(packed.cre), _ptr(packed.cim), float(theta)))

    # -- observables ---------------------------------------------------------
    def paulisum(self, operator) -> PauliSum:
        """Lower + upload ``operator`` once; cached per object identity plus a cheap fingerprint of its term list
        (operators are treated as immutable, like the reference treats them; replacing the term list or editing its
        ends is noticed, an in-place edit in the middle is not)."""
        key = id(operator)
        mark = operator_fingerprint(operator)
        hit = self._ps_cache.get(key)
        if hit is not None and hit[0]() is operator and hit[2] == mark:
            return hit[1]
        packed = operator if isinstance(operator, PackedTerms) else pack_operator(operator, with_constant=True)
        ps = PauliSum(self, packed)
        try:
            self._ps_cache[key] = (weakref.ref(operator), ps, mark)
        except TypeError:
            pass
        if len(self._ps_cache) > 64:
            self._ps_cache.pop(next(iter(self._ps_cache)))
        return ps

    def expectation(self, ps: PauliSum, buf=BUF_PSI) -> complex:
        out = (C.c_double * 2)()
        _lib.check(self._lib.vqe_expectation(self.handle, buf, ps.handle, out))
        return complex(out[0], out[1])

    def apply_paulisum(self, ps: PauliSum, dst=BUF_SIGMA, src=BUF_PSI):
        _lib.check(self._lib.vqe_apply_paulisum(self.handle, dst, src, ps.handle))

    def pool_overlaps(self, pool: PackedTerms, bra=BUF_SIGMA, ket=BUF_PSI):
        """-> complex array, out[k] = <bra| A_k |ket> for every pool operator."""
        n_ops = int(pool.offsets.shape[0]) - 1
        out = np.zeros(n_ops, dtype=np.complex128)
        _lib.check(self._lib.vqe_pool_overlaps(self.handle, bra, ket, n_ops, _ptr(pool.offsets), _ptr(pool.x),
                                               _ptr(pool.z), _ptr(pool.ny), _ptr(pool.cre), _ptr(pool.cim), _ptr(out)))
        return out

    # -- reductions ----------------------------------------------------------
    def norm2(self, buf=BUF_PSI) -> float:
        out = C.c_double()
        _lib.check(self._lib.vqe_norm2(self.handle, buf, C.byref(out)))
        return out.value

    def inner(self, a, b) -> complex:
        out = (C.c_double * 2)()
        _lib.check(self._lib.vqe_inner(self.handle, a, b, out))
        return complex(out[0], out[1])

    def overlap_host(self, vec, buf=BUF_PSI) -> complex:
        v = np.ascontiguousarray(np.asarray(vec, dtype=np.complex128).reshape(-1))
        out = (C.c_double * 2)()
        _lib.check(self._lib.vqe_overlap_host(self.handle, buf, _ptr(v), out))
        return complex(out[0], out[1])

    # -- bookkeeping ---------------------------------------------------------
    @property
    def real_layout(self) -> bool:
        """True while the state buffer is kept as 2^n doubles (purely real state; see include/vqe_b200.h)."""
        return bool(self._lib.vqe_state_layout(self.handle))

    def synchronize(self):
        _lib.check(self._lib.vqe_synchronize(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self._lib.vqe_launch_count(self.handle))

    def profile(self, on: bool):
        _lib.check(self._lib.vqe_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self, which: int, reset=False):
        ms, n = C.c_double(), C.c_uint64()
        _lib.check(self._lib.vqe_profile_read(self.handle, which, C.byref(ms), C.byref(n), 1 if reset else 0))
        return ms.value, n.value

    def timer_begin(self):
        _lib.check(self._lib.vqe_timer_begin(self.handle))

    def timer_end(self) -> float:
        ms = C.c_double()
        _lib.check(self._lib.vqe_timer_end(self.handle, C.byref(ms)))
        return ms.value

    def transfer_bytes(self, reset=False):
        a, b = C.c_uint64(), C.c_uint64()
        _lib.check(self._lib.vqe_transfer_bytes(self.handle, C.byref(a), C.byref(b), 1 if reset else 0))
        return a.value, b.value

    def gather_bytes(self, reset=False) -> int:
        """Bytes this rank has fetched from partner shards in gather-form peer passes (sharded states)."""
        g = C.c_uint64()
        _lib.check(self._lib.vqe_peer_bytes(self.handle, C.byref(g), 1 if reset else 0))
        return g.value

    def relabel_stats(self, reset=False):
        """(qubit swaps executed, bytes this rank read from partner shards in them) -- sharded states, see include/vqe_b200.h."""
        a, b = C.c_uint64(), C.c_uint64()
        _lib.check(self._lib.vqe_relabel_stats(self.handle, C.byref(a), C.byref(b), 1 if reset else 0))
        return a.value, b.value

    def buffer_ptr(self, buf=BUF_PSI):
        p, n = C.c_void_p(), C.c_uint64()
        _lib.check(self._lib.vqe_buffer_ptr(self.handle, buf, C.byref(p), C.byref(n)))
        return p.value, n.value


_ENGINES = {}


_ENGINE_FACTORY = None  # set by openvqe_b200.sharded.enable(): (n_qubits, device) -> Engine or None


def default_device() -> int:
    """Device of this process: VQE_B200_DEVICE, else LOCAL_RANK (one process per GPU under torchrun), else 0."""
    import os
    v = os.environ.get("VQE_B200_DEVICE", os.environ.get("LOCAL_RANK", "0"))
    try:
        dev = int(v)
    except ValueError:
        dev = 0
    n = _lib.load().vqe_device_count()
    return dev % n if n > 0 else dev


def get_engine(n_qubits: int, device=None) -> Engine:
    """Process-wide engine per (n_qubits, device): the state buffers are reused
    across the thousands of objective evaluations of one optimisation."""
    key = (int(n_qubits), default_device() if device is None else int(device))
    eng = _ENGINES.get(key)
    if eng is None:
        if _ENGINE_FACTORY is not None:
            eng = _ENGINE_FACTORY(*key)
        if eng is None:
            eng = Engine(*key)
        _ENGINES[key] = eng
    return eng


def release_engines():
    _ENGINES.clear()
