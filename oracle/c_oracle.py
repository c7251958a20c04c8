"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of the plain-C oracle (oracle/c/vqe_oracle.c).

Used as the CPU baseline of bench.py (``cpu_baseline`` leg and ``--impl reference``) and as a second,
independent checker in tests.  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "c", "libvqe_oracle.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "c")], check=True)
    lib = C.CDLL(LIB_PATH)
    vp, u64, i32, dbl = C.c_void_p, C.c_uint64, C.c_int32, C.c_double
    lib.orc_threads.restype = C.c_int
    lib.orc_set_threads.restype = C.c_int
    lib.orc_set_threads.argtypes = [C.c_int]
    lib.orc_basis_state.argtypes = [vp, C.c_int, u64]
    lib.orc_pauli_rotation.argtypes = [vp, C.c_int, u64, u64, C.c_int, dbl]
    lib.orc_pauli_expectation.argtypes = [vp, C.c_int, u64, u64, C.c_int, vp]
    lib.orc_expectation.restype = dbl
    lib.orc_expectation.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    lib.orc_ucc_energy.restype = dbl
    lib.orc_ucc_energy.argtypes = [vp, C.c_int, u64, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp]
    lib.orc_apply_gates.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp]
    lib.orc_apply_paulisum.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    lib.orc_pool_overlaps.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def threads():
    return int(load().orc_threads())


def set_threads(n_threads=None):
    """Use n_threads OpenMP threads (default: every host core this process may run on), whatever OMP_NUM_THREADS
    says -- torchrun pins it to 1 for every rank.  Returns the thread count now in effect."""
    if n_threads is None:
        try:
            n_threads = len(os.sched_getaffinity(0))
        except AttributeError:
            n_threads = os.cpu_count() or 1
    return int(load().orc_set_threads(int(n_threads)))


def _arrs(x, z, ny, *rest):
    out = [np.ascontiguousarray(x, dtype=np.uint64), np.ascontiguousarray(z, dtype=np.uint64),
           np.ascontiguousarray(ny, dtype=np.int32)]
    out += [np.ascontiguousarray(r, dtype=np.float64) for r in rest]
    return out


def apply_rotations(psi, n, x, z, ny, angles):
    lib = load()
    x, z, ny, a = _arrs(x, z, ny, angles)
    for k in range(len(a)):
        if a[k] != 0.0:
            lib.orc_pauli_rotation(_p(psi), n, int(x[k]), int(z[k]), int(ny[k]), float(a[k]))
    return psi


def expectation(psi, n, x, z, ny, cre, cim):
    x, z, ny, cre, cim = _arrs(x, z, ny, cre, cim)
    return float(load().orc_expectation(_p(psi), n, len(x), _p(x), _p(z), _p(ny), _p(cre), _p(cim)))


def ucc_energy(n, hf_index, rot, angles, ham, psi=None):
    """rot / ham: objects with .x .z .ny (.cre .cim) arrays (openvqe_b200.lowering.PackedTerms layout)."""
    if psi is None:
        psi = np.empty(1 << n, dtype=np.complex128)
    a = np.ascontiguousarray(angles, dtype=np.float64)
    return float(load().orc_ucc_energy(_p(psi), n, int(hf_index), len(a), _p(rot.x), _p(rot.z), _p(rot.ny), _p(a),
                                       len(ham.x), _p(ham.x), _p(ham.z), _p(ham.ny), _p(ham.cre), _p(ham.cim)))


def apply_gates(psi, n, kinds, q0, q1, angles):
    k = np.ascontiguousarray(kinds, dtype=np.int32)
    a0 = np.ascontiguousarray(q0, dtype=np.int32)
    a1 = np.ascontiguousarray(q1, dtype=np.int32)
    an = np.ascontiguousarray(angles, dtype=np.float64)
    load().orc_apply_gates(_p(psi), n, len(k), _p(k), _p(a0), _p(a1), _p(an))
    return psi


def apply_paulisum(psi, n, ham):
    out = np.empty_like(psi)
    load().orc_apply_paulisum(_p(psi), _p(out), n, len(ham.x), _p(ham.x), _p(ham.z), _p(ham.ny), _p(ham.cre), _p(ham.cim))
    return out


def pool_overlaps(bra, ket, n, pool):
    n_ops = len(pool.offsets) - 1
    out = np.zeros(n_ops, dtype=np.complex128)
    work = np.empty_like(ket)
    load().orc_pool_overlaps(_p(bra), _p(ket), _p(work), n, n_ops, _p(pool.offsets), _p(pool.x), _p(pool.z),
                             _p(pool.ny), _p(pool.cre), _p(pool.cim), _p(out))
    return out
