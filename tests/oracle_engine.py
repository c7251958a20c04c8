"""TEST DOUBLE -- an Engine look-alike whose arithmetic is the numpy oracle.

Lets the CPU test-suite (no GPU in the build container) drive the HOST logic of the drop-in -- the reference-shaped
modules under openvqe_b200/ucc_family and openvqe_b200/adapt, their caches and bookkeeping -- end to end, e.g. under the
reference's own facade (tests/test_facade_cpu.py).  Installed through ``openvqe_b200.engine._ENGINE_FACTORY``; it is never
importable from the product package, and the product path still fails loudly without the CUDA library."""
import numpy as np
import scipy.sparse
import scipy.sparse.linalg

from oracle import statevector_oracle as orc
from openvqe_b200.engine import GATE_KINDS
from openvqe_b200.lowering import PackedTerms, pack_operator

_GATE_NAMES = {v: k for k, v in GATE_KINDS.items()}


class _PS:
    def __init__(self, packed):
        self.packed = packed
        self.n_terms = len(packed)
        self.n_groups = len(set(packed.x.tolist()))
        self.n_passes = 1


class OracleEngine:
    n_global = 0
    rank = 0

    def __init__(self, n_qubits, device=0):
        self.n = self.n_local = int(n_qubits)
        self.device = device
        self.buf = {k: np.zeros(1 << self.n, dtype=np.complex128) for k in range(4)}
        self.launch_count = 0

    # -- state
    def set_basis_state(self, index):
        self.buf[0] = orc.basis_state(self.n, int(index))

    def set_state(self, vec, buf=0):
        self.buf[buf] = np.asarray(vec, dtype=np.complex128).reshape(-1).copy()

    def get_state(self, buf=0):
        return self.buf[buf].copy()

    def copy_buffer(self, dst, src):
        self.buf[dst] = self.buf[src].copy()

    def scale_state(self, factor, buf=0):
        self.buf[buf] = self.buf[buf] * complex(factor)

    def axpby(self, dst, x, alpha=1.0, beta=1.0):
        self.buf[dst] = complex(alpha) * self.buf[x] + complex(beta) * self.buf[dst]

    # -- state preparation
    def apply_rotations(self, x, z, ny, angles, buf=0):
        psi = self.buf[buf]
        for xk, zk, nk, a in zip(x, z, ny, angles):
            if a != 0.0:
                psi = orc.pauli_rotation(psi, int(xk), int(zk), int(nk), float(a))
        self.buf[buf] = psi

    def apply_gates(self, kinds, q0, q1, angles):
        gates = []
        for k, a, b, t in zip(kinds, q0, q1, angles):
            name = _GATE_NAMES[int(k)]
            gates.append((name, [int(a), int(b)] if name == "CNOT" else [int(a)], None if name in ("X", "H", "CNOT") else float(t)))
        self.buf[0] = orc.apply_gates(self.buf[0], self.n, gates)

    def apply_plane_rotations(self, xmask, offsets, pattern, cosv, sinv):
        psi = self.buf[0].copy()
        idx = np.arange(1 << self.n)
        for k, x in enumerate(xmask):
            x = int(x)
            for q in range(int(offsets[k]), int(offsets[k + 1])):
                sel = idx[(idx & x) == int(pattern[q])]
                a, b = psi[sel].copy(), psi[sel ^ x].copy()
                psi[sel] = cosv[q] * a - sinv[q] * b
                psi[sel ^ x] = sinv[q] * a + cosv[q] * b
        self.buf[0] = psi

    def _matrix(self, packed):
        dim = 1 << self.n
        idx = np.arange(dim)
        m = scipy.sparse.csr_matrix((dim, dim), dtype=np.complex128)
        for x, z, ny, cr, ci in zip(packed.x, packed.z, packed.ny, packed.cre, packed.cim):
            par = np.array([bin(int(i) & int(z)).count("1") & 1 for i in idx])
            vals = complex(cr, ci) * (1j ** int(ny)) * (1 - 2 * par)
            m = m + scipy.sparse.csr_matrix((vals, (idx ^ int(x), idx)), shape=(dim, dim))
        return m

    def apply_exp(self, packed, theta):
        self.buf[0] = scipy.sparse.linalg.expm_multiply(float(theta) * self._matrix(packed), self.buf[0])

    # -- observables
    def paulisum(self, operator):
        return _PS(operator if isinstance(operator, PackedTerms) else pack_operator(operator, with_constant=True))

    def _apply(self, packed, psi):
        out = np.zeros_like(psi)
        for x, z, ny, cr, ci in zip(packed.x, packed.z, packed.ny, packed.cre, packed.cim):
            if cr != 0 or ci != 0:
                out += complex(cr, ci) * orc.apply_pauli(psi, int(x), int(z), int(ny))
        return out

    def expectation(self, ps, buf=0):
        psi = self.buf[buf]
        return complex(np.vdot(psi, self._apply(ps.packed, psi)))

    def apply_paulisum(self, ps, dst=1, src=0):
        self.buf[dst] = self._apply(ps.packed, self.buf[src])

    def pool_overlaps(self, pool, bra=1, ket=0):
        n_ops = len(pool.offsets) - 1
        out = np.zeros(n_ops, dtype=np.complex128)
        for k in range(n_ops):
            a, b = int(pool.offsets[k]), int(pool.offsets[k + 1])
            sub = PackedTerms(pool.n, pool.x[a:b], pool.z[a:b], pool.ny[a:b], pool.cre[a:b], pool.cim[a:b])
            out[k] = np.vdot(self.buf[bra], self._apply(sub, self.buf[ket]))
        return out

    # -- reductions / bookkeeping
    def norm2(self, buf=0):
        return float(np.vdot(self.buf[buf], self.buf[buf]).real)

    def inner(self, a, b):
        return complex(np.vdot(self.buf[a], self.buf[b]))

    def overlap_host(self, vec, buf=0):
        return complex(np.vdot(np.asarray(vec, dtype=np.complex128).reshape(-1), self.buf[buf]))

    def synchronize(self):
        pass
