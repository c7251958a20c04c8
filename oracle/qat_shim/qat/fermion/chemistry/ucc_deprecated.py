"""TEST INFRASTRUCTURE ONLY -- stand-in for ``qat.fermion.chemistry.ucc_deprecated``
(myqlm-fermion 1.1.4).  ``build_ucc_ansatz`` semantics (SURVEY Appendix A V5 and
Appendix B item 10): X gates for the set bits of ``ket_hf`` (qubit 0 = MSB), then
for every Pauli string c*P of every cluster operator, in ``terms`` order, the
staircase circuit of exp(-i theta c P): H on X letters, RX(+pi/2) on Y letters,
CNOT chain, RZ(2 theta c), uncompute with H / RX(-pi/2)."""
import itertools
from math import pi

import numpy as np

from ...core import Term
from ...lang.AQASM import CNOT, H, RX, RZ, X, QRoutine
from .. import FermionHamiltonian


def _pauli_rotation(rout, op, qbits, angle):
    act = [(l, q) for l, q in zip(op, qbits) if l != "I"]
    if not act:
        return
    for l, q in act:
        if l == "X":
            rout.apply(H, q)
        elif l == "Y":
            rout.apply(RX(pi / 2), q)
    qs = [q for _, q in act]
    for a, b in zip(qs[:-1], qs[1:]):
        rout.apply(CNOT, a, b)
    rout.apply(RZ(2.0 * angle), qs[-1])
    for a, b in reversed(list(zip(qs[:-1], qs[1:]))):
        rout.apply(CNOT, a, b)
    for l, q in act:
        if l == "X":
            rout.apply(H, q)
        elif l == "Y":
            rout.apply(RX(-pi / 2), q)


def build_ucc_ansatz(cluster_ops, ket_hf, n_steps=1):
    nbqbits = cluster_ops[0].nbqbits

    def qfunc(theta):
        rout = QRoutine()
        for q in range(nbqbits):
            if (int(ket_hf) >> (nbqbits - 1 - q)) & 1:
                rout.apply(X, q)
        for _ in range(n_steps):
            for th, op in zip(theta, cluster_ops):
                for t in op.terms:
                    c = complex(t.coeff)
                    if c == 0:
                        continue
                    _pauli_rotation(rout, t.op, t.qbits, float(th) * c.real / n_steps)
        return rout

    return qfunc


def get_cluster_ops_and_init_guess(n_electrons, noons, orbital_energies, hpqrs):
    """UCCSD excitation list + MP2 guess.  Layout of ``terms[0].qbits`` is
    [a, i] / [a, b, i, j] (virtuals first).  The operator
    ORDER emitted by the real myQLM routine is unverified (V9)."""
    n = len(noons)
    occ = list(range(n_electrons))
    virt = list(range(n_electrons, n))
    ops, theta = [], []
    e = list(orbital_energies)
    for i in occ:
        for a in virt:
            if (a - i) % 2 == 0:
                ops.append(FermionHamiltonian(n, [Term(1j, "Cc", [a, i]), Term(-1j, "Cc", [i, a])]))
                theta.append(0.0)
    for i, j in itertools.combinations(occ, 2):
        for a, b in itertools.combinations(virt, 2):
            if (i % 2 + j % 2) != (a % 2 + b % 2):
                continue
            if (i % 2 == j % 2) and (a % 2 != i % 2):
                continue
            ops.append(FermionHamiltonian(n, [Term(1j, "CCcc", [a, b, i, j]),
                                               Term(-1j, "CCcc", [j, i, b, a])]))
            num = hpqrs[a, b, i, j] - hpqrs[a, b, j, i]
            theta.append(float(num / (e[i] + e[j] - e[a] - e[b])))
    hf_init = 0
    for q in occ:
        hf_init |= 1 << (n - 1 - q)
    return ops, theta, hf_init


def get_active_space_hamiltonian(*a, **kw):
    raise NotImplementedError("qat shim: active-space selection not restated")
