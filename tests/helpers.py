"""Shared test helpers: tiny duck-typed operator classes + random generators."""
import numpy as np


class T:
    """Duck-typed qat Term."""

    def __init__(self, coeff, op, qbits):
        self.coeff, self.op, self.qbits = coeff, op, list(qbits)


class Ham:
    """Duck-typed qat Hamiltonian."""

    def __init__(self, nbqbits, terms, constant_coeff=0.0):
        self.nbqbits, self.terms, self.constant_coeff = nbqbits, list(terms), constant_coeff

    def __mul__(self, scalar):  # qat Hamiltonians support scalar multiplication (reference algorithms/ucc.py:31)
        return Ham(self.nbqbits, [T(t.coeff * scalar, t.op, t.qbits) for t in self.terms], self.constant_coeff * scalar)

    __rmul__ = __mul__


def random_pauli(rng, n, max_weight=None, letters="XYZ"):
    w = int(rng.integers(1, (max_weight or n) + 1))
    qb = sorted(rng.choice(n, size=min(w, n), replace=False).tolist())
    op = "".join(rng.choice(list(letters), size=len(qb)))
    return op, qb


def random_hermitian(rng, n, n_terms, max_weight=None, const=0.0):
    terms = []
    for _ in range(n_terms):
        op, qb = random_pauli(rng, n, max_weight)
        terms.append(T(float(rng.normal()), op, qb))
    return Ham(n, terms, const)


def random_antihermitian(rng, n, n_terms, max_weight=None):
    terms = []
    for _ in range(n_terms):
        op, qb = random_pauli(rng, n, max_weight)
        terms.append(T(1j * float(rng.normal()), op, qb))
    return Ham(n, terms)


def random_state(rng, n):
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


# ---- golden fixtures ---------------------------------------------------------------------------
import gzip
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rt") as f:
            return json.load(f)
    with open(path) as f:
        return json.load(f)


def ham_from_json(d):
    terms = [T(complex(cr, ci) if ci != 0 else cr, op, qb) for cr, ci, op, qb in d["terms"]]
    c = d.get("constant", 0.0)
    const = complex(c[0], c[1]) if isinstance(c, list) else c
    if isinstance(const, complex) and const.imag == 0:
        const = const.real
    return Ham(d["nbqbits"], terms, const)


def pool_from_json(nbqbits, plist):
    return [Ham(nbqbits, [T(complex(cr, ci), op, qb) for cr, ci, op, qb in terms]) for terms in plist]


class FermiOp:
    """Duck-typed fermionic cluster operator: QUCCSD only reads terms[0].qbits."""

    def __init__(self, nbqbits, qbits):
        self.nbqbits = nbqbits
        self.terms = [T(1.0, "Cc" if len(qbits) == 2 else "CCcc", qbits)]


def jw_excitation(nbqbits, create, annihilate):
    """i*(T - T^dagger) of one excitation, Jordan-Wigner, via the oracle's qat stand-in
    (Hermitian generator with real Pauli coefficients, as reference algorithms/ucc.py:24-31)."""
    import sys
    shim = os.path.join(os.path.dirname(GOLDEN), "..", "oracle", "qat_shim")
    if shim not in sys.path:
        sys.path.insert(0, os.path.abspath(shim))
    from qat.core import Term as QTerm
    from qat.fermion import FermionHamiltonian
    op = "C" * len(create) + "c" * len(annihilate)
    fwd = QTerm(1.0, op, list(create) + list(annihilate))
    bwd = QTerm(-1.0, op, list(reversed(annihilate)) + list(reversed(create)))
    sp = FermionHamiltonian(nbqbits, [fwd, bwd]).to_spin()
    return Ham(nbqbits, [T((1j * complex(t.coeff)).real, t.op, t.qbits) for t in sp.terms])
