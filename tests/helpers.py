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
