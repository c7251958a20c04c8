"""TEST INFRASTRUCTURE ONLY -- stand-in for ``qat.fermion`` (myqlm-fermion 1.1.4).

Provides ``SpinHamiltonian``, ``FermionHamiltonian`` and
``ElectronicStructureHamiltonian`` with the operations the reference uses:
``.nbqbits``, ``.terms``, ``.constant_coeff``, ``+ - *``, ``.get_matrix(sparse)``,
``.hpqrs``.  Conventions calibrated against the reference's stored notebook
outputs: qubit 0 is the most significant bit of
the basis-state index; duplicate ``(op, qbits)`` keys are merged by the
constructor and zero coefficients are kept.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from ..core import Term

_LETTER_XZ = {"I": (0, 0), "X": (1, 0), "Y": (1, 1), "Z": (0, 1)}
_XZ_LETTER = {(0, 0): "I", (1, 0): "X", (1, 1): "Y", (0, 1): "Z"}


def _merge_terms(terms):
    """Merge duplicate (op, qbits) keys, first-appearance order, keep zeros."""
    out, pos = [], {}
    for t in terms:
        k = t.key()
        if k in pos:
            out[pos[k]].coeff = out[pos[k]].coeff + t.coeff
        else:
            pos[k] = len(out)
            out.append(t.copy())
    return out


# ---------------------------------------------------------------------------
# Pauli algebra in the X^x Z^z product form.  A spin operator is a dict
# {(x, z): coeff} meaning sum coeff * prod_q X_q^{x_q} Z_q^{z_q}; bit q of the
# python ints x, z refers to QUBIT q (not to the index bit).  Y = i X Z.
# ---------------------------------------------------------------------------
def _popc(v):
    return bin(v).count("1")


def xz_mul(a, b):
    out = {}
    for (x1, z1), c1 in a.items():
        for (x2, z2), c2 in b.items():
            sign = -1.0 if (_popc(z1 & x2) & 1) else 1.0
            k = (x1 ^ x2, z1 ^ z2)
            out[k] = out.get(k, 0.0) + sign * c1 * c2
    return out


def xz_add(a, b, scale=1.0):
    out = dict(a)
    for k, c in b.items():
        out[k] = out.get(k, 0.0) + scale * c
    return out


def term_to_xz(term):
    """Pauli-letter term -> ((x, z), coeff) in product form."""
    x = z = 0
    for letter, q in zip(term.op, term.qbits):
        lx, lz = _LETTER_XZ[letter]
        if (x >> q) & 1 or (z >> q) & 1:
            raise ValueError("repeated qubit in Pauli term %r" % term)
        x |= lx << q
        z |= lz << q
    ny = _popc(x & z)
    return (x, z), complex(term.coeff) * (1j ** ny)


def xz_to_term(x, z, coeff, nbqbits):
    ny = _popc(x & z)
    letters, qbits = [], []
    for q in range(nbqbits):
        k = ((x >> q) & 1, (z >> q) & 1)
        if k != (0, 0):
            letters.append(_XZ_LETTER[k])
            qbits.append(q)
    c = coeff * ((-1j) ** ny)
    return Term(c, "".join(letters), qbits)


class _HamiltonianBase:
    def __init__(self, nqbits, terms=None, constant_coeff=0.0, do_clean_up=True):
        self.nbqbits = int(nqbits)
        self.terms = _merge_terms(list(terms) if terms is not None else [])
        self.constant_coeff = constant_coeff

    def _new(self, terms, constant):
        return type(self)(self.nbqbits, terms, constant)

    def copy(self):
        return self._new([t.copy() for t in self.terms], self.constant_coeff)

    # -- linear structure ---------------------------------------------------
    def __add__(self, other):
        if isinstance(other, (int, float, complex)):
            return self._new(self.terms, self.constant_coeff + other)
        return self._new(self.terms + other.terms, self.constant_coeff + other.constant_coeff)

    __radd__ = __add__

    def __neg__(self):
        return self * -1.0

    def __sub__(self, other):
        return self + (other * -1.0)

    def __truediv__(self, scalar):
        return self * (1.0 / scalar)

    def _scale(self, s):
        return self._new([Term(t.coeff * s, t.op, t.qbits) for t in self.terms],
                         self.constant_coeff * s)

    def __rmul__(self, other):
        if isinstance(other, (int, float, complex, np.number)):
            return self._scale(other)
        return NotImplemented

    def __repr__(self):
        body = " +\n".join(repr(t) for t in self.terms)
        return "%s * I^%d +\n%s" % (self.constant_coeff, self.nbqbits, body)


class SpinHamiltonian(_HamiltonianBase):
    """Sum of Pauli strings.  ``terms[k].op`` is over I,X,Y,Z."""

    def to_xz(self):
        d = {}
        if self.constant_coeff != 0:
            d[(0, 0)] = complex(self.constant_coeff)
        for t in self.terms:
            k, c = term_to_xz(t)
            d[k] = d.get(k, 0.0) + c
        return d

    @classmethod
    def from_xz(cls, nbqbits, d, keep_zero_if_empty=True, tol=0.0):
        terms, const = [], 0.0
        for (x, z), c in d.items():
            if x == 0 and z == 0:
                const = c
                continue
            if abs(c) <= tol:
                continue
            terms.append(xz_to_term(x, z, c, nbqbits))
        return cls(nbqbits, terms, const)

    def __mul__(self, other):
        if isinstance(other, (int, float, complex, np.number)):
            return self._scale(other)
        if isinstance(other, SpinHamiltonian):
            return SpinHamiltonian.from_xz(self.nbqbits, xz_mul(self.to_xz(), other.to_xz()))
        return NotImplemented

    def get_matrix(self, sparse=False):
        n = self.nbqbits
        dim = 1 << n
        idx = np.arange(dim, dtype=np.int64)
        mat = sp.csr_matrix((dim, dim), dtype=np.complex128)
        if self.constant_coeff != 0:
            mat = mat + complex(self.constant_coeff) * sp.identity(dim, dtype=np.complex128, format="csr")
        for t in self.terms:
            xm = zm = 0
            ny = 0
            for letter, q in zip(t.op, t.qbits):
                lx, lz = _LETTER_XZ[letter]
                bit = n - 1 - q
                xm |= lx << bit
                zm |= lz << bit
                ny += lx & lz
            par = np.zeros(dim, dtype=np.int64)
            v = idx & zm
            while True:
                nz = v != 0
                if not nz.any():
                    break
                par ^= v & 1
                v = v >> 1
            vals = complex(t.coeff) * (1j ** ny) * (1 - 2 * par)
            mat = mat + sp.csr_matrix((vals, (idx ^ xm, idx)), shape=(dim, dim))
        mat = sp.csr_matrix(mat)
        return mat if sparse else mat.toarray()


def _ladder_xz(kind, p):
    """JW image of c_p / C_p (C = creation) in X^x Z^z form.
    c_p = Z_0..Z_{p-1} (X_p + iY_p)/2 ; Y = iXZ  (SURVEY Appendix A, V1)."""
    chain = (1 << p) - 1
    xp = 1 << p
    if kind == "c":
        return {(xp, chain): 0.5, (xp, chain | xp): -0.5}
    if kind == "C":
        return {(xp, chain): 0.5, (xp, chain | xp): 0.5}
    raise ValueError(kind)


class FermionHamiltonian(_HamiltonianBase):
    """Sum of ladder-operator strings (``op`` over C = creation, c = annihilation)."""

    def __mul__(self, other):
        if isinstance(other, (int, float, complex, np.number)):
            return self._scale(other)
        return NotImplemented

    def to_spin(self, drop_tol=1e-14):
        acc = {}
        order = []
        for t in self.terms:
            prod = {(0, 0): complex(t.coeff)}
            for letter, q in zip(t.op, t.qbits):
                prod = xz_mul(prod, _ladder_xz(letter, q))
            for k, c in prod.items():
                if k not in acc:
                    order.append(k)
                    acc[k] = 0.0
                acc[k] += c
        const = complex(self.constant_coeff)
        terms = []
        for k in order:
            c = acc[k]
            if k == (0, 0):
                const += c
                continue
            if abs(c) <= drop_tol:
                continue
            terms.append(xz_to_term(k[0], k[1], c, self.nbqbits))
        if not terms:
            # myQLM evidently never returns an empty term list (the reference's
            # `terms != []` filter at generator_excitations.py:30 never fires,
            # SURVEY Appendix A V2): keep one explicit zero term.
            terms = [Term(0.0, "I", [0])]
        if abs(const.imag) == 0:
            const = const.real
        return SpinHamiltonian(self.nbqbits, terms, const)

    def get_matrix(self, sparse=False):
        return self.to_spin().get_matrix(sparse=sparse)


class ElectronicStructureHamiltonian(FermionHamiltonian):
    """H = sum hpq C_p c_q + 1/2 sum hpqrs C_p C_q c_r c_s + const (spin-orbital
    tensors; the myQLM convention, cf. reference molecule_factory.py:333-340)."""

    def __init__(self, hpq, hpqrs=None, constant_coeff=0.0, do_clean_up=True):
        if isinstance(hpq, (int, np.integer)):  # generic (nqbits, terms, const) form
            super().__init__(hpq, hpqrs, constant_coeff)
            self.hpq = self.hpqrs = None
            return
        hpq = np.asarray(hpq)
        n = hpq.shape[0]
        self.hpq = hpq
        self.hpqrs = np.zeros((n, n, n, n)) if hpqrs is None else np.asarray(hpqrs)
        terms = []
        for p, q in zip(*np.nonzero(hpq)):
            terms.append(Term(hpq[p, q], "Cc", [int(p), int(q)]))
        for p, q, r, s in zip(*np.nonzero(self.hpqrs)):
            if p == q or r == s:
                continue
            terms.append(Term(0.5 * self.hpqrs[p, q, r, s], "CCcc", [int(p), int(q), int(r), int(s)]))
        super().__init__(n, terms, constant_coeff)

    def _new(self, terms, constant):
        return FermionHamiltonian(self.nbqbits, terms, constant)
