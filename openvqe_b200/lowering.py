"""Operator lowering: qat-style Pauli operators -> packed bit masks.

The reference hands myQLM objects to its hot path (duck-typed: ``.nbqbits``,
``.terms`` whose items have ``.coeff``, ``.op`` -- a string over I,X,Y,Z -- and
``.qbits``; uses at reference openvqe/adapt/qubit_adapt_vqe.py:105-121).  The
engine works on (xmask, zmask, ny, coeff) arrays in INDEX-BIT space: reference
qubit q is index bit n-1-q (myQLM qubit 0 = most significant bit).

Term order is preserved exactly: rotations are applied in ``.terms`` order.
"""
from __future__ import annotations

import numpy as np


class PackedTerms:
    """Flat arrays for one operator or a list of operators (CSR-style offsets)."""

    __slots__ = ("n", "x", "z", "ny", "cre", "cim", "offsets", "const")

    def __init__(self, n, x, z, ny, cre, cim, offsets=None, const=0j):
        self.n = n
        self.x = np.ascontiguousarray(x, dtype=np.uint64)
        self.z = np.ascontiguousarray(z, dtype=np.uint64)
        self.ny = np.ascontiguousarray(ny, dtype=np.int32)
        self.cre = np.ascontiguousarray(cre, dtype=np.float64)
        self.cim = np.ascontiguousarray(cim, dtype=np.float64)
        self.offsets = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int32)
        self.const = const

    def __len__(self):
        return int(self.x.shape[0])


def term_masks(op, qbits, n):
    x = z = ny = 0
    for letter, q in zip(op, qbits):
        q = int(q)
        if q < 0 or q >= n:
            raise ValueError("qubit %d out of range for %d qubits" % (q, n))
        bit = 1 << (n - 1 - q)
        if letter == "X":
            x |= bit
        elif letter == "Y":
            x |= bit
            z |= bit
            ny += 1
        elif letter == "Z":
            z |= bit
        elif letter != "I":
            raise ValueError("not a Pauli letter: %r" % letter)
    return x, z, ny


def pack_operator(op, with_constant=False):
    """One Hamiltonian-like object -> PackedTerms (its ``.terms`` order)."""
    n = int(op.nbqbits)
    xs, zs, nys, cr, ci = [], [], [], [], []
    const = complex(getattr(op, "constant_coeff", 0.0) or 0.0)
    for t in op.terms:
        x, z, ny = term_masks(t.op, t.qbits, n)
        c = complex(t.coeff)
        xs.append(x)
        zs.append(z)
        nys.append(ny)
        cr.append(c.real)
        ci.append(c.imag)
    if with_constant and const != 0:
        xs.append(0)
        zs.append(0)
        nys.append(0)
        cr.append(const.real)
        ci.append(const.imag)
    return PackedTerms(n, xs, zs, nys, cr, ci, None, const)


def pack_pool(ops):
    """List of operators -> one PackedTerms with ``offsets`` (len(ops)+1)."""
    if len(ops) == 0:
        raise ValueError("empty operator pool")
    n = int(ops[0].nbqbits)
    xs, zs, nys, cr, ci, offs = [], [], [], [], [], [0]
    for op in ops:
        for t in op.terms:
            x, z, ny = term_masks(t.op, t.qbits, n)
            c = complex(t.coeff)
            xs.append(x)
            zs.append(z)
            nys.append(ny)
            cr.append(c.real)
            ci.append(c.imag)
        offs.append(len(xs))
    return PackedTerms(n, xs, zs, nys, cr, ci, offs)


# ---- 2^n x 2^n matrices -> Pauli lists -------------------------------------------------------------------
# The reference's module-level ADAPT helpers take scipy matrices (hamiltonian_sparse, cluster_ops_sparse:
# fermionic_adapt_vqe.py:12-122, qubit_adapt_vqe.py:126-150).  The engine works from Pauli lists, so a matrix argument is
# decomposed once (cached per matrix object): non-zeros are grouped by x = row ^ col, and for every x the vector
# f_x(col) = M[col ^ x, col] is Walsh-Hadamard transformed over the Z patterns:
#     M[i ^ x, i] = sum_z c_(x,z) i^ny (-1)^popcount(i & z)   =>   c_(x,z) i^ny = 2^-n sum_i f_x(i) (-1)^popcount(i & z).
MATRIX_MAX_QUBITS = 14


class _Term:
    __slots__ = ("coeff", "op", "qbits")

    def __init__(self, coeff, op, qbits):
        self.coeff, self.op, self.qbits = coeff, op, qbits


class MatrixOperator:
    """Duck-typed Pauli-sum view (``.nbqbits``, ``.terms``) of a 2^n x 2^n matrix."""

    def __init__(self, nbqbits, terms):
        self.nbqbits, self.terms, self.constant_coeff = nbqbits, terms, 0.0


def _fwht(v):
    v = v.copy()
    h, n = 1, v.shape[0]
    while h < n:
        v = v.reshape(-1, 2, h)
        a, b = v[:, 0, :] + v[:, 1, :], v[:, 0, :] - v[:, 1, :]
        v = np.stack([a, b], axis=1).reshape(-1)
        h *= 2
    return v


def is_matrix(obj) -> bool:
    return hasattr(obj, "shape") and not hasattr(obj, "terms") and len(getattr(obj, "shape", ())) == 2 \
        and obj.shape[0] == obj.shape[1] and obj.shape[0] > 1


_MATRIX_CACHE = {}


def operator_from_matrix(mat, tol=1e-13):
    """Pauli-sum operator of a square 2^n x 2^n (scipy-sparse or dense) matrix, n <= 14.  Cached per matrix object."""
    tagged = getattr(mat, "_vqe_operator", None)
    if tagged is not None:
        return tagged
    key = id(mat)
    hit = _MATRIX_CACHE.get(key)
    if hit is not None and hit[0] is mat:
        return hit[1]
    dim = int(mat.shape[0])
    n = dim.bit_length() - 1
    if (1 << n) != dim or mat.shape[0] != mat.shape[1]:
        raise ValueError("expected a 2^n x 2^n matrix, got shape %r" % (mat.shape,))
    if n > MATRIX_MAX_QUBITS:
        raise ValueError("a %d-qubit matrix argument is not decomposed (limit %d): pass the Pauli-list operator" % (n, MATRIX_MAX_QUBITS))
    if hasattr(mat, "tocoo"):
        coo = mat.tocoo()
        row, col, data = np.asarray(coo.row, dtype=np.int64), np.asarray(coo.col, dtype=np.int64), np.asarray(coo.data, dtype=np.complex128)
    else:
        dense = np.asarray(mat, dtype=np.complex128)
        row, col = np.nonzero(dense)
        data = dense[row, col]
    keep = data != 0
    row, col, data = row[keep], col[keep], data[keep]
    xs = row ^ col
    order = np.argsort(xs, kind="stable")
    xs, col, data = xs[order], col[order], data[order]
    terms = []
    scale = float(np.max(np.abs(data))) if data.size else 0.0
    bounds = np.flatnonzero(np.diff(xs)) + 1
    for lo, hi in zip(np.concatenate([[0], bounds]), np.concatenate([bounds, [xs.size]])):
        if hi <= lo:
            continue
        x = int(xs[lo])
        f = np.zeros(dim, dtype=np.complex128)
        np.add.at(f, col[lo:hi], data[lo:hi])
        c = _fwht(f) / dim
        for z in np.flatnonzero(np.abs(c) > tol * max(scale, 1.0)):
            z = int(z)
            ny = bin(x & z).count("1")
            coeff = c[z] * (-1j) ** ny          # c_(x,z) = c' / i^ny
            op, qb = [], []
            for q in range(n):
                b = n - 1 - q
                xb, zb = (x >> b) & 1, (z >> b) & 1
                if xb or zb:
                    op.append("Y" if xb and zb else ("X" if xb else "Z"))
                    qb.append(q)
            if not op:
                op, qb = ["I"], [0]
            coeff = complex(coeff)
            terms.append(_Term(coeff.real if abs(coeff.imag) <= tol * max(scale, 1.0) else coeff, "".join(op), qb))
    out = MatrixOperator(n, terms)
    if len(_MATRIX_CACHE) >= 4096:
        _MATRIX_CACHE.clear()
    _MATRIX_CACHE[key] = (mat, out)
    return out


def as_operator(obj):
    """Pauli-list operator of ``obj``: itself when it already is one, its decomposition when it is a matrix."""
    return operator_from_matrix(obj) if is_matrix(obj) else obj


def matrix_of(op):
    """scipy CSR matrix of a Pauli-sum operator (n <= 14), tagged with the operator it came from so that handing it
    back to an engine entry point costs nothing.  What the reference's term_to_matrix_sparse returns
    (qubit_adapt_vqe.py:81-123), built from the bit masks instead of a kron chain."""
    import scipy.sparse
    n = int(op.nbqbits)
    if n > MATRIX_MAX_QUBITS:
        raise ValueError("matrix of a %d-qubit operator is not built (limit %d)" % (n, MATRIX_MAX_QUBITS))
    dim = 1 << n
    idx = np.arange(dim, dtype=np.int64)
    rows, cols, vals = [], [], []
    for t in op.terms:
        x, z, ny = term_masks(t.op, t.qbits, n)
        par = np.zeros(dim, dtype=np.int64)
        zz = z
        while zz:
            b = zz & -zz
            par ^= (idx & b) != 0
            zz ^= b
        rows.append(idx ^ x)
        cols.append(idx)
        vals.append(complex(t.coeff) * (1j ** ny) * (1 - 2 * par))
    const = complex(getattr(op, "constant_coeff", 0.0) or 0.0)
    if const != 0:
        rows.append(idx); cols.append(idx); vals.append(np.full(dim, const))
    if rows:
        m = scipy.sparse.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dim, dim))
    else:
        m = scipy.sparse.csr_matrix((dim, dim), dtype=np.complex128)
    m._vqe_operator = op
    return m
