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
