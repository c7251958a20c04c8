"""Operator pools generated directly as Pauli lists / packed bit masks.

The reference builds its ADAPT pools as myQLM objects and then turns every operator into a 2^n x 2^n scipy matrix
(``molecule_factory_with_sparse.py:615-617``), which caps ADAPT near 14 qubits.  The engine only needs the Pauli
strings, so the pools are produced here without any matrix:

  qubit pools      ``generate_yxxx_pool`` / ``xyxx`` / ``xxyx`` / ``xxxy`` / ``generate_random_pool``
                   (reference ``qubit_pool.py:278-465``): single-string operators, coefficient -1, the same
                   ``itertools.combinations`` order and parity filters; ``packed_qubit_pool`` is the vectorised form.
  fermionic pools  ``spin_complement_gsd`` (reference ``generator_excitations.py:83-156``), ``uccgsd`` (:555-609, defined
                   in the reference but never wired into ``generate_cluster_ops``), ``singlet_upccgsd`` (:403-465):
                   the same loops over spin-orbital indices, normal ordering of every ladder-operator product by the
                   anticommutation rules, merging of duplicate terms (zero coefficients kept -- the reference's pools
                   contain identically-zero operators), then the Jordan-Wigner image.
  ``generate_cluster_ops``  the dispatcher of ``molecule_factory_with_sparse.py:573-617`` with ``uccgsd`` wired in and
                   without the sparse matrices.

Every generator returns ``(pool_size, cluster_ops, cluster_ops_sp)`` like the reference: ``cluster_ops`` are light
fermionic operators (``.terms[k].op`` over C/c, ``.qbits``), ``cluster_ops_sp`` duck-typed Pauli-sum operators
(``.nbqbits``, ``.terms`` with ``.coeff/.op/.qbits``) that every engine entry point accepts.  Only the Jordan-Wigner
transform is generated here (the engine itself is transform-agnostic: Bravyi-Kitaev / parity pools built elsewhere
work unchanged).  Term order inside an operator is the order of first appearance while the ladder products are
expanded left to right -- the order the fixtures under tests/golden were produced with.
"""
from __future__ import annotations

import itertools

import numpy as np

from ..lowering import PackedTerms


class Term:
    __slots__ = ("coeff", "op", "qbits")

    def __init__(self, coeff, op, qbits):
        self.coeff, self.op, self.qbits = coeff, op, list(qbits)

    def __repr__(self):
        return "Term(%r, %r, %r)" % (self.coeff, self.op, self.qbits)


class Operator:
    """Duck-typed qat Hamiltonian: what the reference hands around (``.nbqbits``, ``.terms``, ``.constant_coeff``)."""

    def __init__(self, nbqbits, terms, constant_coeff=0.0):
        self.nbqbits, self.terms, self.constant_coeff = int(nbqbits), list(terms), constant_coeff

    def __mul__(self, scalar):
        return Operator(self.nbqbits, [Term(t.coeff * scalar, t.op, t.qbits) for t in self.terms], self.constant_coeff * scalar)

    __rmul__ = __mul__


# ---------------------------------------------------------------------------------------------------------
# qubit pools: single Pauli strings
# ---------------------------------------------------------------------------------------------------------
def _string_pool(nbqbits, four_letters):
    pool = []
    for a, b in itertools.combinations(range(nbqbits), 2):
        if (a + b) % 2 == 0:
            pool.append(Operator(nbqbits, [Term(-1.0, "YX", [a, b])]))
    for a, b, c, d in itertools.combinations(range(nbqbits), 4):
        if (a % 2 + b % 2 + c % 2 + d % 2) % 2 == 0:
            pool.append(Operator(nbqbits, [Term(-1.0, four_letters, [a, b, c, d])]))
    return len(pool), pool


def generate_yxxx_pool(nbqbits):
    """reference qubit_pool.py:278-312"""
    return _string_pool(nbqbits, "YXXX")


def generate_xyxx_pool(nbqbits):
    """reference qubit_pool.py:314-350"""
    return _string_pool(nbqbits, "XYXX")


def generate_xxyx_pool(nbqbits):
    """reference qubit_pool.py:352-389"""
    return _string_pool(nbqbits, "XXYX")


def generate_xxxy_pool(nbqbits):
    """reference qubit_pool.py:391-428"""
    return _string_pool(nbqbits, "XXXY")


def generate_random_pool(yxxx_pool, xyxx_pool, xxyx_pool, xxxy_pool):
    """One of the four variants per position, drawn with numpy's global generator exactly as the reference does
    (qubit_pool.py:430-465: ``np.random.randint(0, 4)`` per operator) -- seed ``np.random`` to pin it."""
    options = [yxxx_pool, xyxx_pool, xxyx_pool, xxxy_pool]
    pool = [options[np.random.randint(0, 4)][i] for i in range(len(xxxy_pool))]
    return len(pool), pool


def packed_qubit_pool(nbqbits, four_letters="YXXX"):
    """The same pool as ``generate_*_pool`` as packed masks (vectorised; a 24-qubit pool has 5 445 operators, a 30-qubit
    one 13 905): operator k is the single string k with coefficient -1."""
    n = int(nbqbits)
    bit = lambda q: np.uint64(1) << np.uint64(n - 1 - q)
    xs, zs, nys = [], [], []
    for a, b in itertools.combinations(range(n), 2):
        if (a + b) % 2 == 0:
            xs.append(int(bit(a) | bit(b)))
            zs.append(int(bit(a)))
            nys.append(1)
    ypos = four_letters.index("Y")
    for q4 in itertools.combinations(range(n), 4):
        if sum(q % 2 for q in q4) % 2 == 0:
            xs.append(int(bit(q4[0]) | bit(q4[1]) | bit(q4[2]) | bit(q4[3])))
            zs.append(int(bit(q4[ypos])))
            nys.append(1)
    m = len(xs)
    return PackedTerms(n, xs, zs, nys, -np.ones(m), np.zeros(m), np.arange(m + 1, dtype=np.int32))


# ---------------------------------------------------------------------------------------------------------
# fermionic algebra: normal ordering and the Jordan-Wigner image, on index lists and bit masks
# ---------------------------------------------------------------------------------------------------------
def _normal_order(coeff, op, qbits):
    """Creation operators to the left of annihilation operators by {c_p, C_q} = delta_pq (reference
    fermion_util.order_fermionic_ops): the first annihilator that has a creator somewhere to its right is the anchor;
    the creator nearest to it on the right is moved one position to the left (sign flip, plus the contraction when
    the two act on the same mode), contraction first, depth first."""
    ic = op.find("c")
    if ic < 0:
        return [(coeff, op, qbits)]
    jc = op.find("C", ic)
    if jc < 0:
        return [(coeff, op, qbits)]
    i = jc - 1
    swapped_op = op[:i] + op[i + 1] + op[i] + op[i + 2:]
    swapped_q = qbits[:i] + [qbits[i + 1], qbits[i]] + qbits[i + 2:]
    out = []
    if op[i] != op[i + 1] and qbits[i] == qbits[i + 1]:
        out += _normal_order(coeff, op[:i] + op[i + 2:], qbits[:i] + qbits[i + 2:])
    out += _normal_order(-coeff, swapped_op, swapped_q)
    return out


def _sort_block(qs):
    """Ascending insertion order of a block of like operators; (sorted, sign) or None when a mode repeats."""
    qs = list(qs)
    sign = 1
    while True:
        i = 0
        while i < len(qs) - 1 and qs[i] <= qs[i + 1]:
            if qs[i] == qs[i + 1]:
                return None
            i += 1
        if i >= len(qs) - 1:
            return qs, sign
        i += 1
        j = 0
        while qs[j] < qs[i]:
            j += 1
        qs.insert(j, qs.pop(i))
        if (i - j) & 1:
            sign = -sign


def order_fermionic_term(coeff, op, qbits):
    """List of (coeff, op, qbits) with creators left (ascending modes) and annihilators right (ascending modes)
    (reference fermion_util.order_fermionic_term)."""
    out = []
    for c, o, q in _normal_order(coeff, op, list(qbits)):
        k = o.find("c")
        k = len(o) if k < 0 else k
        left, right = _sort_block(q[:k]), _sort_block(q[k:])
        if left is None or right is None:
            continue
        out.append((c * left[1] * right[1], o, left[0] + right[0]))
    return out


def _merge(terms):
    """Duplicate (op, qbits) keys merged in first-appearance order, zero coefficients kept (the constructor of the
    reference's FermionHamiltonian behaves like this -- it is how the pool sizes 175 / 69 / 70 of the reference's
    tests come about)."""
    out, pos = [], {}
    for c, o, q in terms:
        key = (o, tuple(q))
        if key in pos:
            out[pos[key]][0] += c
        else:
            pos[key] = len(out)
            out.append([c, o, list(q)])
    return out


def _ladder(kind, p):
    """JW image of c_p / C_p as {(x, z): coeff} over products X^x Z^z (bit q = qubit q, Y = i X Z):
    c_p = Z_0 .. Z_(p-1) (X_p + i Y_p) / 2."""
    chain, xp = (1 << p) - 1, 1 << p
    return (((xp, chain), 0.5), ((xp, chain | xp), -0.5 if kind == "c" else 0.5))


def jordan_wigner(nbqbits, fermionic_terms, drop_tol=1e-14):
    """[(coeff, letters, qubits)] of sum_k coeff_k prod ladder operators, in order of first appearance."""
    acc, order = {}, []
    for coeff, op, qbits in fermionic_terms:
        prod = {(0, 0): complex(coeff)}
        for letter, q in zip(op, qbits):
            nxt = {}
            lad = _ladder(letter, q)
            for (x1, z1), c1 in prod.items():
                for (x2, z2), c2 in lad:
                    key = (x1 ^ x2, z1 ^ z2)
                    val = (-c1 * c2) if (bin(z1 & x2).count("1") & 1) else (c1 * c2)
                    nxt[key] = nxt.get(key, 0.0) + val
            prod = nxt
        for key, c in prod.items():
            if key not in acc:
                order.append(key)
                acc[key] = 0.0
            acc[key] += c
    out = []
    for x, z in order:
        c = acc[(x, z)]
        if (x, z) == (0, 0) or abs(c) <= drop_tol:
            continue
        ny = bin(x & z).count("1")
        letters, qubits = [], []
        for q in range(nbqbits):
            k = ((x >> q) & 1, (z >> q) & 1)
            if k != (0, 0):
                letters.append("X" if k == (1, 0) else ("Y" if k == (1, 1) else "Z"))
                qubits.append(q)
        out.append((c * ((-1j) ** ny), "".join(letters), qubits))
    if not out:
        out = [(0.0, "I", [0])]  # an identically-zero operator stays in the pool as one explicit zero term
    return out


def _finish(nbqbits, fermionic_ops, perm=0):
    """reference generator_excitations._apply_transforms for 'JW' (:16-36): nothing is filtered; ``perm`` repeats."""
    cluster_ops = [Operator(nbqbits, [Term(c, o, q) for c, o, q in terms]) for terms in fermionic_ops]
    cluster_ops_sp = [Operator(nbqbits, [Term(c, o, q) for c, o, q in jordan_wigner(nbqbits, terms)]) for terms in fermionic_ops]
    cluster_ops = cluster_ops + cluster_ops * perm
    cluster_ops_sp = cluster_ops_sp + cluster_ops_sp * perm
    return len(cluster_ops_sp), cluster_ops, cluster_ops_sp


def _check_transform(transform):
    if transform != "JW":
        raise NotImplementedError("openvqe_b200.common_files.pools generates Jordan-Wigner pools only (got %r); pools in "
                                  "another encoding can be passed to the engine as Pauli lists" % (transform,))


def spin_complement_gsd(n_elec, orbital_number, transform="JW"):
    """Spin-complemented generalised singles and doubles (reference generator_excitations.py:83-156)."""
    _check_transform(transform)
    n = 2 * orbital_number
    singles, doubles = [], []
    for p in range(0, n, 2):
        for q in range(p, n, 2):
            singles.append(_merge([(1, "Cc", [p, q]), (-1, "Cc", [q, p]), (1, "Cc", [p + 1, q + 1]), (-1, "Cc", [q + 1, p + 1])]))
            for r in range(p, n, 2):
                for s in range(q if r == p else r, n, 2):
                    term_a = [(1, "CcCc", [r, p, s, q]), (-1, "CcCc", [q, s, p, r]),
                              (1, "CcCc", [r + 1, p + 1, s + 1, q + 1]), (-1, "CcCc", [q + 1, s + 1, p + 1, r + 1])]
                    term_b = [(1, "CcCc", [r, p, s + 1, q + 1]), (-1, "CcCc", [q + 1, s + 1, p, r]),
                              (1, "CcCc", [r + 1, p + 1, s, q]), (-1, "CcCc", [q, s, p + 1, r + 1])]
                    term_c = [(1, "CcCc", [r, p + 1, s + 1, q]), (-1, "CcCc", [q, s + 1, p + 1, r]),
                              (1, "CcCc", [r + 1, p, s, q + 1]), (-1, "CcCc", [q + 1, s, p, r + 1])]
                    for term_x in (term_a, term_b, term_c):
                        doubles.append(_merge(sum((order_fermionic_term(*t) for t in term_x), [])))
    return _finish(n, singles + doubles)


def uccgsd(n_elec, orbital_number, transform="JW"):
    """Generalised singles and doubles over spin orbitals (reference generator_excitations.py:555-609; present in the
    reference but not reachable through its ``generate_cluster_ops``)."""
    _check_transform(transform)
    n = 2 * orbital_number
    singles, doubles = [], []
    for p in range(0, n):
        for q in range(p, n):
            singles.append(_merge([(1, "Cc", [p, q]), (-1, "Cc", [q, p])]))
            for r in range(p, n):
                for s in range(q if r == p else r, n):
                    term_a = [(1, "CCcc", [p, q, r, s]), (-1, "CCcc", [s, r, q, p])]
                    doubles.append(_merge(sum((order_fermionic_term(*t) for t in term_a), [])))
    return _finish(n, singles + doubles)


def singlet_upccgsd(n_orb, transform="JW", perm=0):
    """Paired generalised singles and doubles (k-UpCCGSD with k = perm + 1; reference generator_excitations.py:403-465)."""
    _check_transform(transform)
    n = 2 * n_orb
    singles, doubles = [], []
    for p in range(0, n, 2):
        for q in range(0, p, 2):
            singles.append(_merge([(1, "Cc", [q, p]), (-1, "Cc", [p, q]), (1, "Cc", [q + 1, p + 1]), (-1, "Cc", [p + 1, q + 1])]))
    for p, q in itertools.combinations(range(0, n, 2), 2):
        term_a = [(1.0, "CcCc", [q, p, q + 1, p + 1]), (-1.0, "CcCc", [p + 1, q + 1, p, q])]
        doubles.append(_merge(sum((order_fermionic_term(*t) for t in term_a), [])))
    return _finish(n, singles + doubles, perm=perm)


def generate_cluster_ops(type_of_generator, n_elec, orbital_number, transform="JW"):
    """The pool dispatcher of reference molecule_factory_with_sparse.py:573-617 without the 2^n x 2^n matrices and with
    ``uccgsd`` wired in.  Returns ``(pool_size, cluster_ops, cluster_ops_sp)``; None for an unknown name, as the
    reference."""
    if type_of_generator == "spin_complement_gsd":
        return spin_complement_gsd(n_elec, orbital_number, transform)
    if type_of_generator == "uccgsd":
        return uccgsd(n_elec, orbital_number, transform)
    if type_of_generator == "sUPCCGSD":
        return singlet_upccgsd(orbital_number, transform, 0)
    return None
