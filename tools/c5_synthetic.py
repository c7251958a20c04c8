"""Synthetic C5 workload (SURVEY.md section 8d "Concrete inputs", BASELINE.json configs[4]): a UCC-like
Pauli-rotation program and a Hamiltonian with a bounded number of X-mask groups for 30-36 qubits, where a molecular
Hamiltonian is infeasible (~3e4 X-mask groups x 1.1 TB).  Everything is produced directly as packed bit masks in
INDEX-BIT space (reference qubit q = index bit n-1-q), deterministic from the seed.

  generators : K JW excitations i(T - T^dagger): singles X_p Z.. Y_q - Y_p Z.. X_q (2 strings, coefficients +-1/2)
               and spin-conserving doubles on p<q<r<s (8 strings with an odd number of Y, coefficients +-1/8,
               Z chains on p+1..q-1 and r+1..s-1); theta ~ U(-0.1, 0.1)
  Hamiltonian: 64 X-mask groups -- the diagonal one (n Z strings + C(n,2)/4 ZZ strings + a constant), 15 weight-2
               groups (hopping-like X Z.. X / Y Z.. Y, plus number-operator-dressed variants) and 48 weight-4
               groups with the 8 even-Y sign variants; coefficients ~ N(0,1) / number of strings

Used by bench.py (--workload c5) and by the parity tests (small n, against the oracle)."""
from __future__ import annotations

import numpy as np

_DOUBLE = [("XXXY", +1), ("XXYX", +1), ("XYXX", -1), ("YXXX", -1), ("YYYX", -1), ("YYXY", -1), ("YXYY", +1), ("XYYY", +1)]
_QUAD_H = ["XXXX", "XXYY", "XYXY", "XYYX", "YXXY", "YXYX", "YYXX", "YYYY"]


def _masks(n, letters):
    """letters: dict qubit -> 'X'|'Y'|'Z'"""
    x = z = ny = 0
    for q, l in letters.items():
        bit = 1 << (n - 1 - q)
        if l in "XY":
            x |= bit
        if l in "YZ":
            z |= bit
        if l == "Y":
            ny += 1
    return x, z, ny


def _chain(letters, a, b):
    for q in range(a + 1, b):
        letters[q] = "Z"


def generators(n, k_gen=256, seed=2026):
    """-> dict(x, z, ny, coeff, owner, theta): rotation r has angle theta[owner[r]] * coeff[r]."""
    rng = np.random.default_rng(seed)
    xs, zs, nys, cs, own = [], [], [], [], []
    for j in range(k_gen):
        if rng.random() < 0.25:
            while True:
                p, q = sorted(rng.choice(n, size=2, replace=False).tolist())
                if (p - q) % 2 == 0:
                    break
            for lp, lq, c in (("X", "Y", 0.5), ("Y", "X", -0.5)):
                letters = {p: lp, q: lq}
                _chain(letters, p, q)
                x, z, ny = _masks(n, letters)
                xs.append(x); zs.append(z); nys.append(ny); cs.append(c); own.append(j)
        else:
            while True:
                p, q, r, s = sorted(rng.choice(n, size=4, replace=False).tolist())
                if (p + q) % 2 == (r + s) % 2:
                    break
            for pat, sg in _DOUBLE:
                letters = dict(zip((p, q, r, s), pat))
                _chain(letters, p, q)
                _chain(letters, r, s)
                x, z, ny = _masks(n, letters)
                xs.append(x); zs.append(z); nys.append(ny); cs.append(sg / 8.0); own.append(j)
    theta = rng.uniform(-0.1, 0.1, size=k_gen)
    return {"x": np.array(xs, dtype=np.uint64), "z": np.array(zs, dtype=np.uint64), "ny": np.array(nys, dtype=np.int32),
            "coeff": np.array(cs), "owner": np.array(own, dtype=np.int64), "theta": theta, "n_generators": k_gen}


def hamiltonian(n, seed=2026, n_pair_groups=15, n_quad_groups=48):
    """-> dict(x, z, ny, cre): real Hermitian Pauli sum with 1 + n_pair_groups + n_quad_groups X-mask groups."""
    rng = np.random.default_rng(seed + 1)
    terms = []
    diag = [{q: "Z"} for q in range(n)]
    n_zz = max(1, (n * (n - 1) // 2) // 4)
    seen = set()
    while len(seen) < n_zz:
        a, b = sorted(rng.choice(n, size=2, replace=False).tolist())
        seen.add((a, b))
    diag += [{a: "Z", b: "Z"} for a, b in sorted(seen)]
    terms += [({}, float(rng.normal()))]  # constant
    terms += [(l, float(rng.normal()) / len(diag)) for l in diag]
    used = set()
    while len(used) < n_pair_groups:
        p, q = sorted(rng.choice(n, size=2, replace=False).tolist())
        if (p - q) % 2 or (p, q) in used:
            continue
        outside = [k for k in range(n) if k < p or k > q]
        if len(outside) < 4:
            continue
        used.add((p, q))
        dress = [None] + [int(k) for k in rng.choice(outside, size=4, replace=False)]
        for lp in "XY":
            for k in dress:
                letters = {p: lp, q: lp}
                _chain(letters, p, q)
                if k is not None:
                    letters[k] = "Z"
                terms.append((letters, float(rng.normal()) / (2 * len(dress))))
    usedq = set()
    while len(usedq) < n_quad_groups:
        p, q, r, s = sorted(rng.choice(n, size=4, replace=False).tolist())
        if (p, q, r, s) in usedq:
            continue
        usedq.add((p, q, r, s))
        for pat in _QUAD_H:
            letters = dict(zip((p, q, r, s), pat))
            _chain(letters, p, q)
            _chain(letters, r, s)
            terms.append((letters, float(rng.normal()) / 8.0))
    xs, zs, nys, cs = [], [], [], []
    for letters, c in terms:
        x, z, ny = _masks(n, letters)
        xs.append(x); zs.append(z); nys.append(ny); cs.append(c)
    return {"x": np.array(xs, dtype=np.uint64), "z": np.array(zs, dtype=np.uint64), "ny": np.array(nys, dtype=np.int32),
            "cre": np.array(cs), "n_groups": len(set(xs))}


def hf_index(n, n_electrons=None):
    """JW Hartree-Fock determinant: the first n_electrons spin-orbitals (qubits 0..) occupied."""
    ne = n // 2 if n_electrons is None else n_electrons
    return ((1 << ne) - 1) << (n - ne)


def to_terms(n, w, coeff_key):
    """Duck-typed (coeff, op, qbits) triples for the oracle / the reference-shaped API."""
    out = []
    for x, z, c in zip(w["x"], w["z"], w[coeff_key]):
        op, qb = [], []
        for q in range(n):
            b = n - 1 - q
            xb, zb = (int(x) >> b) & 1, (int(z) >> b) & 1
            if xb or zb:
                op.append("Y" if xb and zb else ("X" if xb else "Z"))
                qb.append(q)
        out.append((float(c), "".join(op), qb))
    return out
