"""TEST INFRASTRUCTURE ONLY -- fixture tooling, not product.

Closed-shell RHF over contracted Cartesian Gaussians with s AND p shells (McMurchie-Davidson scheme: Hermite
expansion coefficients E_t^{ij}, Hermite Coulomb integrals R_tuv, Boys function), from scratch in numpy.  It extends
oracle/chem/hchain.py (s functions only) so that the molecules BASELINE.json names -- LiH/STO-3G (reference
openvqe/common_files/molecule_factory_with_sparse.py:63-68) and H2O in 6-31G with the O 1s core frozen (reference
molecule_factory.py:138-148 geometry, :366-394 active space) -- can be turned into Pauli-list fixtures without pyscf.

Validation (tests/test_chem_fixture_cpu.py): the s-only integrals reproduce hchain.py to 1e-12; H2O/STO-3G at the
textbook geometry of the "Crawford programming projects" (R = 1.1 A, 104 deg) gives E_nuc = 8.002367061810 and
E_SCF = -74.942079928192; LiH/STO-3G at R = 1.6 A gives E_SCF = -7.8618 (literature, 4 decimals).

Conventions as hchain.py: spin-orbital index = 2*spatial + spin, MOs in energy order, two_body[p,q,r,s] = (p s|q r).
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import hyp1f1

from .hchain import BOHR, rhf

# shells: (kind, exponents, s-coefficients, p-coefficients); kind "s" or "sp"
_STO3G_S = [0.15432897, 0.53532814, 0.44463454]
_STO3G_2S = [-0.09996723, 0.39951283, 0.70011547]
_STO3G_2P = [0.15591627, 0.60768372, 0.39195739]
BASIS = {
    "sto-3g": {
        "H": [("s", [3.42525091, 0.62391373, 0.16885540], _STO3G_S, None)],
        "Li": [("s", [16.1195750, 2.9362007, 0.7946505], _STO3G_S, None),
               ("sp", [0.6362897, 0.1478601, 0.0480887], _STO3G_2S, _STO3G_2P)],
        "O": [("s", [130.7093200, 23.8088610, 6.4436083], _STO3G_S, None),
              ("sp", [5.0331513, 1.1695961, 0.3803890], _STO3G_2S, _STO3G_2P)],
    },
    "6-31g": {
        "H": [("s", [18.7311370, 2.8253937, 0.6401217], [0.03349460, 0.23472695, 0.81375733], None),
              ("s", [0.1612778], [1.0], None)],
        "O": [("s", [5484.6717000, 825.2349500, 188.0469600, 52.9645000, 16.8975700, 5.7996353],
               [0.0018311, 0.0139501, 0.0684451, 0.2327143, 0.4701930, 0.3585209], None),
              ("sp", [15.5396160, 3.5999336, 1.0137618], [-0.1107775, -0.1480263, 1.1307670],
               [0.0708743, 0.3397528, 0.7271586]),
              ("sp", [0.2700058], [1.0], [1.0])],
    },
}
CHARGE = {"H": 1, "Li": 3, "O": 8}


class BasisFunction:
    __slots__ = ("center", "lmn", "exps", "coefs")

    def __init__(self, center, lmn, exps, coefs):
        self.center = np.asarray(center, dtype=float)
        self.lmn = tuple(lmn)
        self.exps = np.asarray(exps, dtype=float)
        l = sum(lmn)
        # coefficients refer to normalised primitives; the contraction is then renormalised to unit self-overlap
        norm = (2.0 * self.exps / math.pi) ** 0.75 * (4.0 * self.exps) ** (l / 2.0)
        c = np.asarray(coefs, dtype=float) * norm
        a, b = self.exps[:, None], self.exps[None, :]
        pref = (math.pi / (a + b)) ** 1.5 * (1.0 if l == 0 else 1.0 / (2.0 * (a + b)))
        self.coefs = c / math.sqrt(float(c @ pref @ c))


def build_basis(geometry, basis):
    """geometry: [(symbol, (x, y, z) in Angstrom)] -> list of BasisFunction (atom order, s before p, px py pz)."""
    table = BASIS[basis.lower()]
    funcs = []
    for sym, xyz in geometry:
        key = sym.capitalize()
        c = np.asarray(xyz, dtype=float) / BOHR
        for kind, exps, cs, cp in table[key]:
            funcs.append(BasisFunction(c, (0, 0, 0), exps, cs))
            if kind == "sp":
                for lmn in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
                    funcs.append(BasisFunction(c, lmn, exps, cp))
    return funcs


def _E(i, j, t, qx, a, b):
    """Hermite expansion coefficient E_t^{ij} of the Gaussian product (one Cartesian direction)."""
    p = a + b
    q = a * b / p
    if t < 0 or t > i + j:
        return 0.0
    if i == j == t == 0:
        return math.exp(-q * qx * qx)
    if j == 0:
        return (1.0 / (2.0 * p)) * _E(i - 1, j, t - 1, qx, a, b) - (q * qx / a) * _E(i - 1, j, t, qx, a, b) \
            + (t + 1) * _E(i - 1, j, t + 1, qx, a, b)
    return (1.0 / (2.0 * p)) * _E(i, j - 1, t - 1, qx, a, b) + (q * qx / b) * _E(i, j - 1, t, qx, a, b) \
        + (t + 1) * _E(i, j - 1, t + 1, qx, a, b)


def _boys(n, x):
    return hyp1f1(n + 0.5, n + 1.5, -x) / (2.0 * n + 1.0)


def _R(tmax, umax, vmax, p, pc):
    """Hermite Coulomb integrals R^0_{tuv} for 0 <= t <= tmax etc. (bottom-up over the auxiliary index n)."""
    nmax = tmax + umax + vmax
    r2 = float(pc @ pc)
    # tab[n][t][u][v]
    tab = np.zeros((nmax + 2, tmax + 1, umax + 1, vmax + 1))
    for n in range(nmax + 1):
        tab[n, 0, 0, 0] = (-2.0 * p) ** n * _boys(n, p * r2)
    for n in range(nmax - 1, -1, -1):
        for t in range(tmax + 1):
            for u in range(umax + 1):
                for v in range(vmax + 1):
                    if t == u == v == 0 or t + u + v + n > nmax:
                        continue
                    if t > 0:
                        val = pc[0] * tab[n + 1, t - 1, u, v] + ((t - 1) * tab[n + 1, t - 2, u, v] if t > 1 else 0.0)
                    elif u > 0:
                        val = pc[1] * tab[n + 1, t, u - 1, v] + ((u - 1) * tab[n + 1, t, u - 2, v] if u > 1 else 0.0)
                    else:
                        val = pc[2] * tab[n + 1, t, u, v - 1] + ((v - 1) * tab[n + 1, t, u, v - 2] if v > 1 else 0.0)
                    tab[n, t, u, v] = val
    return tab[0]


def _overlap_prim(a, l1, A, b, l2, B):
    p = a + b
    s = (math.pi / p) ** 1.5
    for d in range(3):
        s *= _E(l1[d], l2[d], 0, A[d] - B[d], a, b)
    return s


def _kinetic_prim(a, l1, A, b, l2, B):
    l, m, n = l2
    t0 = b * (2 * (l + m + n) + 3) * _overlap_prim(a, l1, A, b, l2, B)
    t1 = -2.0 * b * b * (_overlap_prim(a, l1, A, b, (l + 2, m, n), B) + _overlap_prim(a, l1, A, b, (l, m + 2, n), B)
                         + _overlap_prim(a, l1, A, b, (l, m, n + 2), B))
    t2 = -0.5 * ((l * (l - 1) * _overlap_prim(a, l1, A, b, (l - 2, m, n), B) if l > 1 else 0.0)
                 + (m * (m - 1) * _overlap_prim(a, l1, A, b, (l, m - 2, n), B) if m > 1 else 0.0)
                 + (n * (n - 1) * _overlap_prim(a, l1, A, b, (l, m, n - 2), B) if n > 1 else 0.0))
    return t0 + t1 + t2


def _nuclear_prim(a, l1, A, b, l2, B, C):
    p = a + b
    P = (a * A + b * B) / p
    R = _R(l1[0] + l2[0], l1[1] + l2[1], l1[2] + l2[2], p, P - C)
    val = 0.0
    for t in range(l1[0] + l2[0] + 1):
        et = _E(l1[0], l2[0], t, A[0] - B[0], a, b)
        for u in range(l1[1] + l2[1] + 1):
            eu = _E(l1[1], l2[1], u, A[1] - B[1], a, b)
            for v in range(l1[2] + l2[2] + 1):
                val += et * eu * _E(l1[2], l2[2], v, A[2] - B[2], a, b) * R[t, u, v]
    return 2.0 * math.pi / p * val


class _PairData:
    """Hermite expansion of every primitive pair of two contracted functions (reused by all ERIs of the pair)."""

    def __init__(self, f1, f2):
        self.L = tuple(f1.lmn[d] + f2.lmn[d] for d in range(3))
        self.items = []  # (p, P, weight, E[t,u,v])
        A, B = f1.center, f2.center
        for a, ca in zip(f1.exps, f1.coefs):
            for b, cb in zip(f2.exps, f2.coefs):
                p = a + b
                P = (a * A + b * B) / p
                e = np.zeros((self.L[0] + 1, self.L[1] + 1, self.L[2] + 1))
                ex = [_E(f1.lmn[0], f2.lmn[0], t, A[0] - B[0], a, b) for t in range(self.L[0] + 1)]
                ey = [_E(f1.lmn[1], f2.lmn[1], t, A[1] - B[1], a, b) for t in range(self.L[1] + 1)]
                ez = [_E(f1.lmn[2], f2.lmn[2], t, A[2] - B[2], a, b) for t in range(self.L[2] + 1)]
                for t in range(self.L[0] + 1):
                    for u in range(self.L[1] + 1):
                        for v in range(self.L[2] + 1):
                            e[t, u, v] = ex[t] * ey[u] * ez[v]
                self.items.append((p, P, ca * cb, e))


def _eri(pab, pcd):
    La, Lc = pab.L, pcd.L
    sign = np.zeros((Lc[0] + 1, Lc[1] + 1, Lc[2] + 1))
    for t in range(Lc[0] + 1):
        for u in range(Lc[1] + 1):
            for v in range(Lc[2] + 1):
                sign[t, u, v] = (-1.0) ** (t + u + v)
    total = 0.0
    for p, P, wab, eab in pab.items:
        for q, Q, wcd, ecd in pcd.items:
            alpha = p * q / (p + q)
            R = _R(La[0] + Lc[0], La[1] + Lc[1], La[2] + Lc[2], alpha, P - Q)
            ecs = ecd * sign
            val = 0.0
            for t in range(La[0] + 1):
                for u in range(La[1] + 1):
                    for v in range(La[2] + 1):
                        if eab[t, u, v] == 0.0:
                            continue
                        val += eab[t, u, v] * float(np.sum(ecs * R[t:t + Lc[0] + 1, u:u + Lc[1] + 1, v:v + Lc[2] + 1]))
            total += wab * wcd * val * 2.0 * math.pi ** 2.5 / (p * q * math.sqrt(p + q))
    return total


def integrals(geometry, basis):
    """-> overlap, core Hamiltonian, ERIs (chemists' (ab|cd)), nuclear repulsion; all in the AO basis."""
    funcs = build_basis(geometry, basis)
    n = len(funcs)
    nuc = [(CHARGE[s.capitalize()], np.asarray(x, dtype=float) / BOHR) for s, x in geometry]
    s = np.zeros((n, n))
    h = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            fi, fj = funcs[i], funcs[j]
            sv = tv = vv = 0.0
            for a, ca in zip(fi.exps, fi.coefs):
                for b, cb in zip(fj.exps, fj.coefs):
                    w = ca * cb
                    sv += w * _overlap_prim(a, fi.lmn, fi.center, b, fj.lmn, fj.center)
                    tv += w * _kinetic_prim(a, fi.lmn, fi.center, b, fj.lmn, fj.center)
                    for z, c in nuc:
                        vv -= w * z * _nuclear_prim(a, fi.lmn, fi.center, b, fj.lmn, fj.center, c)
            s[i, j] = s[j, i] = sv
            h[i, j] = h[j, i] = tv + vv
    pairs = {(i, j): _PairData(funcs[i], funcs[j]) for i in range(n) for j in range(i + 1)}
    eri = np.zeros((n, n, n, n))
    for i in range(n):
        for j in range(i + 1):
            ij = i * (i + 1) // 2 + j
            for k in range(n):
                for l in range(k + 1):
                    if k * (k + 1) // 2 + l > ij:
                        continue
                    v = _eri(pairs[(i, j)], pairs[(k, l)])
                    for a, b, c, d in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k),
                                       (k, l, i, j), (l, k, i, j), (k, l, j, i), (l, k, j, i)):
                        eri[a, b, c, d] = v
    e_nuc = 0.0
    for a in range(len(nuc)):
        for b in range(a + 1, len(nuc)):
            e_nuc += nuc[a][0] * nuc[b][0] / float(np.linalg.norm(nuc[a][1] - nuc[b][1]))
    return s, h, eri, e_nuc


def molecular_integrals(geometry, basis, charge=0, n_frozen=0):
    """RHF + MO integrals in the layout of hchain.molecular_integrals.  ``n_frozen`` lowest MOs are frozen (doubly
    occupied core): their mean field is folded into the one-body integrals and their energy into the constant, and
    the returned integrals cover only the remaining orbitals (reference molecule_factory.py:366-394 does the same
    through get_active_space_hamiltonian with the core selected by natural-orbital occupation)."""
    n_elec = sum(CHARGE[s.capitalize()] for s, _ in geometry) - charge
    s, hcore, eri, e_nuc = integrals(geometry, basis)
    e_el, eps, c = rhf(s, hcore, eri, n_elec)
    h1 = c.T @ hcore @ c
    mo = np.einsum("abcd,ap,bq,cr,ds->pqrs", eri, c, c, c, c, optimize=True)  # (pq|rs)
    e_core = 0.0
    if n_frozen:
        fz = list(range(n_frozen))
        act = list(range(n_frozen, h1.shape[0]))
        e_core = 2.0 * sum(h1[i, i] for i in fz) + sum(2.0 * mo[i, i, j, j] - mo[i, j, j, i] for i in fz for j in fz)
        h_eff = h1.copy()
        for i in fz:
            h_eff += 2.0 * mo[:, :, i, i] - mo[:, i, i, :]
        h1 = h_eff[np.ix_(act, act)]
        mo = mo[np.ix_(act, act, act, act)]
        eps = eps[n_frozen:]
        n_elec -= 2 * n_frozen
    return {"one_body": h1, "two_body": np.transpose(mo, (0, 2, 3, 1)), "nuclear_repulsion": e_nuc + e_core,
            "orbital_energies": eps, "hf_energy": e_el + e_nuc, "n_elec": n_elec, "core_energy": e_core}
