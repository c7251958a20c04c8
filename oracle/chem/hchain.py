"""TEST INFRASTRUCTURE ONLY -- fixture tooling, not product.

Closed-shell RHF for hydrogen chains / H2 in s-type Gaussian bases (STO-3G,
6-31G), from scratch in numpy.  It replaces the pyscf step of the reference's
input preparation (reference openvqe/common_files/molecule_factory.py:306-364,
``perform_pyscf_computation``), which cannot run here (no pyscf, no network), so
that molecular Hamiltonians of the shape the reference feeds into its hot path
can be produced for the golden fixtures.  SURVEY.md Appendix A (V1, V10)
records that this construction reproduces the 15-term H2/STO-3G Hamiltonian
stored in reference notebooks/demo_WSSVQE.ipynb cell[5] to <= 2e-14.

Outputs follow the conventions of the reference pipeline:
  spin-orbital index = 2*spatial + spin, MOs in energy order,
  H = sum h_pq C_p c_q + 1/2 sum (ps|qr) C_p C_q c_r c_s + E_nuc.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import erf

BOHR = 0.52917721092  # Angstrom, the value pyscf 2.7 uses

BASIS = {
    "sto-3g": [
        ([3.42525091, 0.62391373, 0.16885540], [0.15432897, 0.53532814, 0.44463454]),
    ],
    "6-31g": [
        ([18.7311370, 2.8253937, 0.6401217], [0.03349460, 0.23472695, 0.81375733]),
        ([0.1612778], [1.0]),
    ],
}


def _f0(t):
    t = np.asarray(t, dtype=float)
    out = np.empty_like(t)
    small = t < 1e-12
    ts = np.where(small, 1.0, t)
    out = 0.5 * np.sqrt(np.pi / ts) * erf(np.sqrt(ts))
    return np.where(small, 1.0 - t / 3.0, out)


def build_basis(coords_angstrom, basis):
    """-> primitive arrays (exp, coef*norm, center, contracted-function id)."""
    shells = BASIS[basis.lower()]
    exps, coefs, cents, owner = [], [], [], []
    nfun = 0
    for xyz in coords_angstrom:
        c = np.asarray(xyz, dtype=float) / BOHR
        for (es, cs) in shells:
            for e, k in zip(es, cs):
                exps.append(e)
                coefs.append(k * (2.0 * e / math.pi) ** 0.75)
                cents.append(c)
                owner.append(nfun)
            nfun += 1
    return np.array(exps), np.array(coefs), np.array(cents), np.array(owner), nfun


def integrals(coords_angstrom, basis, charges=None):
    exps, coefs, cents, owner, nfun = build_basis(coords_angstrom, basis)
    nuc = np.asarray(coords_angstrom, dtype=float) / BOHR
    charges = np.ones(len(nuc)) if charges is None else np.asarray(charges, float)
    npr = len(exps)
    a = exps[:, None]
    b = exps[None, :]
    p = a + b
    mu = a * b / p
    ab2 = ((cents[:, None, :] - cents[None, :, :]) ** 2).sum(-1)
    pc = (a[..., None] * cents[:, None, :] + b[..., None] * cents[None, :, :]) / p[..., None]
    kab = np.exp(-mu * ab2)
    s_prim = (np.pi / p) ** 1.5 * kab
    t_prim = mu * (3.0 - 2.0 * mu * ab2) * s_prim
    v_prim = np.zeros_like(s_prim)
    for z, c in zip(charges, nuc):
        r2 = ((pc - c) ** 2).sum(-1)
        v_prim += -z * 2.0 * np.pi / p * kab * _f0(p * r2)
    cc = coefs[:, None] * coefs[None, :]
    contr = np.zeros((nfun, npr))
    contr[owner, np.arange(npr)] = 1.0

    def contract2(m):
        return contr @ (m * cc) @ contr.T

    s = contract2(s_prim)
    hcore = contract2(t_prim + v_prim)
    # ERIs over primitive pairs
    pp = p.reshape(-1)
    kk = (kab * cc).reshape(-1)
    pcf = pc.reshape(-1, 3)
    pq2 = ((pcf[:, None, :] - pcf[None, :, :]) ** 2).sum(-1)
    psum = pp[:, None] + pp[None, :]
    rho = pp[:, None] * pp[None, :] / psum
    eri_prim = 2.0 * np.pi ** 2.5 / (pp[:, None] * pp[None, :] * np.sqrt(psum)) \
        * kk[:, None] * kk[None, :] * _f0(rho * pq2)
    pair = (contr[:, None, :, None] * contr[None, :, None, :]).reshape(nfun * nfun, npr * npr)
    eri = (pair @ eri_prim @ pair.T).reshape(nfun, nfun, nfun, nfun)  # chemists' (ab|cd)
    e_nuc = 0.0
    for i in range(len(nuc)):
        for j in range(i + 1, len(nuc)):
            e_nuc += charges[i] * charges[j] / np.linalg.norm(nuc[i] - nuc[j])
    return s, hcore, eri, e_nuc


def rhf(s, hcore, eri, n_elec, tol=1e-13, max_iter=500):
    nocc = n_elec // 2
    sval, svec = np.linalg.eigh(s)
    x = svec @ np.diag(sval ** -0.5) @ svec.T
    f = hcore.copy()
    d = np.zeros_like(s)
    e_old = 0.0
    fs, errs = [], []
    for it in range(max_iter):
        eps, c = np.linalg.eigh(x.T @ f @ x)
        c = x @ c
        d = 2.0 * c[:, :nocc] @ c[:, :nocc].T
        j = np.einsum("pqrs,rs->pq", eri, d)
        k = np.einsum("prqs,rs->pq", eri, d)
        f = hcore + j - 0.5 * k
        e = 0.5 * np.sum(d * (hcore + f))
        err = f @ d @ s - s @ d @ f
        fs.append(f.copy())
        errs.append(err)
        if len(fs) > 8:
            fs.pop(0)
            errs.pop(0)
        if len(fs) >= 2:  # DIIS
            m = len(fs)
            bm = -np.ones((m + 1, m + 1))
            bm[m, m] = 0.0
            for a_ in range(m):
                for b_ in range(m):
                    bm[a_, b_] = np.sum(errs[a_] * errs[b_])
            rhs = np.zeros(m + 1)
            rhs[m] = -1.0
            try:
                w = np.linalg.solve(bm, rhs)[:m]
                f = sum(wi * fi for wi, fi in zip(w, fs))
            except np.linalg.LinAlgError:
                pass
        if abs(e - e_old) < tol and np.abs(err).max() < 1e-10:
            break
        e_old = e
    eps, c = np.linalg.eigh(x.T @ fs[-1] @ x) if False else np.linalg.eigh(x.T @ (hcore + j - 0.5 * k) @ x)
    c = x @ c
    # fix the arbitrary MO signs deterministically: largest-|coef| entry positive
    for m_ in range(c.shape[1]):
        i_ = np.argmax(np.abs(c[:, m_]))
        if c[i_, m_] < 0:
            c[:, m_] *= -1.0
    return e, eps, c


def molecular_integrals(coords_angstrom, basis, n_elec=None):
    """-> dict with MO one/two-body integrals in the layout the reference's
    ``convert_to_h_integrals`` consumes: two_body[p,q,r,s] = (p s | q r)."""
    n_elec = len(coords_angstrom) if n_elec is None else n_elec
    s, hcore, eri, e_nuc = integrals(coords_angstrom, basis)
    e_el, eps, c = rhf(s, hcore, eri, n_elec)
    h1 = c.T @ hcore @ c
    mo = np.einsum("abcd,ap,bq,cr,ds->pqrs", eri, c, c, c, c, optimize=True)  # (pq|rs)
    two_body = np.transpose(mo, (0, 2, 3, 1))  # I[p,q,r,s] = (p s | q r)
    return {
        "one_body": h1,
        "two_body": two_body,
        "nuclear_repulsion": e_nuc,
        "orbital_energies": eps,
        "hf_energy": e_el + e_nuc,
        "n_elec": n_elec,
    }


def chain(n_atoms, r_angstrom):
    return [(0.0, 0.0, i * r_angstrom) for i in range(n_atoms)]
