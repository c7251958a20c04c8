"""TEST INFRASTRUCTURE ONLY -- stand-in for ``qat.fermion.chemistry.ucc``."""
import numpy as np


def convert_to_h_integrals(one_body_integrals, two_body_integrals):
    """Spatial MO integrals -> spin-orbital hpq, hpqrs with
    H = sum hpq C_p c_q + 1/2 sum hpqrs C_p C_q c_r c_s.
    ``two_body_integrals[p,q,r,s]`` is taken in the ordering pyscf-derived myQLM
    uses, I[p,q,r,s] = (p s | q r) (chemists' (ps|qr)); spin-orbital index =
    2*spatial + spin."""
    h1 = np.asarray(one_body_integrals)
    h2 = np.asarray(two_body_integrals)
    m = h1.shape[0]
    n = 2 * m
    hpq = np.zeros((n, n))
    hpqrs = np.zeros((n, n, n, n))
    for s in range(2):
        hpq[s::2, s::2] = h1
    for s1 in range(2):
        for s2 in range(2):
            # C_{p s1} C_{q s2} c_{r s2} c_{s s1}
            hpqrs[s1::2, s2::2, s2::2, s1::2] = h2
    return hpq, hpqrs


def transform_integrals_to_new_basis(one_body_integrals, two_body_integrals, U):
    h1 = U.T @ one_body_integrals @ U
    h2 = np.einsum("pqrs,pa,qb,rc,sd->abcd", two_body_integrals, U, U, U, U, optimize=True)
    return h1, h2
