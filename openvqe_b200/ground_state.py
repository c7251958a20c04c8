"""Lanczos ground state on the device.

Replaces ``np.linalg.eigh(hamiltonian_sp.get_matrix())`` (reference adapt/fermionic_adapt_vqe.py:474), which is
O(8^n) and stops near 14 qubits, by a two-pass Lanczos iteration whose only large operations are the engine's own
kernels: sigma = H v (``vqe_apply_paulisum``), <a|b> (``vqe_inner``) and a x + b y (``vqe_axpby``).  Three vectors
rotate through the context's buffers (psi, sigma, work); the second pass regenerates the Krylov vectors from the
stored recurrence coefficients and accumulates the Ritz vector into the rank-local aux buffer, where it stays for
``fidelity``.  No vector ever leaves the GPU.

The Krylov space of a computational basis state |HF> stays inside that state's symmetry sector (particle number,
S_z, point group), so the result is the lowest eigenpair of that sector -- the state the reference's fidelity is
meant to track.  (``eigh`` over the whole Fock space may return a lower eigenvalue of another particle-number
sector, or an arbitrary vector of a degenerate level; up to 14 qubits the drop-in therefore keeps the host ``eigh``
so that it reproduces the reference's numbers bit for bit, and uses this module above.)
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg

from .engine import BUF_AUX, BUF_PSI, BUF_SIGMA, BUF_WORK


class DeviceGroundState:
    """Handle of a ground state kept in the engine's aux buffer."""

    def __init__(self, engine, energy, iterations, residual):
        self.engine = engine
        self.energy = float(energy)
        self.iterations = int(iterations)
        self.residual = float(residual)

    def fidelity(self, buf=BUF_PSI) -> float:
        """|<gs|buf>|^2 (reference fun_fidelity, fermionic_adapt_vqe.py:331-361), reduced on the device."""
        return abs(self.engine.inner(BUF_AUX, buf)) ** 2

    def vector(self):
        return self.engine.get_state(BUF_AUX)


def _lowest_ritz(alphas, betas):
    if len(alphas) == 1:
        return alphas[0], np.ones(1)
    w, v = scipy.linalg.eigh_tridiagonal(np.asarray(alphas), np.asarray(betas[:len(alphas) - 1]), select="i",
                                         select_range=(0, 0))
    return float(w[0]), v[:, 0]


def lanczos_ground_state(engine, hamiltonian, start_index: int, tol: float = 1e-10, max_iter: int = 500,
                         min_iter: int = 4) -> DeviceGroundState:
    """Lowest eigenpair of ``hamiltonian`` (a Pauli-sum operator or an engine ``PauliSum``) in the Krylov space of
    the basis state ``start_index``.  ``tol`` bounds the residual norm |H y - E y| of the returned unit vector.
    Overwrites the psi, sigma and work buffers; the eigenvector is left in the aux buffer."""
    if getattr(engine, "n_global", 0):
        raise NotImplementedError("Lanczos ground state: not available on a sharded state (the aux vector is rank-local "
                                  "and three peer-attached vectors per rank do not fit above 33 qubits)")
    ps = hamiltonian if hasattr(hamiltonian, "handle") or hasattr(hamiltonian, "packed") else engine.paulisum(hamiltonian)

    def recurrence(accumulate=None):
        """One pass of the three-term recurrence.  accumulate = None: build (alphas, betas) until the lowest Ritz pair
        has converged; else the Ritz coefficients: regenerate v_j from the stored coefficients and add s_j v_j to aux."""
        cur, nxt, prev = BUF_PSI, BUF_SIGMA, BUF_WORK
        engine.set_basis_state(int(start_index))
        a_list, b_list = [], []
        theta, resid = float("nan"), float("inf")
        steps = max_iter if accumulate is None else len(accumulate)
        for j in range(steps):
            if accumulate is not None:
                engine.axpby(BUF_AUX, cur, accumulate[j], 0.0 if j == 0 else 1.0)
                if j + 1 == steps:
                    break
            engine.apply_paulisum(ps, dst=nxt, src=cur)                      # w = H v_j
            alpha = engine.inner(cur, nxt).real if accumulate is None else alphas[j]
            engine.axpby(nxt, cur, -alpha, 1.0)                               # w -= alpha_j v_j
            if j > 0:
                engine.axpby(nxt, prev, -(b_list[-1] if accumulate is None else betas[j - 1]), 1.0)  # w -= beta_j v_{j-1}
            if accumulate is None:
                a_list.append(alpha)
                beta = math.sqrt(max(engine.norm2(nxt), 0.0))
                theta, s = _lowest_ritz(a_list, b_list)
                resid = abs(beta * s[-1])
                if (j + 1 >= min_iter and resid < tol) or beta < 1e-14:
                    break
                b_list.append(beta)
            else:
                beta = betas[j]
            engine.scale_state(1.0 / beta, buf=nxt)
            prev, cur, nxt = cur, nxt, prev
        return a_list, b_list, theta, resid

    alphas, betas, theta, resid = recurrence()
    _, s = _lowest_ritz(alphas, betas)
    recurrence(accumulate=s)
    nrm = math.sqrt(engine.norm2(BUF_AUX))
    engine.scale_state(1.0 / nrm, buf=BUF_AUX)
    return DeviceGroundState(engine, theta, len(alphas), resid)
