"""Shared GPU implementations behind the reference-shaped entry points.

Every function here replaces the *body* of a reference hot-path function with
calls into the CUDA engine; the reference-shaped wrappers live in
``openvqe_b200/ucc_family`` and ``openvqe_b200/adapt``.
"""
from __future__ import annotations

import numpy as np

from .common_files.circuit import hf_index, quccsd_circuit, quccsd_plane_ops
from .engine import BUF_PSI, BUF_SIGMA, GATE_KINDS, get_engine, operator_fingerprint
from .lowering import PackedTerms, pack_operator, pack_pool

_PACK_CACHE = {}
_PACK_CACHE_MAX = 16384


def packed(op) -> PackedTerms:
    """Lowered form of one operator, cached per object (the same generator objects
    are passed on every one of the thousands of objective evaluations)."""
    key = id(op)
    mark = operator_fingerprint(op)
    hit = _PACK_CACHE.get(key)
    if hit is not None and hit[0] is op and hit[2] == mark:
        return hit[1]
    p = pack_operator(op)
    if len(_PACK_CACHE) >= _PACK_CACHE_MAX:
        _PACK_CACHE.clear()
    _PACK_CACHE[key] = (op, p, mark)  # strong ref: the id cannot be recycled while cached
    return p


class RotationProgram:
    """Concatenated Pauli strings of an ordered generator list: rotation k has
    angle theta[owner[k]] * creal[k]."""

    __slots__ = ("x", "z", "ny", "creal", "owner", "n_ops")

    def __init__(self, ops):
        xs, zs, nys, cr, own = [], [], [], [], []
        for j, op in enumerate(ops):
            p = packed(op)
            keep = (p.cre != 0) | (p.cim != 0)  # zero strings are exact identities
            xs.append(p.x[keep])
            zs.append(p.z[keep])
            nys.append(p.ny[keep])
            cr.append(p.cre[keep])
            own.append(np.full(int(keep.sum()), j, dtype=np.int64))
        cat = lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros(0, dt)
        self.x, self.z = cat(xs, np.uint64), cat(zs, np.uint64)
        self.ny, self.creal = cat(nys, np.int32), cat(cr, np.float64)
        self.owner = cat(own, np.int64)
        self.n_ops = len(ops)


_PROG_CACHE = {}


def _spot_marks(ops):
    """Fingerprints of up to 10 operators spread over the list (O(1) per call): operators are treated as immutable,
    as the reference treats them; this catches the common ways of breaking that (a rebuilt or rescaled generator)."""
    n = len(ops)
    if n == 0:
        return ()
    step = max(1, n // 9)
    return tuple(operator_fingerprint(ops[k]) for k in sorted(set(list(range(0, n, step))[:9] + [n - 1])))


def rotation_program(ops) -> RotationProgram:
    key = tuple(id(o) for o in ops)
    hit = _PROG_CACHE.get(key)
    if hit is not None and all(a is b for a, b in zip(hit[0], ops)) and hit[2] == _spot_marks(ops):
        return hit[1]
    prog = RotationProgram(ops)
    if len(_PROG_CACHE) >= 256:
        _PROG_CACHE.clear()
    _PROG_CACHE[key] = (list(ops), prog, _spot_marks(ops))
    return prog


def prepare_ucc_state(engine, cluster_ops_sp, hf_init_sp, theta):
    """|psi> = prod_j prod_k exp(-i theta_j Re(c_jk) P_jk) |HF> on the device.
    Replaces reference get_energy_ucc.py:42-46.  ``zip`` truncation: only the first
    min(len(ops), len(theta)) generators are applied; with no generator at all the
    reference never applies the HF X gates (``init`` is only passed for n_term == 0)."""
    m = min(len(cluster_ops_sp), len(theta))
    if m == 0:
        engine.set_basis_state(0)
        return
    ops = list(cluster_ops_sp[:m])
    prog = rotation_program(ops)
    th = np.asarray(theta, dtype=np.float64)[:m]
    angles = th[prog.owner] * prog.creal
    engine.set_basis_state(int(hf_init_sp))
    engine.apply_rotations(prog.x, prog.z, prog.ny, angles)


def ucc_energy(theta, hamiltonian_sp, cluster_ops_sp, hf_init_sp, device=None):
    """E(theta) of the Trotterised UCC ansatz (reference get_energy_ucc.py:8-50)."""
    engine = get_engine(hamiltonian_sp.nbqbits, device)
    prepare_ucc_state(engine, cluster_ops_sp, hf_init_sp, theta)
    return float(engine.expectation(engine.paulisum(hamiltonian_sp)).real)


def apply_gate_list(engine, gates):
    kinds = [GATE_KINDS[g[0]] for g in gates]
    q0 = [g[1][0] for g in gates]
    q1 = [g[1][1] if len(g[1]) > 1 else 0 for g in gates]
    ang = [0.0 if g[2] is None else g[2] for g in gates]
    engine.apply_gates(kinds, q0, q1, ang)


def prepare_quccsd_state(engine, n, hf_init_sp, cluster_ops, theta, use_tables=True, global_phase=True):
    """|psi> of the gate-defined QUCCSD ansatz (reference get_energy_qucc.py:38-51).  Every excitation template is
    applied as ONE tabulated plane rotation (the exact unitary of its gate list, see common_files/circuit.py); if
    some excitation is not of that form, or touches a global qubit of a sharded state, the gate list is executed.
    ``global_phase=False`` leaves out the accumulated phase e^{i pi/4 * n_singles} of the single-excitation templates: an
    expectation value does not see it, and without it the state stays purely real (real layout, real kernels)."""
    list_exci = [list(op.terms[0].qbits) for op in cluster_ops]
    if len(theta) < len(list_exci):
        raise IndexError("list index out of range")  # as the reference's list_theta[i] (circuit.py:95-106)
    ops = quccsd_plane_ops(n, list_exci, theta) if use_tables else None
    n_global = getattr(engine, "n_global", 0)
    if ops is not None and n_global:
        top = max(int(x).bit_length() for x in ops[0]) if len(ops[0]) else 0
        if top > n - n_global:
            ops = None
    if ops is None:
        circ = quccsd_circuit(n, hf_init_sp, cluster_ops, theta)
        engine.set_basis_state(0)
        apply_gate_list(engine, circ.gates)
        return
    x, offs, pat, cosv, sinv, phase = ops
    engine.set_basis_state(hf_index(n, hf_init_sp))
    engine.apply_plane_rotations(x, offs, pat, cosv, sinv)
    if phase != 1.0 and global_phase:
        engine.scale_state(phase)


def quccsd_energy(theta, hamiltonian_sp, cluster_ops, hf_init_sp, device=None):
    """E(theta) of the gate-defined QUCCSD ansatz (reference get_energy_qucc.py:11-56)."""
    n = hamiltonian_sp.nbqbits
    engine = get_engine(n, device)
    prepare_quccsd_state(engine, n, hf_init_sp, cluster_ops, theta, global_phase=False)
    return float(engine.expectation(engine.paulisum(hamiltonian_sp)).real)


def basis_energy(hamiltonian_sp, hf_init_sp, device=None):
    """<HF|H|HF> (reference hf_energy, fermionic_adapt_vqe.py:216-238)."""
    engine = get_engine(hamiltonian_sp.nbqbits, device)
    engine.set_basis_state(int(hf_init_sp))
    return float(engine.expectation(engine.paulisum(hamiltonian_sp)).real)


_MATLIST_CACHE = {}


def operators_of_matrices(mats):
    """Pauli-list operators of a list of 2^n x 2^n matrices (reference cluster_ops_sparse); the converted LIST is cached
    per list contents so that the pool lowering cache keyed on operator identity keeps hitting."""
    from .lowering import operator_from_matrix
    key = tuple(id(m) for m in mats)
    hit = _MATLIST_CACHE.get(key)
    if hit is not None and all(a is b for a, b in zip(hit[0], mats)):
        return hit[1]
    ops = [operator_from_matrix(m) for m in mats]
    if len(_MATLIST_CACHE) >= 16:
        _MATLIST_CACHE.clear()
    _MATLIST_CACHE[key] = (list(mats), ops)
    return ops


def load_reference_ket(engine, reference_ket):
    """Accept the reference's 2^n x 1 scipy-sparse / dense column or a flat vector."""
    if hasattr(reference_ket, "toarray"):
        reference_ket = reference_ket.toarray()
    engine.set_state(np.asarray(reference_ket, dtype=np.complex128).reshape(-1))


def pool_overlaps(engine, hamiltonian_sp, pool_ops):
    """sigma = H psi once, then <sigma|A_k|psi> for the whole pool in one sweep
    (reference fermionic_adapt_vqe.py:114-121 / qubit_adapt_vqe.py:462-472)."""
    engine.apply_paulisum(engine.paulisum(hamiltonian_sp), dst=BUF_SIGMA, src=BUF_PSI)
    key = ("pool",) + tuple(id(o) for o in pool_ops)
    hit = _PROG_CACHE.get(key)
    if hit is not None and all(a is b for a, b in zip(hit[0], pool_ops)) and hit[2] == _spot_marks(pool_ops):
        pool = hit[1]
    else:
        pool = pack_pool(pool_ops)
        if len(_PROG_CACHE) >= 256:
            _PROG_CACHE.clear()
        _PROG_CACHE[key] = (list(pool_ops), pool, _spot_marks(pool_ops))
    if replica_split_active(engine):
        from . import sharded
        return sharded.replica_pool_overlaps(engine, pool, bra=BUF_SIGMA, ket=BUF_PSI)
    return engine.pool_overlaps(pool, bra=BUF_SIGMA, ket=BUF_PSI)


def replica_split_active(engine) -> bool:
    """True when the caller has OPTED IN to SPMD replica mode (``sharded.enable_replica()`` or VQE_B200_REPLICA=1),
    this process is one of several ranks and the state fits one GPU: the pool sweep is then split over the ranks.
    Never on by default -- under torchrun ranks often work on different problems (a geometry scan, rank-0-only ADAPT),
    and splitting would then mix gradients of different Hamiltonians or deadlock.  VQE_B200_REPLICA_POOL=0 keeps the
    finite-difference split but not the pool split."""
    import os
    if os.environ.get("VQE_B200_REPLICA_POOL", "1") == "0" or getattr(engine, "n_global", 0):
        return False
    return replica_world(which=None)[1] > 1


ZERO_TOL = 1e-14  # Hartree; far below the rounding noise of a 2^n-term fp64 reduction of O(1) values


def snap_ties(values, rel=1e-12, zero_tol=ZERO_TOL):
    """Make the selection robust against reduction-order noise (documented deviation):

    * The reference drops EXACT zeros before sorting (sorted_gradient.py:32,53); in its scipy path
      symmetry-forbidden gradients are exactly 0.0 and a few more cancel exactly.  A parallel reduction
      may leave ~1e-17 there, so |g| <= ``zero_tol`` is snapped to 0.0.
    * Gradients that are equal in exact arithmetic (spin-complement partners carry +-g, SURVEY.md
      Appendix B item 14) may differ in the last bits.  The reference's selection uses float equality
      with ties resolved to the lowest pool index, so magnitudes within ``rel`` of each other are
      snapped to the value of their lowest-index member.  Signs are kept."""
    vals = [0.0 if abs(v) <= zero_tol else v for v in values]
    order = sorted(range(len(vals)), key=lambda k: -abs(vals[k]))
    i = 0
    while i < len(order):
        j = i + 1
        top = abs(vals[order[i]])
        while j < len(order) and top > 0 and top - abs(vals[order[j]]) <= rel * top:
            j += 1
        if j - i > 1:
            members = sorted(order[i:j])
            mag = abs(vals[members[0]])
            for k in members:
                vals[k] = mag if vals[k] >= 0 else -mag
        i = j
    return vals


def _strings_commute(x, z):
    """True when the Pauli strings (x_k, z_k) commute pairwise."""
    n = len(x)
    for a in range(n):
        for b in range(a + 1, n):
            if (bin(int(x[a]) & int(z[b])).count("1") + bin(int(z[a]) & int(x[b])).count("1")) & 1:
                return False
    return True


def ucc_energy_and_gradient(theta, hamiltonian_sp, cluster_ops_sp, hf_init_sp, device=None):
    """E(theta) and dE/dtheta of the Trotterised UCC ansatz by ADJOINT differentiation: one forward state preparation,
    the co-state lambda = H psi, then one reverse sweep that peels the generators off both vectors,

        dE/dtheta_j = 2 Re <lambda_j| (-i G_j) |psi_j> = 2 Im <lambda_j| G_j |psi_j>,
        psi_{j-1} = U_j^-1 psi_j,   lambda_{j-1} = U_j^-1 lambda_j,

    instead of the n+1 energy evaluations of the reference's finite differences (get_energy_ucc.py:158-166).  A
    generator whose strings do not commute is peeled string by string (the derivative of the ordered product).  Opt-in
    (SURVEY.md section 8f item 2): it changes the optimiser's trajectory at the level of the finite-difference error,
    so the drop-in drivers use it only when VQE_B200_ADJOINT=1."""
    from .lowering import PackedTerms
    engine = get_engine(hamiltonian_sp.nbqbits, device)
    n = hamiltonian_sp.nbqbits
    m = min(len(cluster_ops_sp), len(theta))
    th = np.asarray(theta, dtype=np.float64)
    grad = np.zeros(len(theta), dtype=np.float64)
    prepare_ucc_state(engine, cluster_ops_sp, hf_init_sp, theta)
    ps = engine.paulisum(hamiltonian_sp)
    energy = float(engine.expectation(ps).real)
    if m == 0:
        return energy, grad
    engine.apply_paulisum(ps, dst=BUF_SIGMA, src=BUF_PSI)
    for j in range(m - 1, -1, -1):
        p = packed(cluster_ops_sp[j])
        keep = (p.cre != 0) | (p.cim != 0)
        x, z, ny, c = p.x[keep], p.z[keep], p.ny[keep], p.cre[keep]
        if len(x) == 0:
            continue
        blocks = [slice(0, len(x))] if _strings_commute(x, z) else [slice(k, k + 1) for k in range(len(x) - 1, -1, -1)]
        for blk in blocks:
            bx, bz, bny, bc = x[blk], z[blk], ny[blk], c[blk]
            op = PackedTerms(n, bx, bz, bny, bc, np.zeros_like(bc), np.array([0, len(bx)], dtype=np.int32))
            ov = engine.pool_overlaps(op, bra=BUF_SIGMA, ket=BUF_PSI)[0]
            grad[j] += 2.0 * ov.imag
            # undo the block on both vectors: reverse order, negated angles
            ang = -(th[j] * bc)[::-1]
            rx, rz, rny = bx[::-1], bz[::-1], bny[::-1]
            engine.apply_rotations(rx, rz, rny, ang, buf=BUF_PSI)
            engine.apply_rotations(rx, rz, rny, ang, buf=BUF_SIGMA)
    return energy, grad


def adjoint_enabled() -> bool:
    import os
    return os.environ.get("VQE_B200_ADJOINT", "0") == "1"


def adjoint_fun_jac(hamiltonian_sp, cluster_ops_sp, hf_init_sp, energies):
    """(fun, jac) for scipy.optimize.minimize backed by ``ucc_energy_and_gradient`` (one adjoint sweep serves both)."""
    last = {}

    def both(theta):
        x = np.array(theta, dtype=np.float64)
        if "x" not in last or not np.array_equal(last["x"], x):
            last["x"] = x
            last["e"], last["g"] = ucc_energy_and_gradient(x, hamiltonian_sp, cluster_ops_sp, hf_init_sp)
        return last["e"], last["g"]

    def fun(theta):
        e = both(theta)[0]
        energies.append(e)
        return e

    return fun, lambda theta: both(theta)[1].copy()


FD_STEP = 1.4901161193847656e-08  # scipy's absolute 2-point step for BFGS with jac=None (sqrt of machine epsilon)


def replica_world(which="fd"):
    """(rank, world) of the SPMD replica group, or (0, 1) when replica mode is off.  Replica mode is OPT-IN:
    ``openvqe_b200.sharded.enable_replica(group)`` or the environment variable VQE_B200_REPLICA=1 (every rank must then
    run the same problem in lockstep)."""
    import os
    import sys
    from . import sharded
    if not (sharded.replica_enabled() or os.environ.get("VQE_B200_REPLICA", "0") == "1"):
        return 0, 1
    if which == "fd" and os.environ.get("VQE_B200_REPLICA_FD", "1") == "0":
        return 0, 1
    if "torch.distributed" not in sys.modules or not sharded.dist_ready():
        return 0, 1
    dist = sharded._dist()
    group = sharded.replica_group()
    return dist.get_rank(group), dist.get_world_size(group)


def distributed_fd(action, energies):
    """Finite-difference gradient of a BFGS run spread over the SPMD ranks.

    ``action(theta)`` is the energy WITHOUT the bookkeeping append.  Returns ``(fun, jac)`` for
    ``scipy.optimize.minimize``: ``fun`` appends to ``energies`` like the reference's objective; ``jac`` is scipy's own
    2-point forward difference -- same absolute step, same (x + h) - x denominator, evaluation order f(x),
    f(x + h e_0), f(x + h e_1), ... -- but every rank evaluates only its slice of the n displaced points on its own
    GPU and the values are all-gathered, so the gradient, the optimiser trajectory and the ``energies`` list are the
    ones of the serial run.  Outside a replica group ``jac`` is None and scipy differences serially."""
    rank, world = replica_world()
    last = {}

    def fun(theta):
        x = np.array(theta, dtype=np.float64)
        if "x" in last and last.get("pending") and np.array_equal(last["x"], x):
            last["pending"] = False      # already evaluated (and appended) by jac at this point
            return last["f"]
        v = action(x)
        energies.append(v)
        last["x"], last["f"], last["pending"] = x, v, False
        return v

    if world == 1:
        return fun, None
    from . import sharded
    group = sharded.replica_group()

    def jac(theta):
        x0 = np.array(theta, dtype=np.float64)
        if not ("x" in last and np.array_equal(last["x"], x0)):
            # scipy asked for the gradient at a point it has not evaluated yet: evaluate and append f(x) FIRST, as the
            # serial finite-difference path does, and let the following fun(x) reuse it without appending again
            fun(x0)
            last["pending"] = True
        f0 = last["f"]
        n = x0.shape[0]
        if not sharded.ranks_agree(x0.tobytes(), group):
            # the ranks are not running the same problem: every rank differences serially (same decision everywhere)
            vals = np.array([action(np.where(np.arange(n) == i, x0 + FD_STEP, x0)) for i in range(n)])
        else:
            lo, hi = sharded.split_range(n, world, rank)
            width = max(sharded.split_range(n, world, r)[1] - sharded.split_range(n, world, r)[0] for r in range(world))
            mine = np.zeros(width, dtype=np.float64)
            for i in range(lo, hi):
                xi = x0.copy()
                xi[i] = x0[i] + FD_STEP
                mine[i - lo] = action(xi)
            rows = sharded.allgather_f64(mine, group)
            vals = np.concatenate([rows[r, :sharded.split_range(n, world, r)[1] - sharded.split_range(n, world, r)[0]]
                                   for r in range(world)])
        energies.extend(float(v) for v in vals)
        dx = (x0 + FD_STEP) - x0
        return (vals - f0) / dx

    return fun, jac
