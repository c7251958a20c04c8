"""CPU: the plain-C oracle (bench CPU baseline) against the numpy oracle."""
import numpy as np

from oracle import c_oracle
from oracle import statevector_oracle as orc
from openvqe_b200.lowering import pack_operator, pack_pool, term_masks
from tests.helpers import random_antihermitian, random_hermitian, random_pauli, random_state


def test_c_rotations_expectation_gates_match_numpy():
    n = 9
    rng = np.random.default_rng(3)
    psi = random_state(rng, n)
    ref = psi.copy()
    xs, zs, nys, angs = [], [], [], []
    for k in range(25):
        op, qb = random_pauli(rng, n, 5)
        if k % 5 == 0:
            op = "Z" * len(qb)
        x, z, ny = term_masks(op, qb, n)
        a = float(rng.uniform(-1, 1))
        xs.append(x); zs.append(z); nys.append(ny); angs.append(a)
        ref = orc.pauli_rotation(ref, x, z, ny, a)
    got = c_oracle.apply_rotations(psi.copy(), n, xs, zs, nys, angs)
    assert np.abs(got - ref).max() < 1e-13
    ham = random_hermitian(rng, n, 40, 6, const=0.5)
    p = pack_operator(ham, with_constant=True)
    assert abs(c_oracle.expectation(got, n, p.x, p.z, p.ny, p.cre, p.cim) - orc.expectation(ref, ham)) < 1e-12
    gates = [("H", [0], None), ("CNOT", [0, 3], None), ("RY", [3], 0.4), ("RZ", [8], -1.1), ("RX", [5], 2.0),
             ("X", [2], None), ("CNOT", [7, 1], None)]
    kinds = {"X": 0, "H": 1, "RX": 2, "RY": 3, "RZ": 4, "CNOT": 5}
    g2 = c_oracle.apply_gates(got.copy(), n, [kinds[g[0]] for g in gates], [g[1][0] for g in gates],
                              [g[1][1] if len(g[1]) > 1 else 0 for g in gates], [g[2] or 0.0 for g in gates])
    assert np.abs(g2 - orc.apply_gates(ref, n, gates)).max() < 1e-13
    sig = c_oracle.apply_paulisum(got, n, p)
    assert np.abs(sig - orc.apply_pauli_sum(ref, ham)).max() < 1e-12
    pool = [random_antihermitian(rng, n, 4, 4) for _ in range(7)]
    ov = c_oracle.pool_overlaps(sig, got, n, pack_pool(pool))
    ref_ov = np.array([np.vdot(sig, orc.apply_pauli_sum(ref, o)) for o in pool])
    assert np.abs(ov - ref_ov).max() < 1e-12
