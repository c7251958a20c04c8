"""CPU checks of the multi-GPU host logic: the peer-pass planner (pure host code in the C-ABI library), the
rank-ordered reductions and the replica-mode pool split over a 2-rank gloo group (no GPU, no compute calls into
the CUDA library)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rot(n, op, qb):
    from openvqe_b200.lowering import term_masks
    return term_masks(op, qb, n)


def test_planner_local_vs_peer_passes():
    from openvqe_b200.sharded import plan_rotations
    n, g = 20, 2  # qubits 0,1 are global (index bits 19, 18); 18 local bits
    prog = [
        ("XY", [5, 9]),          # local
        ("ZXZY", [0, 7, 8, 9]),  # Z on a global qubit only: still local, no traffic
        ("XZY", [1, 3, 4]),      # X on global qubit 1 -> peer pass, pattern 0b01
        ("YZZX", [1, 2, 3, 6]),  # same pattern: fused into the same peer pass
        ("XX", [0, 1]),          # pattern 0b11: a new peer pass
        ("XY", [10, 11]),        # local again (absorbed by the open peer pass: it costs no extra sweep)
        ("YX", [0, 12]),         # pattern 0b10
    ]
    xs, zs, nys = zip(*[_rot(n, op, qb) for op, qb in prog])
    passes = plan_rotations(n, g, xs, zs, nys, [0.1] * len(prog))
    kinds = [p[0] for p in passes]
    pats = [p[1] for p in passes]
    assert sum(p[2] for p in passes) == len(prog)           # every rotation exactly once, order kept
    assert kinds == [1, 1, 1] and pats == [0b01, 0b11, 0b10]
    assert [p[2] for p in passes] == [4, 2, 1]
    nl = n - g
    for (kind, pat, nops, tmask), lo_hi in zip(passes, [(0, 4), (4, 6), (6, 7)]):
        need = 0
        for x in xs[lo_hi[0]:lo_hi[1]]:
            need |= x & ((1 << nl) - 1)
        assert need & ~tmask == 0                           # every local X bit is inside the tile
        assert bin(tmask).count("1") == 11 and tmask & 31 == 31   # 12-bit tile minus the virtual shard bit; low 5 bits fixed
    # without global qubits the same program only has local passes over 12-bit tiles
    one = plan_rotations(n, 0, xs, zs, nys, [0.1] * len(prog))
    assert all(p[0] == 0 and p[1] == 0 for p in one) and sum(p[2] for p in one) == len(prog)
    assert all(bin(p[3]).count("1") == 12 for p in one)


def test_planner_gather_form_for_collapsed_excitations():
    """JW excitations collapse to one plane rotation per generator; a peer pass made of them runs in gather form
    (kind 2), a peer pass with a lone Pauli string that flips a global qubit keeps the exchange form (kind 1)."""
    from openvqe_b200.lowering import pack_operator
    from openvqe_b200.sharded import plan_rotations
    from tests.helpers import jw_excitation
    n, g = 20, 2
    xs, zs, nys, angs = [], [], [], []
    def add(cre, ann):
        pk = pack_operator(jw_excitation(n, cre, ann))
        xs.extend(int(v) for v in pk.x); zs.extend(int(v) for v in pk.z); nys.extend(int(v) for v in pk.ny)
        angs.extend(0.1 * float(c) for c in pk.cre)

    for cre, ann in (([9, 12], [0, 5]), ([15], [1]), ([10, 11], [6, 7])):
        add(cre, ann)
    kinds = [p[0] for p in plan_rotations(n, g, xs, zs, nys, angs)]
    assert kinds and all(k in (0, 2) for k in kinds) and 2 in kinds
    # a generator that flips BOTH global qubits keeps per-string outside-tile signs: not collapsed, exchange form
    add([8, 19], [0, 1])
    assert plan_rotations(n, g, xs, zs, nys, angs)[-1][0] == 1
    # so does a lone Pauli string on a global qubit
    x, z, ny = _rot(n, "XZY", [1, 3, 4])
    kinds = [p[0] for p in plan_rotations(n, g, xs + [x], zs + [z], nys + [ny], angs + [0.3])]
    assert kinds[-1] == 1


def test_planner_zero_angles_dropped_and_capacity_split():
    from openvqe_b200.sharded import plan_rotations
    n = 24
    prog = [("X" * 1, [q]) for q in range(n)]
    xs, zs, nys = zip(*[_rot(n, op, qb) for op, qb in prog])
    angles = [0.2] * n
    angles[3] = 0.0
    passes = plan_rotations(n, 0, xs, zs, nys, angles)
    assert sum(p[2] for p in passes) == n - 1
    assert len(passes) >= 3          # 24 distinct X bits cannot share one 12-bit tile with 5 fixed low bits
    for p in passes:
        assert p[0] == 0


def test_split_range_and_rank_order_sum():
    from openvqe_b200.sharded import n_global_for, split_range, sum_in_rank_order
    for n_items in [0, 1, 7, 285, 3159]:
        for world in [1, 2, 4, 8]:
            cover = []
            for r in range(world):
                lo, hi = split_range(n_items, world, r)
                cover.extend(range(lo, hi))
            assert cover == list(range(n_items))
    assert [n_global_for(w) for w in (1, 2, 4, 8)] == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        n_global_for(6)
    rows = np.array([[1e16, 1.0], [1.0, 1.0], [-1e16, 1.0]])
    assert sum_in_rank_order(rows).tolist() == [0.0, 3.0]    # ((1e16 + 1) - 1e16): the fixed order is part of the contract


def test_gloo_world2_host_logic():
    """Two CPU processes over gloo: all-gather + rank-ordered sums give bit-identical totals on both ranks, and the
    replica-mode pool sweep (each rank evaluates its slice, slices are gathered) equals the unsplit sweep."""
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "gloo_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "gloo worker ok" in res.stdout


def test_quccsd_template_tables_match_the_gate_level_oracle():
    """Host logic of the QUCCSD fast path (no GPU): the (pattern, cos, sin) tables derived from the templates' own
    unitaries, applied with numpy, reproduce the oracle's gate-by-gate state -- ladders, both index layouts and the
    global phase of the single-excitation template included."""
    from openvqe_b200.common_files import circuit
    from oracle import statevector_oracle as orc
    from tests.helpers import FermiOp
    tabs = circuit.template_tables()
    assert tabs[2] is not None and tabs[4] is not None
    assert sorted(abs(c) for c in tabs[4][0]) == [0.5] * 7 + [4.5]      # the reference's double template: 7 pairs turn by theta/2, one by 4.5 theta
    n = 8
    rng = np.random.default_rng(3)
    ops = [FermiOp(n, [5, 7, 0, 2]), FermiOp(n, [6, 1]), FermiOp(n, [0, 3, 4, 7]), FermiOp(n, [2, 6]), FermiOp(n, [4, 5, 2, 3])]
    theta = rng.uniform(-0.9, 0.9, size=len(ops)).tolist()
    hf = 0b11110000
    x, offs, pat, cosv, sinv, phase = circuit.quccsd_plane_ops(n, [list(o.terms[0].qbits) for o in ops], theta)
    psi = orc.basis_state(n, circuit.hf_index(n, hf))
    idx = np.arange(1 << n)
    for k in range(len(x)):
        xm = int(x[k])
        new = psi.copy()
        for q in range(offs[k], offs[k + 1]):
            a_idx = idx[(idx & xm) == int(pat[q])]
            b_idx = a_idx ^ xm
            new[a_idx] = cosv[q] * psi[a_idx] - sinv[q] * psi[b_idx]
            new[b_idx] = sinv[q] * psi[a_idx] + cosv[q] * psi[b_idx]
        psi = new
    psi = psi * phase
    ref = orc.quccsd_state(n, hf, ops, theta)
    assert np.max(np.abs(psi - ref)) < 1e-13
    # a template whose ladder runs over one of its own core qubits is not tabulable
    assert circuit.quccsd_plane_ops(n, [[0, 3, 2, 4]], [0.1]) is None
