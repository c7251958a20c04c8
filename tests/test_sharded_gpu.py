"""GPU parity of the SHARDED path against the CPU oracle.

A ``ShardGroup`` with several virtual ranks on ONE device runs exactly the kernels, the peer-pass planner and the
lo/hi tile split of the multi-GPU path (the partner's shard is then simply another allocation on the same GPU), so
these tests run on the 1-GPU box.  The IPC + device-flag-barrier plumbing of the one-process-per-GPU mode is covered
by ``test_two_process_ipc`` (two processes; uses two GPUs when present)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import statevector_oracle as orc
from tests.helpers import Ham, T, random_antihermitian, random_hermitian, random_pauli, random_state

pytestmark = pytest.mark.gpu
TOL = 1e-12
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def groups(gpu_required):
    from openvqe_b200.sharded import ShardGroup
    cache = {}

    def get(n, g):
        if (n, g) not in cache:
            cache[(n, g)] = ShardGroup(n, g)
        return cache[(n, g)]
    return get


@pytest.fixture(params=["relabel", "peer-passes"], autouse=True)
def relabel_mode(request, monkeypatch):
    """Every test runs twice: with qubit relabelling (rotations that flip a global qubit trigger a global<->local swap, all
    passes stay local; VQE_RELABEL_FLOOR lowered so that small shards take part) and without (peer passes)."""
    monkeypatch.setenv("VQE_RELABEL", "1" if request.param == "relabel" else "0")
    monkeypatch.setenv("VQE_RELABEL_FLOOR", "4")
    return request.param


CASES = [(4, 1), (6, 2), (9, 1), (9, 3), (13, 1), (13, 2), (14, 3), (16, 2), (18, 1)]


@pytest.mark.parametrize("n,g", CASES)
def test_sharded_rotations_match_oracle(groups, n, g):
    """Random Pauli rotations over ALL qubits: local passes, Z-on-global signs, peer passes for every pattern."""
    from openvqe_b200.lowering import term_masks
    rng = np.random.default_rng(1000 + 10 * n + g)
    grp = groups(n, g)
    psi = random_state(rng, n)
    grp.set_state(psi)
    xs, zs, nys, angs = [], [], [], []
    ref = psi.copy()
    for k in range(48):
        op, qb = random_pauli(rng, n, max_weight=min(n, 6))
        if k % 7 == 3:
            op = "Z" * len(qb)
        if k % 5 == 1:  # make sure the global qubits see X/Y letters often
            qb = sorted(set(qb) | {int(rng.integers(g))})
            op = "".join(rng.choice(list("XYZ"), size=len(qb)))
        x, z, ny = term_masks(op, qb, n)
        a = float(rng.uniform(-1, 1)) if k % 3 else float(rng.uniform(-0.2, 0.2))
        xs.append(x); zs.append(z); nys.append(ny); angs.append(a)
        ref = orc.pauli_rotation(ref, x, z, ny, a)
    grp.apply_rotations(xs, zs, nys, angs)
    got = grp.get_state()
    assert np.max(np.abs(got - ref)) < TOL
    assert abs(grp.norm2() - 1.0) < 1e-12


@pytest.mark.parametrize("n,g", [(10, 1), (12, 2), (14, 3)])
def test_sharded_ucc_like_program_equals_single_gpu(groups, n, g):
    """JW singles/doubles (8 strings sharing one X-mask -> register-resident runs), small angles (tangent fast path):
    the sharded state must equal the single-context state bit-for-bit up to fp64 reordering."""
    from openvqe_b200.engine import Engine
    from tests.helpers import jw_excitation
    from openvqe_b200.lowering import pack_operator
    rng = np.random.default_rng(77 + n)
    xs, zs, nys, angs = [], [], [], []
    for _ in range(30):
        if rng.random() < 0.3:
            p, q = sorted(rng.choice(n, size=2, replace=False).tolist())
            op = jw_excitation(n, [q], [p])
        else:
            p, q, r, s = sorted(rng.choice(n, size=4, replace=False).tolist())
            op = jw_excitation(n, [r, s], [p, q])
        pk = pack_operator(op)
        th = float(rng.uniform(-0.1, 0.1))
        for k in range(len(pk)):
            if pk.cim[k] == 0 and pk.cre[k] == 0:
                continue
            xs.append(int(pk.x[k])); zs.append(int(pk.z[k])); nys.append(int(pk.ny[k]))
            angs.append(th * float(pk.cim[k] if pk.cre[k] == 0 else pk.cre[k]))
    hf = ((1 << (n // 2)) - 1) << (n - n // 2)
    grp = groups(n, g)
    grp.set_basis_state(hf)
    grp.apply_rotations(xs, zs, nys, angs)
    eng = Engine(n)
    eng.set_basis_state(hf)
    eng.apply_rotations(xs, zs, nys, angs)
    a, b = grp.get_state(), eng.get_state()
    assert np.max(np.abs(a - b)) < 1e-14
    ref = orc.basis_state(n, hf)
    for x, z, ny, t in zip(xs, zs, nys, angs):
        ref = orc.pauli_rotation(ref, x, z, ny, t)
    assert np.max(np.abs(a - ref)) < TOL
    # structural zeros: the unsharded context collapses every same-X-mask run into one plane rotation and leaves the
    # untouched occupation patterns exactly 0.0; peer passes with two global X bits apply the strings one by one,
    # which leaves rounding residue ~1e-19 * |amplitude| there (the host snaps those, _hotpath.snap_ties)
    assert np.max(np.abs(a[b == 0])) < 1e-15


@pytest.mark.parametrize("n,g", [(5, 1), (9, 2), (12, 3), (14, 1)])
def test_sharded_gates_match_oracle(groups, n, g):
    from openvqe_b200.engine import GATE_KINDS
    rng = np.random.default_rng(2000 + n)
    grp = groups(n, g)
    psi = random_state(rng, n)
    grp.set_state(psi)
    gates = []
    for k in range(80):
        name = str(rng.choice(["X", "H", "RX", "RY", "RZ", "CNOT"]))
        lowq = lambda: int(rng.integers(g)) if k % 2 else int(rng.integers(n))  # global qubits often
        if name == "CNOT":
            c = lowq()
            t = int(rng.integers(n))
            if t == c:
                t = (c + 1) % n
            if k % 3 == 0:
                c, t = t, c
            gates.append(("CNOT", [c, t], None))
        else:
            gates.append((name, [lowq()], float(rng.uniform(-3, 3))))
    ref = orc.apply_gates(psi, n, gates)
    grp.apply_gates([GATE_KINDS[x[0]] for x in gates], [x[1][0] for x in gates],
                    [x[1][1] if len(x[1]) > 1 else 0 for x in gates], [x[2] or 0.0 for x in gates])
    assert np.max(np.abs(grp.get_state() - ref)) < TOL


@pytest.mark.parametrize("n,g,nterms", [(4, 1, 12), (8, 2, 120), (12, 3, 400), (13, 1, 300), (16, 2, 150)])
def test_sharded_expectation_matches_oracle(groups, n, g, nterms):
    rng = np.random.default_rng(3000 + n)
    grp = groups(n, g)
    psi = random_state(rng, n)
    grp.set_state(psi)
    ham = random_hermitian(rng, n, nterms, max_weight=min(n, 8), const=0.37)
    got = grp.expectation(grp.paulisum(ham))
    ref = orc.expectation(psi, ham)
    assert abs(got.real - ref) < 1e-11
    assert abs(got.imag) < 1e-11


@pytest.mark.parametrize("n,g,nterms", [(5, 1, 30), (10, 2, 150), (13, 3, 200)])
def test_sharded_apply_paulisum_matches_oracle(groups, n, g, nterms):
    from openvqe_b200.engine import BUF_SIGMA
    rng = np.random.default_rng(4000 + n)
    grp = groups(n, g)
    psi = random_state(rng, n)
    grp.set_state(psi)
    ham = random_hermitian(rng, n, nterms, max_weight=min(n, 8), const=-1.25)
    grp.apply_paulisum(grp.paulisum(ham))
    got = grp.get_state(BUF_SIGMA)
    ref = orc.apply_pauli_sum(psi, ham)
    assert np.max(np.abs(got - ref)) < 1e-11


@pytest.mark.parametrize("n,g,npool", [(6, 1, 20), (10, 2, 60), (13, 3, 40)])
def test_sharded_pool_overlaps_match_oracle(groups, n, g, npool):
    """Pool operators whose strings carry DIFFERENT global X patterns are split per pattern and re-added."""
    from openvqe_b200.lowering import pack_pool
    rng = np.random.default_rng(5000 + n)
    grp = groups(n, g)
    psi = random_state(rng, n)
    grp.set_state(psi)
    ham = random_hermitian(rng, n, 50, max_weight=min(n, 6))
    grp.apply_paulisum(grp.paulisum(ham))
    pool = []
    for k in range(npool):
        if k % 9 == 4:
            pool.append(Ham(n, [T(0.0, "X", [0])]))
        else:
            pool.append(random_antihermitian(rng, n, int(rng.integers(1, 9)), max_weight=min(n, 4)))
    got = grp.pool_overlaps(pack_pool(pool))
    sig = orc.apply_pauli_sum(psi, ham)
    ref = np.array([np.vdot(sig, orc.apply_pauli_sum(psi, op)) for op in pool])
    assert np.max(np.abs(got - ref)) < 1e-11
    assert got[4] == 0.0


def test_sharded_basis_state_lands_on_the_owner_rank(groups):
    grp = groups(9, 3)
    idx = 0b101_110011
    grp.set_basis_state(idx)
    v = grp.get_state()
    assert v[idx] == 1.0 and np.count_nonzero(v) == 1


def test_two_process_ipc(gpu_required):
    """One process per rank: CUDA IPC handle exchange (gloo), peer passes ordered by the device-side flag barrier.
    Uses GPUs 0 and 1 when two are present, otherwise both ranks share GPU 0 (time-sliced)."""
    from openvqe_b200 import _lib
    ndev = _lib.load().vqe_device_count()
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "sharded_worker.py"), "--devices", str(min(ndev, 2))]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "sharded worker ok" in res.stdout


@pytest.mark.parametrize("n,g", [(12, 1), (14, 2)])
def test_c5_synthetic_energy_vs_oracle(groups, n, g):
    """BASELINE config 5 at a size the oracle finishes in seconds: 256 synthetic generators + 64-group Hamiltonian."""
    from openvqe_b200.lowering import PackedTerms
    from tools import c5_synthetic as c5
    gen, ham = c5.generators(n), c5.hamiltonian(n)
    ang = gen["theta"][gen["owner"]] * gen["coeff"]
    hf = c5.hf_index(n)
    grp = groups(n, g)
    grp.set_basis_state(hf)
    grp.apply_rotations(gen["x"], gen["z"], gen["ny"], ang)
    ps = grp.paulisum(PackedTerms(n, ham["x"], ham["z"], ham["ny"], ham["cre"], np.zeros_like(ham["cre"])))
    e = grp.expectation(ps)
    ref = orc.basis_state(n, hf)
    for x, z, ny, a in zip(gen["x"], gen["z"], gen["ny"], ang):
        ref = orc.pauli_rotation(ref, int(x), int(z), int(ny), float(a))
    e_ref = sum(c * np.vdot(ref, orc.apply_pauli(ref, int(x), int(z), int(ny))) for x, z, ny, c in
                zip(ham["x"], ham["z"], ham["ny"], ham["cre"]))
    assert np.max(np.abs(grp.get_state() - ref)) < TOL
    assert abs(e.real - e_ref.real) < 1e-10 and abs(e.imag) < 1e-10   # north_star tolerance: 1e-10 Ha
    assert abs(grp.norm2() - 1.0) < 1e-12


@pytest.mark.parametrize("n,g,overlap", [(14, 2, "1"), (16, 3, "1"), (16, 3, "0"), (15, 1, "1")])
def test_gather_form_chunks_prefetched_under_the_previous_chunk(gpu_required, monkeypatch, n, g, overlap):
    """Gather-form peer passes cut into many chunks (staging buffer forced down to one tile): the partner amplitudes of
    chunk k + 1 are fetched by the gather CTAs riding in the pass kernel of chunk k (VQE_GATHER_OVERLAP=1, two staging
    buffers) or by a separate launch per chunk (=0); both must equal the oracle."""
    from openvqe_b200.sharded import ShardGroup
    from tools import c5_synthetic as c5
    monkeypatch.setenv("VQE_GATHER_STAGE_MB", "0")
    monkeypatch.setenv("VQE_GATHER_OVERLAP", overlap)
    monkeypatch.setenv("VQE_RELABEL", "0")   # this test is about peer passes
    gen = c5.generators(n, k_gen=96)
    ang = gen["theta"][gen["owner"]] * gen["coeff"]
    hf = c5.hf_index(n)
    grp = ShardGroup(n, g)
    grp.set_basis_state(hf)
    grp.apply_rotations(gen["x"], gen["z"], gen["ny"], ang)
    ref = orc.basis_state(n, hf)
    for x, z, ny, a in zip(gen["x"], gen["z"], gen["ny"], ang):
        ref = orc.pauli_rotation(ref, int(x), int(z), int(ny), float(a))
    assert np.max(np.abs(grp.get_state() - ref)) < TOL
    assert sum(e.gather_bytes() for e in grp.ranks) > 0     # the program did take gather-form passes


@pytest.mark.parametrize("n,g", [(8, 1), (11, 2), (13, 3)])
def test_sharded_exact_exponential_of_noncommuting_generators(groups, n, g):
    """Fermionic prepare_adapt_state (reference fermionic_adapt_vqe.py:12-38, expm_multiply of each whole generator) on a
    sharded state: generators with strings on global qubits that do not commute (host-driven Taylor series, three
    vectors per rank, peer passes between sigma and work) and commuting ones (rotations), against scipy."""
    from openvqe_b200.lowering import pack_operator
    rng = np.random.default_rng(1000 + n)
    grp = groups(n, g)
    psi = random_state(rng, n)
    gens = [random_antihermitian(rng, n, 5, max_weight=4) for _ in range(3)]
    gens.append(Ham(n, [T(0.7j, "XY", [0, n - 1]), T(-0.7j, "YX", [0, n - 1])]))   # commuting pair on a global qubit
    thetas = [0.3, -0.8, 1.7, 0.4]
    grp.set_state(psi)
    for op, th in zip(gens, thetas):
        grp.apply_exp(pack_operator(op), th)
    ref = orc.fermionic_adapt_state(psi, gens, thetas)
    assert np.max(np.abs(grp.get_state() - ref)) < 1e-11
    assert abs(grp.norm2() - 1.0) < 1e-11


def test_relabelling_swaps_instead_of_peer_passes(gpu_required, monkeypatch):
    """The C5 program on 8 virtual ranks with relabelling: far fewer qubit swaps than rotations that flip a global qubit, no
    gather-form traffic, the state (read back in the caller's labelling: swaps undone) and the energy (evaluated on the
    relabelled state with the relabelled twin of H) equal the oracle; then the state is re-used in the caller's labelling."""
    from openvqe_b200.lowering import PackedTerms
    from openvqe_b200.sharded import ShardGroup
    from tools import c5_synthetic as c5
    monkeypatch.setenv("VQE_RELABEL", "1")
    monkeypatch.setenv("VQE_RELABEL_FLOOR", "5")
    n, g = 16, 3
    gen, ham = c5.generators(n, k_gen=128), c5.hamiltonian(n)
    ang = gen["theta"][gen["owner"]] * gen["coeff"]
    hf = c5.hf_index(n)
    grp = ShardGroup(n, g)
    grp.set_basis_state(hf)
    grp.apply_rotations(gen["x"], gen["z"], gen["ny"], ang)
    swaps = grp.ranks[0].relabel_stats()[0]
    touching = len({int(o) for o, x in zip(gen["owner"], gen["x"]) if int(x) >> (n - g)})
    assert 0 < swaps < touching and sum(e.gather_bytes() for e in grp.ranks) == 0
    ps = grp.paulisum(PackedTerms(n, ham["x"], ham["z"], ham["ny"], ham["cre"], np.zeros_like(ham["cre"])))
    e = grp.expectation(ps)                     # on the relabelled state
    ref = orc.basis_state(n, hf)
    for x, z, ny, a in zip(gen["x"], gen["z"], gen["ny"], ang):
        ref = orc.pauli_rotation(ref, int(x), int(z), int(ny), float(a))
    e_ref = sum(c * np.vdot(ref, orc.apply_pauli(ref, int(x), int(z), int(ny))) for x, z, ny, c in
                zip(ham["x"], ham["z"], ham["ny"], ham["cre"]))
    assert abs(e.real - e_ref.real) < 1e-10 and abs(e.imag) < 1e-10
    swaps_e = grp.ranks[0].relabel_stats()[0]   # the split evaluation moved the qubits H still flipped out of the global slots
    assert swaps < swaps_e <= swaps + g
    assert np.max(np.abs(grp.get_state() - ref)) < TOL      # undoes the swaps
    assert grp.ranks[0].relabel_stats()[0] == 2 * swaps_e
    e2 = grp.expectation(ps)                    # now in the caller's labelling
    assert abs(e2 - e) < 1e-12
    # a second program on top (relabels again), then sigma = H psi (needs the caller's labelling)
    grp.apply_rotations(gen["x"][:200], gen["z"][:200], gen["ny"][:200], -ang[:200])
    for x, z, ny, a in zip(gen["x"][:200], gen["z"][:200], gen["ny"][:200], -ang[:200]):
        ref = orc.pauli_rotation(ref, int(x), int(z), int(ny), float(a))
    grp.apply_paulisum(ps, dst=1, src=0)
    sig_ref = sum(c * orc.apply_pauli(ref, int(x), int(z), int(ny)) for x, z, ny, c in zip(ham["x"], ham["z"], ham["ny"], ham["cre"]))
    assert np.max(np.abs(grp.get_state(1) - sig_ref)) < 1e-11


def test_c5_synthetic_sharded_equals_unsharded_at_22_qubits(gpu_required):
    """Size-independent property at a size the oracle cannot do: 8 virtual ranks vs one context, same energy."""
    from openvqe_b200.engine import Engine
    from openvqe_b200.lowering import PackedTerms
    from openvqe_b200.sharded import ShardGroup
    from tools import c5_synthetic as c5
    n = 22
    gen, ham = c5.generators(n, k_gen=64), c5.hamiltonian(n)
    ang = gen["theta"][gen["owner"]] * gen["coeff"]
    hp = PackedTerms(n, ham["x"], ham["z"], ham["ny"], ham["cre"], np.zeros_like(ham["cre"]))
    out = []
    for make in (lambda: Engine(n), lambda: ShardGroup(n, 3)):
        e = make()
        e.set_basis_state(c5.hf_index(n))
        e.apply_rotations(gen["x"], gen["z"], gen["ny"], ang)
        out.append((e.expectation(e.paulisum(hp)), e.norm2()))
        del e
    assert abs(out[0][0] - out[1][0]) < 1e-10
    assert abs(out[0][1] - 1.0) < 1e-12 and abs(out[1][1] - 1.0) < 1e-12


def test_replica_mode_two_ranks(gpu_required):
    """n <= 33: SPMD replicas under torchrun -- the BFGS finite-difference evaluations and the ADAPT pool sweep are
    split over the ranks and must reproduce the single-process run bit for bit (energies list, optimum, gradients)."""
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(ROOT, "tests", "replica_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "replica worker ok" in res.stdout
