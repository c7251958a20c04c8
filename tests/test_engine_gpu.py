"""GPU parity: every C-ABI entry point against the CPU oracle on seeded inputs."""
import numpy as np
import pytest

from oracle import statevector_oracle as orc
from tests.helpers import Ham, T, random_antihermitian, random_hermitian, random_pauli, random_state

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def engines(gpu_required):
    from openvqe_b200.engine import Engine
    cache = {}

    def get(n):
        if n not in cache:
            cache[n] = Engine(n)
        return cache[n]
    return get


@pytest.mark.parametrize("n", [1, 2, 3, 5, 8, 11, 12, 13, 15, 18])
def test_pauli_rotations_match_oracle(engines, n):
    from openvqe_b200.lowering import term_masks
    rng = np.random.default_rng(100 + n)
    eng = engines(n)
    psi = random_state(rng, n)
    eng.set_state(psi)
    xs, zs, nys, angs = [], [], [], []
    ref = psi.copy()
    for k in range(40):
        op, qb = random_pauli(rng, n, max_weight=min(n, 6))
        if k % 7 == 3:
            op = "Z" * len(qb)  # diagonal rotations
        x, z, ny = term_masks(op, qb, n)
        a = float(rng.uniform(-1, 1))
        xs.append(x); zs.append(z); nys.append(ny); angs.append(a)
        ref = orc.pauli_rotation(ref, x, z, ny, a)
    eng.apply_rotations(xs, zs, nys, angs)
    got = eng.get_state()
    assert np.max(np.abs(got - ref)) < TOL
    assert abs(eng.norm2() - 1.0) < 1e-12


def test_same_xmask_runs_and_zero_angles(engines):
    """8 strings of a JW double excitation share one X-mask: applied as one register-resident run."""
    from openvqe_b200.lowering import term_masks
    n = 10
    rng = np.random.default_rng(7)
    eng = engines(n)
    psi = random_state(rng, n)
    eng.set_state(psi)
    ref = psi.copy()
    xs, zs, nys, angs = [], [], [], []
    for op in ["XXXY", "XXYX", "XYXX", "YXXX", "YYYX", "YYXY", "YXYY", "XYYY"]:
        full = op[0] + "ZZ" + op[1] + op[2] + "Z" + op[3]
        qb = [1, 2, 3, 4, 6, 7, 8]
        x, z, ny = term_masks(full, qb, n)
        a = float(rng.uniform(-0.3, 0.3))
        xs.append(x); zs.append(z); nys.append(ny); angs.append(a)
        ref = orc.pauli_rotation(ref, x, z, ny, a)
    xs.append(xs[0]); zs.append(zs[0]); nys.append(nys[0]); angs.append(0.0)  # exact identity
    eng.apply_rotations(xs, zs, nys, angs)
    assert np.max(np.abs(eng.get_state() - ref)) < TOL


@pytest.mark.parametrize("n,real_start", [(6, False), (10, True), (12, False), (13, True), (14, False)])
def test_collapsed_runs_jw_excitations(engines, n, real_start):
    """Same-X-mask runs are collapsed into ONE plane rotation with a tabulated angle (the 8 strings of a JW double
    excitation touch 1/8 of the pairs).  Checked on a complex random state (general kernel variant) and on a basis
    state (purely-real variant), for +-1 phases (ny odd) and +-i phases (ny even), against string-by-string oracle."""
    from openvqe_b200.lowering import pack_operator, term_masks
    from tests.helpers import jw_excitation
    rng = np.random.default_rng(900 + n)
    eng = engines(n)
    xs, zs, nys, angs = [], [], [], []
    for g in range(12):
        if g % 4 == 3:
            p, q = sorted(rng.choice(n, size=2, replace=False).tolist())
            pk = pack_operator(jw_excitation(n, [q], [p]))
        else:
            p, q, r, s = sorted(rng.choice(n, size=4, replace=False).tolist())
            pk = pack_operator(jw_excitation(n, [r, s], [p, q]))
        th = float(rng.uniform(-0.8, 0.8))
        for k in range(len(pk)):
            xs.append(int(pk.x[k])); zs.append(int(pk.z[k])); nys.append(int(pk.ny[k])); angs.append(th * float(pk.cre[k]))
        if g % 3 == 1 and g % 4 != 3 and not real_start:
            # a run of even-ny strings on the same X-mask (+-i phases): collapsed with the imaginary-phase formula
            for op in ["XXXX", "XXYY", "YYXX", "XYXY"]:
                x, z, ny = term_masks(op, [p, q, r, s], n)
                xs.append(x); zs.append(z); nys.append(ny); angs.append(float(rng.uniform(-0.4, 0.4)))
    if real_start:
        hf = ((1 << (n // 2)) - 1) << (n - n // 2)
        eng.set_basis_state(hf)
        ref = orc.basis_state(n, hf)
    else:
        ref = random_state(rng, n)
        eng.set_state(ref)
    for x, z, ny, a in zip(xs, zs, nys, angs):
        ref = orc.pauli_rotation(ref, x, z, ny, a)
    eng.apply_rotations(xs, zs, nys, angs)
    got = eng.get_state()
    assert np.max(np.abs(got - ref)) < TOL
    if real_start:
        assert np.all(got.imag == 0.0)
        # occupation patterns a fermionic excitation cannot reach stay EXACTLY zero
        assert np.all(got[np.abs(ref) < 1e-13] == 0.0)


def test_plan_cache_reuse_and_invalidation(engines):
    """The pass plan of a rotation program is cached and re-used with new angles; it must be rebuilt when a collapsed
    pattern stops (or starts) cancelling, when a rotation is dropped (angle 0) and when the masks change."""
    from openvqe_b200.lowering import pack_operator, term_masks
    from tests.helpers import jw_excitation
    n = 12
    rng = np.random.default_rng(77)
    eng = engines(n)
    xs, zs, nys, base = [], [], [], []
    for cre, ann in (([7], [2]), ([8, 11], [0, 3]), ([6, 9], [1, 4])):
        pk = pack_operator(jw_excitation(n, cre, ann))
        xs += [int(v) for v in pk.x]; zs += [int(v) for v in pk.z]; nys += [int(v) for v in pk.ny]
        base += [float(c) for c in pk.cre]
    base = np.array(base)

    def check(angles, masks=None):
        mx, mz, mny = masks or (xs, zs, nys)
        psi = random_state(rng, n)
        eng.set_state(psi)
        eng.apply_rotations(mx, mz, mny, angles)
        ref = psi.copy()
        for x, z, ny, a in zip(mx, mz, mny, angles):
            if a != 0.0:
                ref = orc.pauli_rotation(ref, x, z, ny, a)
        assert np.max(np.abs(eng.get_state() - ref)) < TOL

    check(0.2 * base)                      # builds the plan (every generator collapses to one pattern)
    check(-0.55 * base)                    # same structure: cached plan, new angles
    broken = 0.3 * base
    broken[1] *= 0.25                      # the two strings of the single no longer cancel on |00>,|11>
    check(broken)
    check(0.4 * base)                      # ... and cancel again
    dropped = 0.1 * base
    dropped[3] = 0.0                       # one string of a double dropped
    check(dropped)
    x2, z2, ny2 = term_masks("XZY", [0, 5, 10], n)
    check(np.append(0.2 * base, 0.7), (xs + [x2], zs + [z2], nys + [ny2]))   # other masks
    check(0.2 * base)


def test_structural_zeros_stay_exact(engines):
    """Amplitudes outside the reachable sector must remain exactly 0.0."""
    from openvqe_b200.lowering import term_masks
    n = 8
    eng = engines(n)
    eng.set_basis_state(0b11000000)
    x, z, ny = term_masks("XZY", [1, 2, 3], n)
    eng.apply_rotations([x], [z], [ny], [0.37])
    got = eng.get_state()
    nz = np.nonzero(got)[0].tolist()
    assert nz == sorted([0b11000000, 0b11000000 ^ x])


@pytest.mark.parametrize("n", [2, 4, 9, 12, 14])
def test_gates_match_oracle(engines, n):
    from openvqe_b200.engine import GATE_KINDS
    rng = np.random.default_rng(200 + n)
    eng = engines(n)
    psi = random_state(rng, n)
    eng.set_state(psi)
    gates = []
    for _ in range(60):
        name = str(rng.choice(["X", "H", "RX", "RY", "RZ", "CNOT"]))
        if name == "CNOT" and n >= 2:
            c, t = rng.choice(n, size=2, replace=False).tolist()
            gates.append(("CNOT", [c, t], None))
        elif name != "CNOT":
            gates.append((name, [int(rng.integers(n))], float(rng.uniform(-3, 3))))
    ref = orc.apply_gates(psi, n, gates)
    eng.apply_gates([GATE_KINDS[g[0]] for g in gates], [g[1][0] for g in gates],
                    [g[1][1] if len(g[1]) > 1 else 0 for g in gates], [g[2] or 0.0 for g in gates])
    assert np.max(np.abs(eng.get_state() - ref)) < TOL


@pytest.mark.parametrize("n,nterms", [(1, 3), (3, 10), (6, 60), (10, 200), (12, 400), (14, 300), (16, 100)])
def test_expectation_matches_oracle(engines, n, nterms):
    rng = np.random.default_rng(300 + n)
    eng = engines(n)
    psi = random_state(rng, n)
    eng.set_state(psi)
    ham = random_hermitian(rng, n, nterms, max_weight=min(n, 8), const=0.37)
    ps = eng.paulisum(ham)
    got = eng.expectation(ps)
    ref = orc.expectation(psi, ham)
    assert abs(got.real - ref) < 1e-11
    assert abs(got.imag) < 1e-11


@pytest.mark.parametrize("n", [5, 12, 14, 17])
def test_diagonal_quadratic_form_on_the_real_layout(engines, n):
    """The X-mask-0 part of a molecular Hamiltonian (Z_p, Z_p Z_q) is evaluated as a quadratic form directly on the real
    layout of a UCC state (k_expect_diag2_rl): no expansion to interleaved complex, the state stays in the real layout.
    A diagonal part with a three-Z string is not a quadratic form and takes the general pass (expansion)."""
    from openvqe_b200.lowering import pack_operator
    from tests.helpers import jw_excitation
    rng = np.random.default_rng(5100 + n)
    eng = engines(n)
    # a purely real state: |HF> followed by JW excitations (odd-ny rotations)
    xs, zs, nys, angs = [], [], [], []
    for g in range(6):
        if n >= 4 and g % 3:
            p, q, r, s = sorted(rng.choice(n, size=4, replace=False).tolist())
            pk = pack_operator(jw_excitation(n, [r, s], [p, q]))
        else:
            p, q = sorted(rng.choice(n, size=2, replace=False).tolist())
            pk = pack_operator(jw_excitation(n, [q], [p]))
        th = float(rng.uniform(-0.7, 0.7))
        for k in range(len(pk)):
            xs.append(int(pk.x[k])); zs.append(int(pk.z[k])); nys.append(int(pk.ny[k])); angs.append(th * float(pk.cre[k]))
    hf = ((1 << (n // 2)) - 1) << (n - n // 2)
    ref = orc.basis_state(n, hf)
    for x, z, ny, a in zip(xs, zs, nys, angs):
        ref = orc.pauli_rotation(ref, x, z, ny, a)
    terms = [T(float(rng.normal()), "Z", [q]) for q in range(n)]
    terms += [T(float(rng.normal()), "ZZ", [p, q]) for p in range(n) for q in range(p + 1, n) if rng.random() < 0.7]
    terms += [T(float(rng.normal()), "XX", sorted(rng.choice(n, size=2, replace=False).tolist())) for _ in range(3)]
    ham = Ham(n, terms, 0.41)
    for extra, stays_real in (([], True), ([T(0.3, "ZZZ", [0, 1, 2])], False)):
        if extra and n < 3:
            continue
        h2 = Ham(n, terms + extra, 0.41)
        ps = eng.paulisum(h2)
        eng.set_basis_state(hf)
        eng.apply_rotations(xs, zs, nys, angs)
        was_real = eng.real_layout
        got = eng.expectation(ps)
        assert abs(got.real - orc.expectation(ref, h2)) < 1e-11 and abs(got.imag) < 1e-11
        if was_real and n >= 13:
            assert eng.real_layout == stays_real
        assert np.max(np.abs(eng.get_state() - ref)) < TOL


def test_expectation_complex_coefficients(engines):
    n = 7
    rng = np.random.default_rng(5)
    eng = engines(n)
    psi = random_state(rng, n)
    eng.set_state(psi)
    op = random_antihermitian(rng, n, 30)
    got = eng.expectation(eng.paulisum(op))
    ref = np.vdot(psi, orc.apply_pauli_sum(psi, op))
    assert abs(got - ref) < 1e-11


@pytest.mark.parametrize("n,nterms", [(2, 4), (5, 30), (10, 150), (12, 300), (13, 200), (15, 80)])
def test_apply_paulisum_matches_oracle(engines, n, nterms):
    from openvqe_b200.engine import BUF_SIGMA
    rng = np.random.default_rng(400 + n)
    eng = engines(n)
    psi = random_state(rng, n)
    eng.set_state(psi)
    ham = random_hermitian(rng, n, nterms, max_weight=min(n, 8), const=-1.25)
    eng.apply_paulisum(eng.paulisum(ham))
    got = eng.get_state(BUF_SIGMA)
    ref = orc.apply_pauli_sum(psi, ham)
    assert np.max(np.abs(got - ref)) < 1e-11


@pytest.mark.parametrize("n,npool", [(3, 5), (8, 60), (12, 120), (13, 40), (15, 30)])
def test_pool_overlaps_match_oracle(engines, n, npool):
    from openvqe_b200.engine import BUF_SIGMA
    from openvqe_b200.lowering import pack_pool
    rng = np.random.default_rng(500 + n)
    eng = engines(n)
    psi = random_state(rng, n)
    eng.set_state(psi)
    ham = random_hermitian(rng, n, 50, max_weight=min(n, 6))
    eng.apply_paulisum(eng.paulisum(ham))
    pool = []
    for k in range(npool):
        if k % 9 == 4:
            pool.append(Ham(n, [T(0.0, "X", [0])]))  # identically-zero operator stays in the pool
        else:
            pool.append(random_antihermitian(rng, n, int(rng.integers(1, 9)), max_weight=min(n, 4)))
    got = eng.pool_overlaps(pack_pool(pool))
    sig = orc.apply_pauli_sum(psi, ham)
    ref = np.array([np.vdot(sig, orc.apply_pauli_sum(psi, op)) for op in pool])
    assert np.max(np.abs(got - ref)) < 1e-11
    assert got[4] == 0.0


@pytest.mark.parametrize("n", [4, 8, 12])
def test_exact_exponential_matches_expm_multiply(engines, n):
    from openvqe_b200.lowering import pack_operator
    rng = np.random.default_rng(600 + n)
    eng = engines(n)
    psi = random_state(rng, n)
    # non-commuting anti-Hermitian generator -> Taylor path
    gen = random_antihermitian(rng, n, 6, max_weight=min(n, 4))
    eng.set_state(psi)
    eng.apply_exp(pack_operator(gen), 0.8)
    ref = orc.fermionic_adapt_state(psi, [gen], [0.8])
    assert np.max(np.abs(eng.get_state() - ref)) < 1e-11
    # commuting strings (same string twice + a disjoint one) -> rotation path
    gen2 = Ham(n, [T(0.3j, "XY", [0, 1]), T(-0.7j, "Z", [n - 1]), T(0.2j, "XY", [0, 1])])
    eng.set_state(psi)
    eng.apply_exp(pack_operator(gen2), -1.1)
    ref2 = orc.fermionic_adapt_state(psi, [gen2], [-1.1])
    assert np.max(np.abs(eng.get_state() - ref2)) < 1e-11


def test_overlap_and_inner(engines):
    n = 9
    rng = np.random.default_rng(11)
    eng = engines(n)
    a, b = random_state(rng, n), random_state(rng, n)
    eng.set_state(a)
    assert abs(eng.overlap_host(b) - np.vdot(b, a)) < 1e-12


def test_error_paths(engines):
    from openvqe_b200._lib import VQEError
    eng = engines(3)
    with pytest.raises(VQEError):
        eng.set_basis_state(8)
    with pytest.raises(VQEError):
        eng.apply_rotations([16], [0], [0], [0.1])  # mask outside the register
    with pytest.raises(VQEError):
        eng.apply_rotations([1], [1], [0], [0.1])  # ny inconsistent with masks


@pytest.mark.parametrize("n", [13, 16, 18])
def test_wide_pauli_strings_get_a_smaller_low_bit_floor(engines, n):
    """Strings with 8-12 X/Y letters on high qubits (eight-letter qubit pools, generic circuits) do not fit a tile
    with 5 fixed low bits: the planner gives such a pass shorter contiguous segments instead of refusing."""
    from openvqe_b200.engine import BUF_SIGMA
    from openvqe_b200.lowering import pack_pool, term_masks
    rng = np.random.default_rng(3100 + n)
    eng = engines(n)
    psi = random_state(rng, n)
    eng.set_state(psi)
    ref = psi.copy()
    xs, zs, nys, angs, terms = [], [], [], [], []
    for k in range(14):
        w = int(rng.integers(8, min(n - 2, 12) + 1))
        qb = sorted(rng.choice(n - 2, size=w, replace=False).tolist())       # qubits 0 .. n-3: the high index bits
        op = "".join(rng.choice(list("XY"), size=w))
        x, z, ny = term_masks(op, qb, n)
        a = float(rng.uniform(-0.6, 0.6))
        xs.append(x); zs.append(z); nys.append(ny); angs.append(a)
        terms.append(T(float(rng.normal()), op, qb))
        ref = orc.pauli_rotation(ref, x, z, ny, a)
    eng.apply_rotations(xs, zs, nys, angs)
    assert np.max(np.abs(eng.get_state() - ref)) < TOL
    ham = Ham(n, terms, 0.1)
    got = eng.expectation(eng.paulisum(ham))
    assert abs(got.real - orc.expectation(ref, ham)) < 1e-11 and abs(got.imag) < 1e-11
    eng.apply_paulisum(eng.paulisum(ham))
    assert np.max(np.abs(eng.get_state(BUF_SIGMA) - orc.apply_pauli_sum(ref, ham))) < 1e-11
    pool = [Ham(n, [T(1j * t.coeff, t.op, t.qbits)]) for t in terms[:6]]
    ov = eng.pool_overlaps(pack_pool(pool))
    sig = orc.apply_pauli_sum(ref, ham)
    want = np.array([np.vdot(sig, orc.apply_pauli_sum(ref, op)) for op in pool])
    assert np.max(np.abs(ov - want)) < 1e-11
