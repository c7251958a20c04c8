"""Host logic of the section-8f-item-4 widening on the oracle-backed engine double (no GPU here; the same code runs on the
CUDA engine in tests/test_widen_gpu.py): the QPU plugin's circuit parsing / global-phase bookkeeping against the qat
shim's numpy simulator, and the two-pass Lanczos ground state against a dense eigh."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "qat_shim")


@pytest.fixture()
def double_engine():
    from openvqe_b200 import engine as engine_mod
    from tests.oracle_engine import OracleEngine
    saved = engine_mod._ENGINE_FACTORY
    engine_mod.release_engines()
    engine_mod._ENGINE_FACTORY = lambda n, device: OracleEngine(n, device)
    added = SHIM not in sys.path
    if added:
        sys.path.insert(0, SHIM)
    yield
    engine_mod._ENGINE_FACTORY = saved
    engine_mod.release_engines()
    if added:
        sys.path.remove(SHIM)
    for name in [m for m in sys.modules if m == "qat" or m.startswith("qat.")]:
        del sys.modules[name]


def _random_program(rng, n, n_gates):
    from qat.lang.AQASM import CNOT, H, RX, RY, RZ, X, Y, Z, Program
    prog = Program()
    reg = prog.qalloc(n)
    for _ in range(n_gates):
        k = rng.integers(0, 9)
        q = int(rng.integers(0, n))
        if k == 8:
            t = int((q + 1 + rng.integers(0, n - 1)) % n)
            prog.apply(CNOT, reg[q], reg[t])
        elif k < 5:
            prog.apply([X, Y, Z, H, H][k], reg[q])
        else:
            prog.apply([RX, RY, RZ][k - 5](float(rng.uniform(-3, 3))), reg[q])
    return prog.to_circ()


def test_qpu_matches_the_shim_simulator(double_engine):
    from qat.qpus import get_default_qpu
    from openvqe_b200.qpu import B200QPU
    from tests.helpers import random_hermitian
    rng = np.random.default_rng(5)
    for n in (3, 5):
        circ = _random_program(rng, n, 60)
        obs = random_hermitian(rng, n, 12, const=0.3)
        ref = get_default_qpu().submit(circ.to_job(job_type="OBS", observable=obs)).value
        mine = B200QPU().submit(circ.to_job(job_type="OBS", observable=obs)).value
        assert abs(ref - mine) < 1e-12
        want = np.zeros(1 << n, dtype=np.complex128)
        for s in get_default_qpu().submit(circ.to_job()):
            want[s.state.int] = s.amplitude
        got = np.zeros(1 << n, dtype=np.complex128)
        for s in B200QPU().submit(circ.to_job()):      # the loop of reference get_statevector (fermionic_adapt_vqe.py:326-327)
            got[s.state.int] = s.amplitude
            assert abs(s.probability - abs(s.amplitude) ** 2) < 1e-15
        assert np.abs(want - got).max() < 1e-12         # global phase of the Y / Z rewriting included


def test_qpu_iterate_simple_and_phase_gates(double_engine):
    """myQLM's public flattening (``Circuit.iterate_simple``) and the S / T / PH rewriting, against explicit matrices."""
    from openvqe_b200.qpu import B200QPU

    class Circ:
        nbqbits = 2

        def iterate_simple(self):
            return iter([("H", [], [0]), ("H", [], [1]), ("S", [], [0]), ("T", [], [1]), ("PH", [0.7], [0]), ("CNOT", [], [0, 1]),
                         ("I", [], [1])])

    class Job:
        circuit = Circ()
        observable = None

    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    u0 = np.diag([1, np.exp(0.7j)]) @ np.diag([1, 1j]) @ h
    u1 = np.diag([1, np.exp(0.25j * np.pi)]) @ h
    psi = np.kron(u0[:, 0], u1[:, 0])            # qubit 0 = most significant bit
    cnot = np.eye(4)[[0, 1, 3, 2]]
    want = cnot @ psi
    got = np.zeros(4, dtype=np.complex128)
    for s in B200QPU().submit(Job()):
        got[s.state.int] = s.amplitude
    assert np.abs(want - got).max() < 1e-14
    with pytest.raises(NotImplementedError):
        class Bad(Circ):
            def iterate_simple(self):
                return iter([("SWAP", [], [0, 1])])
        Job.circuit = Bad()
        B200QPU().submit(Job())


def test_lanczos_ground_state_equals_dense_eigh(double_engine):
    from openvqe_b200.engine import get_engine
    from openvqe_b200.ground_state import lanczos_ground_state
    from tests.helpers import ham_from_json, load_golden
    for name, e_fci in (("h2_631g.json.gz", -1.1516885475166094), ("h4_sto3g.json.gz", -2.178313632880399)):
        fx = load_golden(name)
        ham = ham_from_json(fx["hamiltonian"])
        eng = get_engine(ham.nbqbits)
        gs = lanczos_ground_state(eng, ham, fx["hf_init_sp"])
        assert abs(gs.energy - e_fci) < 2e-8                         # the notebooks' FCI energies (pins G2, G4)
        ps = eng.paulisum(ham)
        dim = 1 << ham.nbqbits
        hmat = np.stack([eng._apply(ps.packed, np.eye(dim, dtype=np.complex128)[i]) for i in range(dim)], axis=1)
        w, v = np.linalg.eigh(hmat)
        assert abs(gs.energy - w[0]) < 1e-12
        y = gs.vector()
        assert np.linalg.norm(hmat @ y - gs.energy * y) < 1e-10
        assert abs(abs(np.vdot(v[:, 0], y)) ** 2 - 1.0) < 1e-12
        eng.set_state(v[:, 0])
        assert abs(gs.fidelity() - 1.0) < 1e-12
