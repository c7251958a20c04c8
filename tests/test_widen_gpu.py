"""GPU tests of the section-8f-item-4 widening: the Lanczos ground state that replaces the dense eigh of
reference adapt/fermionic_adapt_vqe.py:474 and the engine as a generic QPU (reference call form
``qpu.submit(circ.to_job(observable=H)).value``, common_files/get_energy_WSSVQE.py:151-178)."""
import contextlib
import io
import math
import os
import sys

import numpy as np
import pytest

from oracle import statevector_oracle as orc
from tests.helpers import ham_from_json, load_golden, pool_from_json

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lanczos_matches_fci_pins(gpu_required):
    """E0 of the HF sector against the FCI energies of the fixtures (H2/6-31G and H4/STO-3G: notebook pins G2 / G4;
    H6/STO-3G: the shim-run reference's own eigh), residual |H y - E y| measured with the engine's own kernels."""
    from openvqe_b200.engine import BUF_AUX, BUF_WORK, get_engine
    from openvqe_b200.ground_state import lanczos_ground_state
    for name in ("h2_631g.json.gz", "h4_sto3g.json.gz", "h6_sto3g.json.gz"):
        fx = load_golden(name)
        ham = ham_from_json(fx["hamiltonian"])
        eng = get_engine(ham.nbqbits)
        gs = lanczos_ground_state(eng, ham, fx["hf_init_sp"])
        assert abs(gs.energy - fx["fci"]) < 1e-9, (name, gs.energy, fx["fci"])
        eng.apply_paulisum(eng.paulisum(ham), dst=BUF_WORK, src=BUF_AUX)
        eng.axpby(BUF_WORK, BUF_AUX, -gs.energy, 1.0)
        assert math.sqrt(eng.norm2(BUF_WORK)) < 1e-9
        assert abs(eng.norm2(BUF_AUX) - 1.0) < 1e-12
        if ham.nbqbits <= 8:   # against the dense eigenvector of the reference's own definition (eigh of the full matrix)
            w, v = np.linalg.eigh(orc.sparse_matrix(ham).toarray())
            assert abs(abs(np.vdot(v[:, 0], gs.vector())) ** 2 - 1.0) < 1e-10


def test_lanczos_24_qubits(gpu_required):
    """The bench Hamiltonian (H12/STO-3G, 14 905 terms): residual below 1e-8, energy below the HF and the UCCSD(MP2) energies."""
    from openvqe_b200.engine import BUF_AUX, BUF_WORK, get_engine
    from openvqe_b200.ground_state import lanczos_ground_state
    from openvqe_b200.lowering import PackedTerms
    d = np.load(os.path.join(ROOT, "tests", "golden", "h12_sto3g_24q.npz"))
    n = int(d["n"])
    ham = PackedTerms(n, d["ham_x"], d["ham_z"], d["ham_ny"], d["ham_cre"], np.zeros_like(d["ham_cre"]))
    eng = get_engine(n)
    ps = eng.paulisum(ham)
    eng.set_basis_state(int(d["hf_init_sp"]))
    e_hf = eng.expectation(ps).real
    gs = lanczos_ground_state(eng, ps, int(d["hf_init_sp"]), tol=1e-8)
    assert gs.energy < e_hf - 0.1 and gs.iterations < 300
    eng.apply_paulisum(ps, dst=BUF_WORK, src=BUF_AUX)
    eng.axpby(BUF_WORK, BUF_AUX, -gs.energy, 1.0)
    assert math.sqrt(eng.norm2(BUF_WORK)) < 5e-8
    eng.set_basis_state(int(d["hf_init_sp"]))
    assert 0.3 < gs.fidelity() < 1.0          # weight of the HF determinant in the ground state


def test_adapt_fidelity_from_the_device_ground_state(gpu_required, monkeypatch):
    """fermionic_adapt_vqe with the Lanczos path forced (as above 14 qubits) returns the fidelities of the eigh path."""
    from openvqe_b200.adapt import fermionic_adapt_vqe as fa
    fx = load_golden("h2_631g.json.gz")
    ham = ham_from_json(fx["hamiltonian"])
    pool = pool_from_json(8, fx["spin_complement_gsd"])
    ham.get_matrix = lambda sparse=False: orc.sparse_matrix(ham).toarray()
    ket = orc.basis_state(8, fx["hf_init_sp"]).reshape(-1, 1)
    runs = []
    for limit in (14, 4):
        monkeypatch.setattr(fa, "FIDELITY_MAX_QUBITS", limit)
        with contextlib.redirect_stdout(io.StringIO()):
            it, res = fa.fermionic_adapt_vqe(None, None, ket, ham, pool, fx["hf_init_sp"], 1, fx["fci"], "COBYLA", 1e-6, "norm", 1e-2, 35)
        runs.append((it, res))
    (it_a, res_a), (it_b, res_b) = runs
    assert res_a["indices"] == res_b["indices"] == [38, 32, 29, 23, 2]
    assert len(it_a["fidelity"]) == len(it_b["fidelity"]) >= 4
    assert np.abs(np.array(it_a["fidelity"]) - np.array(it_b["fidelity"])).max() < 1e-9
    assert it_b["fidelity"][-1] > 0.999


def test_qpu_runs_the_quccsd_circuit(gpu_required):
    """B200QPU against the qat shim's numpy simulator on the gate-defined QUCCSD circuit of H4/STO-3G (292 CNOTs) and on the
    reference value of E(theta_MP2) (notebook pin G4, through the shim-run reference: fixture ``action_quccsd``)."""
    shim = os.path.join(ROOT, "oracle", "qat_shim")
    sys.path.insert(0, shim)
    try:
        from qat.core import Circuit
        from qat.qpus import get_default_qpu
        from openvqe_b200.common_files.circuit import quccsd_circuit
        from openvqe_b200.qpu import B200QPU
        from tests.helpers import FermiOp
        fx = load_golden("h4_sto3g.json.gz")
        ham = ham_from_json(fx["hamiltonian"])
        ops = [FermiOp(8, e) for e in fx["excitations"]]
        for case in fx["action_quccsd"]:
            summary = quccsd_circuit(8, fx["hf_init_sp"], ops, case["theta"])
            mine = B200QPU().submit(summary.to_job(observable=ham)).value
            assert abs(mine - case["energy"]) < 1e-10
        ref_circ = Circuit(8, summary.gates)
        want = np.zeros(256, dtype=np.complex128)
        for s in get_default_qpu().submit(ref_circ.to_job()):
            want[s.state.int] = s.amplitude
        got = np.zeros(256, dtype=np.complex128)
        for s in B200QPU().submit(summary.to_job()):
            got[s.state.int] = s.amplitude
        assert np.abs(want - got).max() < 1e-12
    finally:
        sys.path.remove(shim)
        for name in [m for m in sys.modules if m == "qat" or m.startswith("qat.")]:
            del sys.modules[name]
