"""The reference's own facade on the drop-in modules (Option A), executed -- not just asserted.

Runs in the BUILD container only (needs /root/reference; skipped elsewhere): the UNMODIFIED reference package
(openvqe/vqe.py, openvqe/algorithms/*.py, molecule factories, pool generators) is imported through oracle/qat_shim and
``VQE.algorithm(...).execute()`` is run twice --

  1. as shipped: the reference's own openvqe.ucc_family / openvqe.adapt modules on the shim's numpy simulator;
  2. with the four hot-path modules aliased to openvqe_b200.* before ``openvqe.vqe`` is imported (the sys.modules switch of
     INTEGRATION.md), the engine being the oracle-backed test double of tests/oracle_engine.py (there is no GPU here;
     the same modules run on the CUDA engine in tests/test_boundary_gpu.py).

Both runs must produce the same ``iterations`` / ``result`` structures: keys, gate counts, operator indices, energies.
pyscf is absent, so ``perform_pyscf_computation`` is served from the fixture integral code (oracle/chem/hchain.py)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OPENVQE_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "openvqe")), reason="needs the reference tree (build container)")

ALIASES = {"openvqe.ucc_family.get_energy_ucc": "openvqe_b200.ucc_family.get_energy_ucc",
           "openvqe.ucc_family.get_energy_qucc": "openvqe_b200.ucc_family.get_energy_qucc",
           "openvqe.adapt.fermionic_adapt_vqe": "openvqe_b200.adapt.fermionic_adapt_vqe",
           "openvqe.adapt.qubit_adapt_vqe": "openvqe_b200.adapt.qubit_adapt_vqe"}


def _fake_pyscf(geometry, basis, spin, charge, run_fci=True):
    from oracle.chem.hchain import molecular_integrals
    from qat.fermion.chemistry.ucc import transform_integrals_to_new_basis
    mi = molecular_integrals([xyz for _, xyz in geometry], basis.lower())
    n_orb, ne = mi["one_body"].shape[0], mi["n_elec"]
    if n_orb == 4:  # H2/6-31G: the MO sign class of the reference's notebooks, as oracle/make_golden.py
        mi["one_body"], mi["two_body"] = transform_integrals_to_new_basis(mi["one_body"], mi["two_body"], np.diag([1.0, 1.0, 1.0, -1.0]))
    rdm1 = np.diag([2.0] * (ne // 2) + [0.0] * (n_orb - ne // 2))
    return rdm1, mi["orbital_energies"], mi["nuclear_repulsion"], ne, mi["one_body"], mi["two_body"], \
        {"HF": mi["hf_energy"], "MP2": None, "FCI": -1.1516885475166094}


def _purge():
    for name in [m for m in sys.modules if m == "openvqe" or m.startswith("openvqe.")]:
        del sys.modules[name]


def _run(algo, generator, opts, aliased):
    import importlib
    paths = [os.path.join(ROOT, "oracle", "qat_shim"), REF]
    added = [p for p in paths if p not in sys.path]
    sys.path[:0] = added
    from openvqe_b200 import engine as engine_mod
    saved_factory = engine_mod._ENGINE_FACTORY
    try:
        import qat.fermion.chemistry.pyscf_tools as pt
        pt.perform_pyscf_computation = _fake_pyscf
        _purge()
        if aliased:
            from tests.oracle_engine import OracleEngine
            engine_mod.release_engines()
            engine_mod._ENGINE_FACTORY = lambda n, device: OracleEngine(n, device)
            for ref_name, mine in ALIASES.items():
                sys.modules[ref_name] = importlib.import_module(mine)
        from openvqe.vqe import VQE
        np.random.seed(11)   # the qubit-ADAPT facade draws its 'random' pool from numpy's global generator
        a = VQE.algorithm(algo, "H2", generator, "JW", False, dict(opts))
        with contextlib.redirect_stdout(io.StringIO()):
            a.execute()
        return a
    finally:
        engine_mod._ENGINE_FACTORY = saved_factory
        engine_mod.release_engines()
        _purge()
        for p in added:
            sys.path.remove(p)


def test_ucc_facade_runs_on_the_drop_in():
    """main_ucc.py: VQE.algorithm('ucc', 'H2', 'sUPCCGSD', 'JW', False).execute() (reference algorithms/ucc.py:37-86)."""
    ref = _run("ucc", "sUPCCGSD", {}, aliased=False)
    mine = _run("ucc", "sUPCCGSD", {}, aliased=True)
    assert set(mine.iterations) == set(ref.iterations) and set(mine.result) == set(ref.result)
    assert mine.result["CNOT1"] == ref.result["CNOT1"] == 608          # notebook pin G2
    assert mine.result["CNOT2"] == ref.result["CNOT2"] and mine.result["len_op1"] == ref.result["len_op1"] == 18
    for k in ("1", "2"):
        assert abs(mine.iterations["minimum_energy_result%s_guess" % k][0] - ref.iterations["minimum_energy_result%s_guess" % k][0]) < 1e-8
        a, b = mine.result["energies_" + k], ref.result["energies_" + k]
        assert abs(a[0] - b[0]) < 1e-12 and np.abs(np.array(a[:19]) - np.array(b[:19])).max() < 1e-10   # E(0.01 * 1_18) and the first FD sweep
    assert abs(mine.result["energies_1"][0] - (-1.1167300964889262)) < 2e-8                           # notebook pin G2


def test_fermionic_adapt_facade_runs_on_the_drop_in():
    """main_fermionic_adapt.py (reference algorithms/fermionic_adapt.py:57-71) with the facade's own defaults (COBYLA,
    norm threshold 1e-2): the 175-operator pool and the sparse matrices are built by the reference factory and handed to the
    drop-in, which ignores the matrices.  Notebook G6: operators [38, 32, 29, 23, 2]."""
    ref = _run("fermionic_adapt", "spin_complement_gsd", {}, aliased=False)
    mine = _run("fermionic_adapt", "spin_complement_gsd", {}, aliased=True)
    assert set(mine.iterations) == set(ref.iterations) and set(mine.result) == set(ref.result) and "indices" in ref.result
    assert mine.result["indices"] == ref.result["indices"] == [38, 32, 29, 23, 2]
    assert mine.result["Number_operators"] == ref.result["Number_operators"]
    for key in ("CNOTs", "Hadamard", "RY", "RX"):
        assert mine.iterations[key] == ref.iterations[key]
    assert np.abs(np.array(mine.iterations["energies"]) - np.array(ref.iterations["energies"])).max() < 1e-7
    assert np.abs(np.array(mine.iterations["norms"]) - np.array(ref.iterations["norms"])).max() < 5e-6   # COBYLA (tol 1e-6) trajectories differ at rounding level
    assert abs(mine.result["final_energy_last_iteration"] - ref.result["final_energy_last_iteration"]) < 1e-7
