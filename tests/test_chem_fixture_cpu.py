"""Fixture tooling for the molecules BASELINE.json names (oracle/chem/gto.py: s and p Gaussians, RHF, frozen core) against
textbook numbers, and the oracle against the reference outputs stored in the LiH fixture (config C2)."""
import numpy as np
import pytest

from oracle import statevector_oracle as orc
from oracle.chem import gto, hchain
from tests.helpers import ham_from_json, load_golden


def test_sp_integrals_reproduce_textbook_scf_energies():
    b = hchain.BOHR
    # H2O / STO-3G, R = 1.1 A, 104 deg (the geometry of the "Crawford programming projects"): E_nuc and E_SCF to 1e-9
    geo = [("O", (0.0, -0.143225816552 * b, 0.0)), ("H", (1.638036840407 * b, 1.136548822547 * b, 0.0)),
           ("H", (-1.638036840407 * b, 1.136548822547 * b, 0.0))]
    s, h, eri, e_nuc = gto.integrals(geo, "sto-3g")
    assert abs(e_nuc - 8.002367061810450) < 1e-9
    assert abs(s[0, 1] - 0.2367039) < 1e-6 and abs(np.diag(s) - 1.0).max() < 1e-12
    e_el, eps, c = hchain.rhf(s, h, eri, 10)
    assert abs(e_el + e_nuc - (-74.942079928192)) < 1e-9
    # LiH / STO-3G at 1.6 A: literature RHF energy -7.8618 (4 decimals)
    mi = gto.molecular_integrals([("Li", (0, 0, 0)), ("H", (0, 0, 1.6))], "sto-3g")
    assert abs(mi["hf_energy"] - (-7.8618)) < 1e-4 and mi["n_elec"] == 4 and mi["one_body"].shape == (6, 6)


def test_s_only_integrals_equal_the_round1_tooling():
    xyz = hchain.chain(4, 0.85)
    s2, h2, eri2, en2 = hchain.integrals(xyz, "sto-3g")
    s3, h3, eri3, en3 = gto.integrals([("H", x) for x in xyz], "sto-3g")
    # gto.py renormalises each contraction to unit self-overlap (the published 8-digit coefficients give 1 + 2e-8)
    assert abs(en2 - en3) < 1e-12
    for a, b in ((s2, s3), (h2, h3), (eri2, eri3)):
        assert np.abs(a - b).max() < 1e-7


def test_frozen_core_energy_bookkeeping():
    """Freezing the lowest MO folds its mean field into h and its energy into the constant: <HF|H_active|HF> is unchanged."""
    geo = [("Li", (0, 0, 0)), ("H", (0, 0, 1.45))]
    full = gto.molecular_integrals(geo, "sto-3g")
    act = gto.molecular_integrals(geo, "sto-3g", n_frozen=1)
    assert act["n_elec"] == 2 and act["one_body"].shape == (5, 5)
    h1, g = act["one_body"], act["two_body"]      # g[p,q,r,s] = (p s | q r)
    e = act["nuclear_repulsion"] + 2.0 * h1[0, 0] + g[0, 0, 0, 0]
    assert abs(e - full["hf_energy"]) < 1e-10


def test_lih_fixture_oracle_equals_reference_outputs():
    """Config C2 inputs: the numpy oracle against the outputs of the unmodified reference modules stored in lih_sto3g.json.gz."""
    import os
    from openvqe_b200.common_files.pools import generate_yxxx_pool
    if not os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lih_sto3g.json.gz")):
        pytest.skip("fixture not generated yet: python oracle/make_golden_r3.py --lih")
    fx = load_golden("lih_sto3g.json.gz")
    ham = ham_from_json(fx["hamiltonian"])
    assert ham.nbqbits == 12 and fx["yxxx_pool_size"] == 285
    hf = orc.basis_state(12, fx["hf_init_sp"])
    assert abs(orc.expectation(hf, ham) - fx["hf_energy"]) < 1e-10
    assert fx["fci"] < fx["hf_energy"] - 0.01
    _, pool = generate_yxxx_pool(12)
    g0 = orc.qubit_pool_gradients(hf, ham, pool)
    assert np.abs(np.asarray(g0) - np.asarray(fx["qubit_gradients_at_hf"])).max() < 1e-10
    a = fx["qubit_gradients_at_ansatz"]
    st = orc.qubit_adapt_state(hf, [pool[i] for i in a["indices"]], a["parameters"])
    assert np.abs(st - (np.array(a["state_re"]) + 1j * np.array(a["state_im"]))).max() < 1e-12
    g1 = orc.qubit_pool_gradients(st, ham, pool)
    assert np.abs(np.asarray(g1) - np.asarray(a["gradients"])).max() < 1e-10
