"""TEST INFRASTRUCTURE ONLY -- fixtures for the two molecules BASELINE.json names that need p-type Gaussians.

Run in the BUILD container (needs /root/reference; the GPU box has only the outputs):

    python oracle/make_golden_r3.py [--lih] [--h2o]

Integrals come from oracle/chem/gto.py (s and p shells, validated against textbook SCF energies); everything after the
integrals -- Jordan-Wigner image, operator pools, gradients, energies -- is produced by the UNMODIFIED reference modules
run through oracle/qat_shim, as in make_golden.py.

  lih_sto3g.json.gz   config C2: LiH / STO-3G, r = 1.45 A (reference molecule_factory_with_sparse.py:63-68), full space,
                      6 orbitals / 4 electrons / 12 qubits.  Hamiltonian, E_HF, FCI; outputs of the reference's qubit-ADAPT
                      gradient sweep (term_to_matrix_sparse + calculate_gradient, qubit_adapt_vqe.py:81-150) over the
                      285-operator YXXX pool at |HF> and at a 3-operator ADAPT state (prepare_adapt_state, :20-55); a
                      short run of the reference qubit_adapt_vqe loop itself.
  h2o_631g_24q.npz    config C4: H2O / 6-31G (geometry of reference molecule_factory.py:138-148), O 1s core frozen:
                      12 active orbitals / 8 electrons / 24 qubits.  Packed Hamiltonian, the UCCSD rotation program and the
                      QUCCSD excitation list with MP2 amplitudes (64 singles + 1 360 doubles), same layout as
                      h12_sto3g_24q.npz (the 24-qubit parity tests and bench.py --workload h2o read it).
"""
from __future__ import annotations

import gzip
import json
import os
import sys

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import OUT, fci_energy, ham_to_json, hf_integer, quiet  # noqa: E402  (also sets sys.path for shim + reference)


def hamiltonian_of(mi):
    from qat.fermion import ElectronicStructureHamiltonian
    from qat.fermion.chemistry.ucc import convert_to_h_integrals
    from qat.fermion.transforms import transform_to_jw_basis
    hpq, hpqrs = convert_to_h_integrals(mi["one_body"], mi["two_body"])
    hf = ElectronicStructureHamiltonian(hpq, hpqrs, constant_coeff=mi["nuclear_repulsion"])
    return hf, transform_to_jw_basis(hf), hpqrs


def make_lih():
    from openvqe.adapt import qubit_adapt_vqe as ref_qa
    from openvqe.common_files.qubit_pool import QubitPool
    from oracle import statevector_oracle as orc
    from oracle.chem import gto
    r = 1.45
    mi = gto.molecular_integrals([("Li", (0, 0, 0)), ("H", (0, 0, r))], "sto-3g")
    _, h_sp, _ = hamiltonian_of(mi)
    n, hf = 12, hf_integer(4, 12)
    fci = fci_energy(h_sp, 4)
    fx = {"system": "LiH STO-3G r=1.45 A (reference molecule_factory_with_sparse.py:63-68), full space, 12 qubits",
          "hamiltonian": ham_to_json(h_sp), "hf_init_sp": hf, "hf_energy": mi["hf_energy"], "fci": fci,
          "orbital_energies": list(map(float, mi["orbital_energies"])), "nuclear_repulsion": mi["nuclear_repulsion"]}
    _, yxxx = QubitPool().generate_yxxx_pool(n)
    fx["yxxx_pool_size"] = len(yxxx)
    hs = h_sp.get_matrix(sparse=True)
    ket = scipy.sparse.csr_matrix(orc.basis_state(n, hf).reshape(-1, 1))
    mats = [ref_qa.term_to_matrix_sparse(op) for op in yxxx]
    g0 = [float(ref_qa.calculate_gradient(m, ket, hs)) for m in mats]
    fx["qubit_gradients_at_hf"] = g0
    order = list(np.argsort(-np.abs(np.array(g0)), kind="stable")[:3])
    idx, par = [int(i) for i in order], [0.11, -0.07, 0.05]
    st = ref_qa.prepare_adapt_state(ket, [yxxx[i] for i in idx], par)
    st = np.asarray(st.todense()).reshape(-1)
    stc = scipy.sparse.csr_matrix(st.reshape(-1, 1))
    fx["qubit_gradients_at_ansatz"] = {"indices": idx, "parameters": par, "state_re": st.real.tolist(),
                                       "state_im": st.imag.tolist(),
                                       "gradients": [float(ref_qa.calculate_gradient(m, stc, hs)) for m in mats]}
    fx["ucc_action"] = [{"indices": idx, "theta": th, "energy": float(ref_qa.ucc_action(h_sp, [yxxx[i] for i in idx], hf, th))}
                        for th in ([0.01, 0.01, 0.01], par)]
    out = quiet(ref_qa.qubit_adapt_vqe, h_sp, hs, ket, n, yxxx, hf, fci, n_max_grads=1, adapt_conver="norm",
                adapt_thresh=1e-7, adapt_maxiter=3, tolerance_sim=1e-9, method_sim="BFGS")
    fx["qubit_adapt_run"] = {"iterations_sim": out[0], "result_sim": out[2], "adapt_maxiter": 3}
    with gzip.open(os.path.join(OUT, "lih_sto3g.json.gz"), "wt") as f:
        json.dump(fx, f)
    print("LiH: E_HF %.10f  FCI %.10f  terms %d  max|g| at HF %.6f" % (mi["hf_energy"], fci, len(h_sp.terms), max(g0)))


def make_h2o():
    from qat.fermion.chemistry.ucc_deprecated import get_cluster_ops_and_init_guess
    from qat.fermion.transforms import transform_to_jw_basis
    from openvqe_b200.lowering import pack_operator
    from oracle.chem import gto
    r, theta = 1.0285, 0.538 * np.pi
    geometry = [("O", (0, 0, 0)), ("H", (0, 0, r)), ("H", (0, r * np.sin(np.pi - theta), r * np.cos(np.pi - theta)))]
    mi = gto.molecular_integrals(geometry, "6-31g", n_frozen=1)
    n_orb, n_el = mi["one_body"].shape[0], mi["n_elec"]
    n = 2 * n_orb
    assert (n_orb, n_el) == (12, 8)
    _, h_sp, hpqrs = hamiltonian_of(mi)
    hp = pack_operator(h_sp, with_constant=True)
    eps = np.repeat(mi["orbital_energies"], 2)
    ops_f, theta_mp2, hf_init = get_cluster_ops_and_init_guess(n_el, [2.0] * n_el + [0.0] * (n - n_el), eps, hpqrs)
    rx, rz, rny, rc, own, exci, exci_len = [], [], [], [], [], [], []
    for j, op in enumerate(ops_f):
        p = pack_operator(transform_to_jw_basis(op))
        assert np.abs(p.cim).max() < 1e-15
        keep = p.cre != 0
        rx += p.x[keep].tolist(); rz += p.z[keep].tolist(); rny += p.ny[keep].tolist(); rc += p.cre[keep].tolist()
        own += [j] * int(keep.sum())
        q = list(map(int, op.terms[0].qbits))
        exci += q
        exci_len.append(len(q))
    meta = {"system": "H2O 6-31G (r = 1.0285 A, 0.538 pi; reference molecule_factory.py:138-148), O 1s frozen: (8e, 12o), 24 qubits",
            "hf_energy": mi["hf_energy"], "nuclear_repulsion_plus_core": mi["nuclear_repulsion"], "n_generators": len(ops_f),
            "n_rotations": len(rx), "n_terms": len(hp), "n_xmask_groups": int(len(set(hp.x.tolist())))}
    np.savez_compressed(os.path.join(OUT, "h2o_631g_24q.npz"), n=n, hf_init_sp=hf_integer(n_el, n),
                        ham_x=hp.x, ham_z=hp.z, ham_ny=hp.ny.astype(np.int8), ham_cre=hp.cre,
                        rot_x=np.array(rx, dtype=np.uint64), rot_z=np.array(rz, dtype=np.uint64),
                        rot_ny=np.array(rny, dtype=np.int8), rot_c=np.array(rc), rot_owner=np.array(own, dtype=np.int32),
                        exci=np.array(exci, dtype=np.int8), exci_len=np.array(exci_len, dtype=np.int8),
                        theta_mp2=np.array(theta_mp2), meta=json.dumps(meta))
    print(meta)


if __name__ == "__main__":
    if "--lih" in sys.argv or len(sys.argv) == 1:
        make_lih()
    if "--h2o" in sys.argv or len(sys.argv) == 1:
        make_h2o()
