"""TEST INFRASTRUCTURE ONLY -- generates the committed fixtures under tests/golden/.

Run in the BUILD container (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

What it writes
  g1_h2_sto3g.json       transcription of the reference's input-complete golden
                         (notebooks/demo_WSSVQE.ipynb cells[5,9]): 15-term Hamiltonian + spectrum.
  notebook_pins.json     stored notebook outputs G2-G7 (energies, gradient norms, indices, gate counts),
                         parsed from the ``iterations are: {...}`` / ``results are: {...}`` lines.
  <system>.json.gz       Pauli-list problem instances (Hamiltonian from oracle/chem/hchain.py, operator pools
                         from the reference's own generator_excitations.py / qubit_pool.py executed through
                         oracle/qat_shim) together with OUTPUTS OF THE UNMODIFIED REFERENCE MODULES
                         (openvqe/ucc_family, openvqe/adapt) run through the shim on those inputs.

Every number a parity test compares against is produced here by reference code (or transcribed from
reference notebooks); the CUDA engine and the numpy restatement are both checked against these files.
"""
from __future__ import annotations

import ast
import contextlib
import gzip
import io
import json
import os
import re
import sys

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("OPENVQE_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "qat_shim"), REF, ROOT]
OUT = os.path.join(ROOT, "tests", "golden")


def _cell_text(nb, k):
    txt = ""
    for o in nb["cells"][k].get("outputs", []):
        t = o.get("text", o.get("data", {}).get("text/plain", []))
        txt += "".join(t)
    return txt


def _load_nb(name):
    with open(os.path.join(REF, "notebooks", name)) as f:
        return json.load(f)


def transcribe_g1():
    nb = _load_nb("demo_WSSVQE.ipynb")
    txt = _cell_text(nb, 5)
    const = None
    terms = []
    for line in txt.splitlines():
        m = re.match(r"\(([-+0-9.e]+)\+0j\) \* I\^4", line)
        if m:
            const = float(m.group(1))
            continue
        m = re.match(r"\(([-+0-9.e]+)\+0j\) \* \((\w+)\|\[([0-9, ]+)\]\)", line)
        if m:
            terms.append([float(m.group(1)), 0.0, m.group(2), [int(v) for v in m.group(3).split(",")]])
    eig_txt = _cell_text(nb, 9)
    m = re.search(r"\[(-0\.34365999[^\]]+)\]", eig_txt)
    eig = [float(v) for v in m.group(1).split()]
    assert const is not None and len(terms) == 14 and len(eig) == 16
    nuc = float(re.search(r"Nuclear repulsion =\s+([0-9.]+)", _cell_text(nb, 4)).group(1))
    return {"source": "reference notebooks/demo_WSSVQE.ipynb cells[4,5,9]", "nbqbits": 4, "constant": const,
            "terms": terms, "eigenvalues": eig, "nuclear_repulsion": nuc, "hf_index": 12}


def transcribe_pins():
    pins = {}
    for key, name in [("G2", "demo_puccgsd.ipynb"), ("G3", "demo_puccgsd_active_space.ipynb"),
                      ("G4", "demo_quccsd.ipynb"), ("G5", "demo_quccsd_active_space.ipynb"),
                      ("G6", "demo_fermionic_adapt.ipynb"), ("G7", "demo_qubit_adapt.ipynb")]:
        txt = _cell_text(_load_nb(name), 3)
        entry = {"source": "reference notebooks/%s cell[3]" % name}
        m = re.search(r"Hamiltonian info (\{.*?\})", txt)
        if m:
            entry["info"] = ast.literal_eval(m.group(1))
        m = re.search(r"^iterations are: (\{.*\})$", txt, re.M)
        if m:
            it = ast.literal_eval(m.group(1))
            entry["iterations"] = it
        m = re.search(r"^results are: (\{.*\})$", txt, re.M)
        if m:
            res = ast.literal_eval(m.group(1))
            # keep the first objective values only (the full traces are long)
            for k in ("energies_1", "energies_2"):
                if k in res:
                    res[k] = res[k][:40]
            entry["result"] = res
        m = re.search(r"sorted_mylist_value of gradient_without_0 (\[.*?\])", txt)
        if m:
            entry["first_sorted_gradients"] = ast.literal_eval(m.group(1))
        m = re.search(r"initial parameters (\[.*?\])", txt)
        if m:
            entry["first_initial_parameters"] = ast.literal_eval(m.group(1))
        pins[key] = entry
    return pins


# ---------------------------------------------------------------------------------------------
def ham_to_json(h):
    return {"nbqbits": h.nbqbits, "constant": [complex(h.constant_coeff).real, complex(h.constant_coeff).imag],
            "terms": [[complex(t.coeff).real, complex(t.coeff).imag, t.op, list(map(int, t.qbits))] for t in h.terms]}


def pool_to_json(ops):
    return [[[complex(t.coeff).real, complex(t.coeff).imag, t.op, list(map(int, t.qbits))] for t in op.terms]
            for op in ops]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def molecular_hamiltonian(n_atoms, r, basis, mo_signs=None):
    from oracle.chem.hchain import chain, molecular_integrals
    from qat.fermion import ElectronicStructureHamiltonian
    from qat.fermion.chemistry.ucc import convert_to_h_integrals, transform_integrals_to_new_basis
    from qat.fermion.transforms import transform_to_jw_basis
    mi = molecular_integrals(chain(n_atoms, r), basis)
    h1, h2 = mi["one_body"], mi["two_body"]
    if mo_signs is not None:
        h1, h2 = transform_integrals_to_new_basis(h1, h2, np.diag(np.asarray(mo_signs, float)))
    hpq, hpqrs = convert_to_h_integrals(h1, h2)
    hf = ElectronicStructureHamiltonian(hpq, hpqrs, constant_coeff=mi["nuclear_repulsion"])
    return mi, hf, transform_to_jw_basis(hf), hpqrs


def hf_integer(n_elec, nbqbits):
    v = 0
    for q in range(n_elec):
        v |= 1 << (nbqbits - 1 - q)
    return v


def fci_energy(h_sp, n_elec):
    import scipy.sparse.linalg as spla
    m = h_sp.get_matrix(sparse=True)
    n = h_sp.nbqbits
    idx = np.array([i for i in range(1 << n) if bin(i).count("1") == n_elec])
    sub = m[idx][:, idx]
    if sub.shape[0] <= 600:
        return float(np.linalg.eigvalsh(sub.toarray().real)[0])
    return float(spla.eigsh(sub.real, k=1, which="SA")[0][0])


def run_reference_fermionic_adapt(h_sp, ops_sp, hf, fci, optimizer, tol, thr, max_it):
    from openvqe.adapt.fermionic_adapt_vqe import fermionic_adapt_vqe
    n = h_sp.nbqbits
    ref_ket = scipy.sparse.csr_matrix(np.eye(1, 1 << n, hf).reshape(-1, 1).astype(complex))
    hs = h_sp.get_matrix(sparse=True)
    pool_sparse = [o.get_matrix(sparse=True) for o in ops_sp]
    try:
        it, res = quiet(fermionic_adapt_vqe, hs, pool_sparse, ref_ket, h_sp, ops_sp, hf, 1, fci, optimizer, tol,
                        "norm", thr, max_it)
    except UnboundLocalError:  # reference quirk: converged at iteration 0
        it, res = {}, {}
    return it, res


def main():
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "g1_h2_sto3g.json"), "w") as f:
        json.dump(transcribe_g1(), f, indent=1)
    with open(os.path.join(OUT, "notebook_pins.json"), "w") as f:
        json.dump(transcribe_pins(), f, indent=1)

    import openvqe.common_files.generator_excitations as gen
    from openvqe.adapt import fermionic_adapt_vqe as ref_fa
    from openvqe.adapt import qubit_adapt_vqe as ref_qa
    from openvqe.common_files.qubit_pool import QubitPool
    from openvqe.ucc_family.get_energy_qucc import EnergyUCC as RefQUCC
    from openvqe.ucc_family.get_energy_ucc import EnergyUCC as RefUCC
    from qat.fermion.chemistry.ucc_deprecated import get_cluster_ops_and_init_guess
    from oracle import statevector_oracle as orc

    rng = np.random.default_rng(2026)

    # ---------------- C1a / G2 / G6 / G7: H2 6-31G, 8 qubits --------------------------------------
    mi, h_f, h_sp, hpqrs = molecular_hamiltonian(2, 0.75, "6-31g", mo_signs=[1, 1, 1, -1])
    n, hf = 8, hf_integer(2, 8)
    fci = fci_energy(h_sp, 2)
    fx = {"system": "H2 6-31G r=0.75 A (reference molecule_factory.py:51-56), MO sign class of the notebooks",
          "hamiltonian": ham_to_json(h_sp), "hf_init_sp": hf, "hf_energy": mi["hf_energy"], "fci": fci}
    _, _, upcc_sp = quiet(gen.singlet_upccgsd, 4, "JW", 2)
    ansatz = [o * 1j for o in upcc_sp]
    fx["supccgsd_ansatz"] = pool_to_json(ansatz)
    thetas = [[0.01] * 18, list(rng.uniform(-0.3, 0.3, 18)), list(rng.uniform(-0.1, 0.1, 36)), [0.2, -0.1, 0.05]]
    fx["ucc_action"] = [{"theta": th, "energy": RefUCC().ucc_action(th, h_sp, ansatz, hf, [])} for th in thetas]
    circ = RefUCC().prepare_state_ansatz(h_sp, ansatz, hf, thetas[1])
    from openvqe.common_files.circuit import count
    fx["ucc_gate_counts"] = {"theta_index": 1, "CNOT": count("CNOT", circ.ops), "H": count("H", circ.ops),
                             "_2": count("_2", circ.ops), "_4": count("_4", circ.ops)}
    # fermionic ADAPT pool (spin_complement_gsd, 175 operators incl. identically-zero ones)
    _, _, scg_sp = quiet(gen.spin_complement_gsd, 2, 4, "JW")
    fx["spin_complement_gsd"] = pool_to_json(scg_sp)
    hs = h_sp.get_matrix(sparse=True)
    scg_sparse = [o.get_matrix(sparse=True) for o in scg_sp]
    ref_ket = scipy.sparse.csr_matrix(orc.basis_state(n, hf).reshape(-1, 1))
    lg, nrm, nd, ni = ref_fa.return_gradient_list(scg_sparse, hs, ref_ket)
    fx["gradients_at_hf"] = {"list_grad": [float(v) for v in lg], "curr_norm": float(nrm), "next_deriv": float(nd),
                             "next_index": int(ni)}
    # gradients + exact-exponential state at a non-trivial ansatz
    idxs, pars = [38, 32, 29], [-0.0214, -0.0251, -0.046]
    st = ref_fa.prepare_adapt_state(ref_ket, [scg_sparse[i] for i in idxs], pars)
    st = np.asarray(st.todense() if hasattr(st, "todense") else st).reshape(-1)
    lg2, nrm2, nd2, ni2 = ref_fa.return_gradient_list(scg_sparse, hs, scipy.sparse.csr_matrix(st.reshape(-1, 1)))
    fx["gradients_at_ansatz"] = {"indices": idxs, "parameters": pars, "state_re": st.real.tolist(),
                                 "state_im": st.imag.tolist(), "list_grad": [float(v) for v in lg2],
                                 "curr_norm": float(nrm2), "next_deriv": float(nd2), "next_index": int(ni2)}
    it, res = run_reference_fermionic_adapt(h_sp, scg_sp, hf, fci, "COBYLA", 1e-6, 1e-2, 35)
    fx["fermionic_adapt_run"] = {"iterations": it, "result": res, "options": ["COBYLA", 1e-6, "norm", 1e-2, 35]}
    # qubit ADAPT: 50-operator 'random' pool of the reference, seeded
    np.random.seed(7)
    _, _, sgsd_sp = quiet(gen.singlet_gsd, 2, 4, "JW")
    _, ops_f, _ = quiet(gen.singlet_gsd, 2, 4, "JW")
    qp = QubitPool()
    qubit_pool = quiet(qp.generate_pool, ops_f)
    _, pool_mix = quiet(qp.generate_pool_without_cluster, pool_type="random", nbqbits=n, qubit_pool=qubit_pool,
                        molecule_symbol="H2")
    fx["qubit_pool_random_seed7"] = pool_to_json(pool_mix)
    grads = [ref_qa.calculate_gradient(ref_qa.term_to_matrix_sparse(op), ref_ket, hs) for op in pool_mix]
    fx["qubit_gradients_at_hf"] = [float(g) for g in grads]
    ans_ops, ans_par = [pool_mix[i] for i in (3, 11, 27)], [0.27, -0.11, 0.06]
    stq = ref_qa.prepare_adapt_state(ref_ket, ans_ops, ans_par)
    stq = np.asarray(stq.todense()).reshape(-1)
    grads2 = [ref_qa.calculate_gradient(ref_qa.term_to_matrix_sparse(op), scipy.sparse.csr_matrix(stq.reshape(-1, 1)), hs)
              for op in pool_mix]
    fx["qubit_gradients_at_ansatz"] = {"indices": [3, 11, 27], "parameters": ans_par, "state_re": stq.real.tolist(),
                                       "state_im": stq.imag.tolist(), "gradients": [float(g) for g in grads2]}
    qa_out = quiet(ref_qa.qubit_adapt_vqe, h_sp, hs, ref_ket, n, pool_mix, hf, fci, n_max_grads=1,
                   adapt_conver="norm", adapt_thresh=1e-7, adapt_maxiter=4, tolerance_sim=1e-9, method_sim="BFGS")
    fx["qubit_adapt_run"] = {"iterations_sim": qa_out[0], "result_sim": qa_out[2], "adapt_maxiter": 4}
    with gzip.open(os.path.join(OUT, "h2_631g.json.gz"), "wt") as f:
        json.dump(fx, f)

    # ---------------- G4: H4 STO-3G QUCCSD, 8 qubits -----------------------------------------------
    mi, h_f, h_sp, hpqrs = molecular_hamiltonian(4, 0.85, "sto-3g")
    n, hf = 8, hf_integer(4, 8)
    eps = np.repeat(mi["orbital_energies"], 2)
    noons = [2.0] * 4 + [0.0] * 4
    ops_f, theta_mp2, hf_init = get_cluster_ops_and_init_guess(4, noons, eps, hpqrs)
    fx = {"system": "H4 STO-3G r=0.85 A (reference molecule_factory.py:57-68)", "hamiltonian": ham_to_json(h_sp),
          "hf_init_sp": hf, "hf_energy": mi["hf_energy"], "fci": fci_energy(h_sp, 4),
          "excitations": [list(map(int, op.terms[0].qbits)) for op in ops_f], "theta_mp2": theta_mp2}
    thetas = [[0.01] * len(ops_f), theta_mp2, list(rng.uniform(-0.4, 0.4, len(ops_f)))]
    fx["action_quccsd"] = [{"theta": list(map(float, th)), "energy": RefQUCC().action_quccsd(th, h_sp, ops_f, hf, [])}
                           for th in thetas]
    circ = RefQUCC().prepare_state_ansatz(h_sp, hf, ops_f, thetas[0])
    fx["cnot_count"] = count("CNOT", circ.ops)
    with gzip.open(os.path.join(OUT, "h4_sto3g.json.gz"), "wt") as f:
        json.dump(fx, f)

    # ---------------- C2/C3: H6 STO-3G r=1.5 (ADAPT factory geometry), 12 qubits -------------------
    mi, h_f, h_sp, hpqrs = molecular_hamiltonian(6, 1.5, "sto-3g")
    n, hf = 12, hf_integer(6, 12)
    fci = fci_energy(h_sp, 6)
    fx = {"system": "H6 STO-3G r=1.5 A (reference molecule_factory_with_sparse.py:101-113)",
          "hamiltonian": ham_to_json(h_sp), "hf_init_sp": hf, "hf_energy": mi["hf_energy"], "fci": fci}
    _, _, ugsd_sp = quiet(gen.uccgsd, 6, 6, "JW")
    _, _, scg_sp = quiet(gen.spin_complement_gsd, 6, 6, "JW")
    fx["pool_sizes"] = {"uccgsd": len(ugsd_sp), "spin_complement_gsd": len(scg_sp)}
    # keep the fixture small: every 9th uccgsd operator, every 3rd spin-complement operator
    sub_u = list(range(0, len(ugsd_sp), 9))
    sub_s = list(range(0, len(scg_sp), 3))
    fx["uccgsd_subset_indices"] = sub_u
    fx["uccgsd_subset"] = pool_to_json([ugsd_sp[i] for i in sub_u])
    fx["spin_complement_gsd_subset_indices"] = sub_s
    fx["spin_complement_gsd_subset"] = pool_to_json([scg_sp[i] for i in sub_s])
    hs = h_sp.get_matrix(sparse=True)
    ref_ket = scipy.sparse.csr_matrix(orc.basis_state(n, hf).reshape(-1, 1))
    # a correlated state: three exact generator exponentials on |HF>
    pick = [j for j in range(len(sub_u)) if any(abs(complex(t.coeff)) > 0 for t in ugsd_sp[sub_u[j]].terms)]
    gens = [ugsd_sp[sub_u[pick[k]]] for k in (5, 40, 77)]
    pars = [0.21, -0.13, 0.08]
    st = ref_fa.prepare_adapt_state(ref_ket, [g.get_matrix(sparse=True) for g in gens], pars)
    st = np.asarray(st.todense() if hasattr(st, "todense") else st).reshape(-1)
    fx["state"] = {"uccgsd_subset_positions": [pick[k] for k in (5, 40, 77)], "parameters": pars,
                   "state_re": st.real.tolist(), "state_im": st.imag.tolist()}
    stc = scipy.sparse.csr_matrix(st.reshape(-1, 1))
    for name, pool in (("uccgsd_subset", [ugsd_sp[i] for i in sub_u]),
                       ("spin_complement_gsd_subset", [scg_sp[i] for i in sub_s])):
        sparse_pool = [o.get_matrix(sparse=True) for o in pool]
        lg, nrm, nd, ni = ref_fa.return_gradient_list(sparse_pool, hs, stc)
        fx["gradients_" + name] = {"list_grad": [float(v) for v in lg], "curr_norm": float(nrm),
                                   "next_deriv": float(nd), "next_index": int(ni)}
    # qubit pool (single Pauli strings, YXXX order: 30 + 255 = 285 operators) gradients on the same state
    _, yxxx = QubitPool().generate_yxxx_pool(n)
    fx["yxxx_pool_size"] = len(yxxx)
    fx["yxxx_pool"] = pool_to_json(yxxx)
    fx["qubit_gradients"] = [float(ref_qa.calculate_gradient(ref_qa.term_to_matrix_sparse(op), stc, hs)) for op in yxxx]
    # Trotterised energies through the reference's ucc_action
    ans = [1j * g for g in gens]
    fx["ucc_action"] = [{"theta": th, "energy": ref_fa.ucc_action(h_sp, ans, hf, th)}
                        for th in ([0.01, 0.01, 0.01], pars)]
    with gzip.open(os.path.join(OUT, "h6_sto3g.json.gz"), "wt") as f:
        json.dump(fx, f)
    print("golden fixtures written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print("  %-28s %8d bytes" % (fn, os.path.getsize(os.path.join(OUT, fn))))


if __name__ == "__main__" and "--h12" not in sys.argv:
    main()


def make_h12():
    """24-qubit workload of bench.py / full-size parity tests: H12 chain, STO-3G, r = 1.0 A (a real molecular
    Hamiltonian of the size of config C4; H2O/6-31G itself needs p-type integrals, which this tooling lacks),
    UCCSD generator list (72 singles + 1746 doubles, Hermitian i(T-T^+) with real Pauli coefficients) and the
    matching QUCCSD excitation list + MP2 amplitudes.  Stored as packed bit masks (index-bit space)."""
    from qat.fermion.chemistry.ucc_deprecated import get_cluster_ops_and_init_guess
    from qat.fermion.transforms import transform_to_jw_basis
    from openvqe_b200.lowering import pack_operator
    n_at, n = 12, 24
    mi, h_f, h_sp, hpqrs = molecular_hamiltonian(n_at, 1.0, "sto-3g")
    hp = pack_operator(h_sp, with_constant=True)
    eps = np.repeat(mi["orbital_energies"], 2)
    ops_f, theta_mp2, hf_init = get_cluster_ops_and_init_guess(n_at, [2.0] * n_at + [0.0] * n_at, eps, hpqrs)
    rx, rz, rny, rc, own, exci, exci_len = [], [], [], [], [], [], []
    for j, op in enumerate(ops_f):
        sp = transform_to_jw_basis(op)  # operators already carry the factor i: Hermitian, real coefficients
        p = pack_operator(sp)
        assert np.abs(p.cim).max() < 1e-15
        keep = p.cre != 0
        rx += p.x[keep].tolist(); rz += p.z[keep].tolist(); rny += p.ny[keep].tolist(); rc += p.cre[keep].tolist()
        own += [j] * int(keep.sum())
        q = list(map(int, op.terms[0].qbits))
        exci += q
        exci_len.append(len(q))
    meta = {"system": "H12 chain STO-3G r=1.0 A, 24 qubits (C4-scale stand-in)", "hf_energy": mi["hf_energy"],
            "nuclear_repulsion": mi["nuclear_repulsion"], "n_generators": len(ops_f), "n_rotations": len(rx),
            "n_terms": len(hp), "n_xmask_groups": int(len(set(hp.x.tolist())))}
    np.savez_compressed(os.path.join(OUT, "h12_sto3g_24q.npz"), n=n, hf_init_sp=hf_integer(n_at, n),
                        ham_x=hp.x, ham_z=hp.z, ham_ny=hp.ny.astype(np.int8), ham_cre=hp.cre,
                        rot_x=np.array(rx, dtype=np.uint64), rot_z=np.array(rz, dtype=np.uint64),
                        rot_ny=np.array(rny, dtype=np.int8), rot_c=np.array(rc), rot_owner=np.array(own, dtype=np.int32),
                        exci=np.array(exci, dtype=np.int8), exci_len=np.array(exci_len, dtype=np.int8),
                        theta_mp2=np.array(theta_mp2), meta=json.dumps(meta))
    print(meta)


if __name__ == "__main__" and "--h12" in sys.argv:
    make_h12()
