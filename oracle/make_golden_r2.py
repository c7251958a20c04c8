"""TEST INFRASTRUCTURE ONLY -- round-2 additions to the committed fixtures (tests/golden/).

Run in the BUILD container (needs /root/reference; the GPU box has only the outputs):

    python oracle/make_golden_r2.py

Writes, next to the round-1 fixtures (which are left untouched):
  h6_full_pools.npz      the reference's own 12-qubit operator pools -- ``uccgsd`` (3 159 operators), ``spin_complement_gsd``
                         (714), ``singlet_upccgsd`` (k = 2) from openvqe/common_files/generator_excitations.py and the
                         285-operator YXXX qubit pool from qubit_pool.py, executed unmodified through oracle/qat_shim -- as
                         packed Pauli lists (x, z, ny, coefficient, offsets; term order preserved), and the outputs of the
                         unmodified reference ``return_gradient_list`` (fermionic_adapt_vqe.py:77-122) / ``calculate_gradient``
                         (qubit_adapt_vqe.py:126-150) for EVERY operator of those pools on the correlated H6 state of
                         h6_sto3g.json.gz.
  h4_get_energies.json   outputs of the unmodified reference QUCCSD driver ``EnergyUCC.get_energies``
                         (get_energy_qucc.py:136-244) on the H4/STO-3G instance of h4_sto3g.json.gz (both BFGS runs).
"""
from __future__ import annotations

import contextlib
import gzip
import io
import json
import os
import sys
import time

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("OPENVQE_REFERENCE", "/root/reference")
sys.path[:0] = [os.path.join(HERE, "qat_shim"), REF, ROOT]
OUT = os.path.join(ROOT, "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def ham_of(d, n):
    from qat.core import Term
    from qat.fermion import SpinHamiltonian
    c = d.get("constant", 0.0)
    const = complex(c[0], c[1]) if isinstance(c, list) else c
    if isinstance(const, complex) and const.imag == 0:
        const = const.real
    return SpinHamiltonian(n, [Term(complex(cr, ci) if ci else cr, op, qb) for cr, ci, op, qb in d["terms"]], const)


def pack(ops, n):
    from openvqe_b200.lowering import pack_pool
    p = pack_pool(ops)
    return {"x": p.x, "z": p.z, "ny": p.ny, "cre": p.cre, "cim": p.cim, "offsets": p.offsets}


def main():
    import openvqe.common_files.generator_excitations as gen
    from openvqe.adapt import fermionic_adapt_vqe as ref_fa
    from openvqe.adapt import qubit_adapt_vqe as ref_qa
    from openvqe.common_files.qubit_pool import QubitPool
    from openvqe.ucc_family.get_energy_qucc import EnergyUCC as RefQUCC
    from qat.core import Term
    from qat.fermion import SpinHamiltonian
    from tests.helpers import load_golden

    # ---------------- H6 STO-3G, 12 qubits: full pools ------------------------------------------------
    fx = load_golden("h6_sto3g.json.gz")
    n = 12
    h_sp = ham_of(fx["hamiltonian"], n)
    hs = h_sp.get_matrix(sparse=True)
    st = np.array(fx["state"]["state_re"]) + 1j * np.array(fx["state"]["state_im"])
    stc = scipy.sparse.csr_matrix(st.reshape(-1, 1))
    out = {}
    pools = {"uccgsd": quiet(gen.uccgsd, 6, 6, "JW")[2], "spin_complement_gsd": quiet(gen.spin_complement_gsd, 6, 6, "JW")[2],
             "singlet_upccgsd_k2": quiet(gen.singlet_upccgsd, 6, "JW", 1)[2]}
    for name, ops in pools.items():
        t0 = time.time()
        for k, v in pack(ops, n).items():
            out[name + "_" + k] = v
        mats = [o.get_matrix(sparse=True) for o in ops]
        lg, nrm, nd, ni = ref_fa.return_gradient_list(mats, hs, stc)
        out[name + "_list_grad"] = np.array([float(v) for v in lg])
        out[name + "_summary"] = np.array([float(nrm), float(nd), float(ni)])
        print("%-22s %5d operators, reference sweep %.1f s, max |g| %.6f at %d" % (name, len(ops), time.time() - t0, abs(nd), ni))
    _, yxxx = QubitPool().generate_yxxx_pool(n)
    for k, v in pack(yxxx, n).items():
        out["yxxx_" + k] = v
    out["yxxx_gradients"] = np.array([float(ref_qa.calculate_gradient(ref_qa.term_to_matrix_sparse(op), stc, hs)) for op in yxxx])
    np.savez_compressed(os.path.join(OUT, "h6_full_pools.npz"), **out)

    # ---------------- H4 STO-3G QUCCSD driver ---------------------------------------------------------
    fx = load_golden("h4_sto3g.json.gz")
    n = 8
    h_sp = ham_of(fx["hamiltonian"], n)

    class Fermi:
        def __init__(self, qbits):
            self.nbqbits, self.terms = n, [Term(1.0, "Cc" if len(qbits) == 2 else "CCcc", qbits)]

    ops_f = [Fermi(ex) for ex in fx["excitations"]]
    theta1, theta2 = list(fx["theta_mp2"]), [0.01] * len(ops_f)
    t0 = time.time()
    it, res = quiet(RefQUCC().get_energies, h_sp, ops_f, fx["hf_init_sp"], theta1, theta2, fx["fci"])
    print("reference QUCCSD get_energies: %.1f s, minima %.12f / %.12f, evaluations %d / %d" % (
        time.time() - t0, it["minimum_energy_result1_guess"][0], it["minimum_energy_result2_guess"][0],
        len(res["energies_1"]), len(res["energies_2"])))
    clean = lambda o: json.loads(json.dumps(o, default=lambda v: float(v) if isinstance(v, (np.floating, float)) else (
        int(v) if isinstance(v, np.integer) else [float(x) for x in v])))
    with open(os.path.join(OUT, "h4_get_energies.json"), "w") as f:
        json.dump({"theta_current1": theta1, "theta_current2": theta2, "iterations": clean(it), "result": clean(res)}, f)
    for fn in ("h6_full_pools.npz", "h4_get_energies.json"):
        print("  %-24s %8d bytes" % (fn, os.path.getsize(os.path.join(OUT, fn))))


if __name__ == "__main__":
    main()
