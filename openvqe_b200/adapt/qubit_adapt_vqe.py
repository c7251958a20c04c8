"""Drop-in for reference ``openvqe/adapt/qubit_adapt_vqe.py``.

Same function names, arguments, printed progress and the 4-tuple of returned
dictionaries.  The reference rebuilds a 2^n x 2^n scipy matrix for every pool
operator in every ADAPT iteration (``term_to_matrix_sparse`` kron chains, :465) and
exponentiates 2^n x 2^n matrices with a sparse Pade ``expm`` (:52); here the pool is
lowered once to X/Z bit masks, sigma = H|psi> is formed once per iteration and the
whole pool is swept in one batched kernel; exp(-i theta A) of a Pauli string is a
rotation.  ``hamiltonian_sp_sparse`` is accepted and ignored.
"""
import numpy as np
import scipy.optimize

from .. import _hotpath
from ..common_files.circuit import CircuitSummary, count, hf_gates, ucc_circuit
from ..common_files.sorted_gradient import abs_sort_desc, corresponding_index, index_without_0, value_without_0
from ..engine import get_engine
from ..lowering import MATRIX_MAX_QUBITS, as_operator, matrix_of, pack_operator


def _hermitian_rotation_form(operator):
    """A = sum c_k P_k Hermitian with commuting strings -> exp(-i theta A) = prod_k
    exp(-i theta c_k P_k).  Qubit pools are single strings (reference qubit_pool.py)."""
    p = _hotpath.packed(operator)
    if np.any(p.cim != 0):
        raise ValueError("qubit-ADAPT operators must have real coefficients (Hermitian Pauli sums)")
    return p


def _apply_exp_minus_i(engine, operator, theta):
    """psi <- exp(-i theta A) psi for a Hermitian Pauli sum A (reference :44-55)."""
    p = _hermitian_rotation_form(operator)
    # exp(-i theta A) = exp(theta * (-i A)); -i A is anti-Hermitian with coefficients -i c_k
    from ..lowering import PackedTerms
    anti = PackedTerms(p.n, p.x, p.z, p.ny, np.zeros_like(p.cre), -p.cre)
    engine.apply_exp(anti, float(theta))


def prepare_adapt_state(reference_state, ansatz, coefficients):
    """psi = prod_k exp(-i theta_k A_k) |ref> (reference :20-55).  Returns a dense column."""
    ket = reference_state.toarray() if hasattr(reference_state, "toarray") else np.asarray(reference_state)
    n = int(np.log2(ket.reshape(-1).shape[0]))
    engine = get_engine(n)
    _hotpath.load_reference_ket(engine, ket)
    for i, operator in enumerate(ansatz):
        _apply_exp_minus_i(engine, operator, coefficients[i])
    return engine.get_state().reshape(-1, 1)


def term_to_matrix_sparse(spin_operator):
    """2^n x 2^n scipy CSR matrix of a pool operator (reference :81-123), built from its bit masks instead of a kron
    chain and tagged with the operator it came from, so ``calculate_gradient(term_to_matrix_sparse(op), state, H)``
    costs no decomposition.  Above 14 qubits the matrix is not built: the operator itself is handed through (the
    engine entry points take either)."""
    if not hasattr(spin_operator, "terms"):
        raise TypeError("term_to_matrix_sparse expects a Pauli-sum operator with .terms")
    if int(spin_operator.nbqbits) > MATRIX_MAX_QUBITS:
        return spin_operator
    return matrix_of(spin_operator)


def calculate_gradient(sparse_operator, state, sparse_hamiltonian):
    """2 |<state| H A |state>| (reference :126-150).  ``sparse_operator`` / ``sparse_hamiltonian``: the reference's
    scipy matrices (n <= 14; decomposed into Pauli lists once, cached) or the Pauli-sum operators themselves."""
    ham = as_operator(sparse_hamiltonian)
    op = as_operator(sparse_operator)
    engine = get_engine(ham.nbqbits)
    _hotpath.load_reference_ket(engine, state)
    ov = _hotpath.pool_overlaps(engine, ham, [op])
    return 2 * float(np.abs(ov[0]))


def prepare_state_ansatz(cluster_ops_sp, hf_init_sp, parameters):
    """reference :153-185"""
    return ucc_circuit(cluster_ops_sp[0].nbqbits, cluster_ops_sp, hf_init_sp, parameters)


def compute_commutator_i(commutator, curr_state):
    """<curr_state| commutator |curr_state> (reference :188-210)."""
    engine = get_engine(commutator.nbqbits)
    engine.set_basis_state(0)
    _hotpath.apply_gate_list(engine, curr_state.gates)
    return float(engine.expectation(engine.paulisum(commutator)).real)


def prepare_hf_state(hf_init_sp, cluster_ops_sp):
    """reference :213-243"""
    n = cluster_ops_sp[0].nbqbits
    return CircuitSummary(n, hf_gates(n, hf_init_sp, padded=False))


def hf_energy(hf_state, hamiltonian_sp):
    """reference :246-268"""
    engine = get_engine(hamiltonian_sp.nbqbits)
    engine.set_basis_state(0)
    _hotpath.apply_gate_list(engine, hf_state.gates)
    return float(engine.expectation(engine.paulisum(hamiltonian_sp)).real)


def ucc_action(hamiltonian_sp, cluster_ops_sp, hf_init_sp, theta_current):
    """reference :271-307"""
    return _hotpath.ucc_energy(theta_current, hamiltonian_sp, cluster_ops_sp, hf_init_sp)


def qubit_adapt_vqe(hamiltonian_sp, hamiltonian_sp_sparse, reference_ket, nqubits, pool_mix, hf_init_sp, fci,
                    n_max_grads=2, adapt_conver="norm", adapt_thresh=1e-08, adapt_maxiter=45,
                    tolerance_sim=1e-07, method_sim="BFGS"):
    """Qubit ADAPT-VQE loop, reference qubit_adapt_vqe.py:310-605."""
    iterations_sim = {"energies": [], "energies_substracted_from_fci": [], "norms": [], "Max_gradient": [],
                      "CNOTs": [], "Hadamard": [], "RY": [], "RX": []}
    result_sim = {}
    iterations_ana = {"energies": [], "energies_substracted_from_fci": [], "norms": [], "Max_gradient": []}
    result_ana = {}
    parameters_sim, parameters_ana, ansatz_ops = [], [], []
    engine = get_engine(hamiltonian_sp.nbqbits)
    curr_state = prepare_hf_state(hf_init_sp, pool_mix)
    ref_energy = hf_energy(curr_state, hamiltonian_sp)
    _hotpath.load_reference_ket(engine, reference_ket)
    ref_energy_ana = float(engine.expectation(engine.paulisum(hamiltonian_sp)).real)
    print("reference_energy from the simulator:", ref_energy)
    print("reference_energy from the analytical calculations:", ref_energy_ana)
    print(" --------------------------------------------------------------------------")
    print("                                                          ")
    print("                      Start Qubit ADAPT-VQE algorithm:")
    print("                                                          ")
    print(" --------------------------------------------------------------------------")
    print("                                                          ")
    Y = int(n_max_grads)
    print(" ------------------------------------------------------")
    print("        The number of maximum gradients inserted in each iteration:", Y)
    print(" ------------------------------------------------------")
    op_indices = []
    prev_norm = 0.0
    for n_iter in range(adapt_maxiter):
        print("\n")
        print(" --------------------------------------------------------------------------")
        print("                         Qubit ADAPT-VQE iteration: ", n_iter)
        print(" --------------------------------------------------------------------------")
        next_deriv = 0
        curr_norm = 0
        print("\n")
        print(" ------------------------------------------------------")
        print("        Start the analytical gradient calculation:")
        print(" ------------------------------------------------------")
        # exact state of the current ansatz on the device, then one batched pool sweep
        _hotpath.load_reference_ket(engine, reference_ket)
        for operator, th in zip(ansatz_ops, parameters_sim):
            _apply_exp_minus_i(engine, operator, th)
        ov = _hotpath.pool_overlaps(engine, hamiltonian_sp, pool_mix)
        list_grad = _hotpath.snap_ties((2.0 * np.abs(ov)).tolist())
        for gi in list_grad:
            curr_norm += gi * gi
            if abs(gi) > abs(next_deriv):
                next_deriv = gi
        mylist_value_without_0 = value_without_0(list_grad)
        mylist_index_without_0 = index_without_0(list_grad)
        sorted_mylist_value_without_0 = abs_sort_desc(value_without_0(list_grad))
        print("sorted_mylist_value of gradient_without_0", sorted_mylist_value_without_0)
        sorted_index = corresponding_index(mylist_value_without_0, mylist_index_without_0,
                                           sorted_mylist_value_without_0)
        curr_norm = np.sqrt(curr_norm)
        max_of_gi = next_deriv
        print(" Norm of <[H,A]> = %12.8f" % curr_norm)
        print(" Max  of <[H,A]> = %12.8f" % max_of_gi)
        converged = False
        if adapt_conver == "norm":
            if curr_norm < adapt_thresh:
                converged = True
        else:
            print(" FAIL: Convergence criterion not defined")
            exit()
        if converged or (abs(curr_norm - prev_norm) < 10 ** (-7)):
            print(" Ansatz Growth Converged!")
            result_sim["optimizer"] = method_sim
            result_sim["final_norm"] = curr_norm
            result_sim["indices"] = op_indices
            result_sim["len_operators"] = len(op_indices)
            result_sim["parameters"] = parameters_sim
            result_sim["final_energy"] = opt_result_sim.fun  # unbound at iteration 0, as in the reference (:508)
            print(" -----------Final ansatz----------- ")
            print(" %4s %12s %18s" % ("#", "Coeff", "Term"))
            for si in range(len(ansatz_ops)):
                print(" %4i %12.8f" % (si, parameters_sim[si]))
            break
        chosen_batch = sorted_mylist_value_without_0
        gamma1, sorted_index1 = [], []
        curr_norm1 = 0
        for z in chosen_batch:
            curr_norm1 += z * z
            curr_norm1 = np.sqrt(curr_norm1)  # sqrt INSIDE the loop: reference quirk (:529-532)
        for i in range(Y):
            gamma1.append(chosen_batch[i] / curr_norm1)
            sorted_index1.append(sorted_index[i])
        for m in range(len(gamma1)):
            parameters_sim.append(gamma1[m])
            parameters_ana.append(gamma1[m])
            ansatz_ops.append(pool_mix[sorted_index1[m]])
            op_indices.append(sorted_index1[m])
        print("initial parameters", parameters_sim)
        print("op_indices of iteration_%d" % n_iter, op_indices)
        opt_result_sim = scipy.optimize.minimize(
            lambda theta: ucc_action(hamiltonian_sp, ansatz_ops, hf_init_sp, theta),
            x0=parameters_sim, method=method_sim, tol=tolerance_sim, options={"maxiter": 100000, "disp": False})
        xlist_sim = opt_result_sim.x
        print(" ----------- ansatz from the simulator----------- ")
        print(" %s\t %s\t\t %s" % ("#", "Coeff", "Term"))
        parameters_sim = []
        for si in range(len(ansatz_ops)):
            print(" %i\t %f\t %s" % (si, xlist_sim[si], op_indices[si]))
            parameters_sim.append(xlist_sim[si])
        print(" Energy reached from the simulator: %20.20f" % opt_result_sim.fun)
        curr_state = prepare_state_ansatz(ansatz_ops, hf_init_sp, parameters_sim)
        prev_norm = curr_norm
        gates = curr_state.ops
        iterations_sim["energies"].append(opt_result_sim.fun)
        iterations_sim["energies_substracted_from_fci"].append(abs(opt_result_sim.fun - fci))
        iterations_sim["norms"].append(curr_norm)
        iterations_sim["Max_gradient"].append(sorted_mylist_value_without_0[0])
        iterations_sim["CNOTs"].append(count("CNOT", gates))
        iterations_sim["Hadamard"].append(count("H", gates))
        iterations_sim["RY"].append(count("_4", gates))
        iterations_sim["RX"].append(count("_2", gates))
    return iterations_sim, iterations_ana, result_sim, result_ana
