"""Drop-in for reference ``openvqe/adapt/fermionic_adapt_vqe.py``.

Same function names, arguments, printed progress and returned dictionaries.  What
changes is where the arithmetic runs:

  reference                                         here
  ------------------------------------------------  ---------------------------------------------
  sig = H_sparse.dot(psi); |pool| sparse matvecs    one H|psi> kernel + one batched pool sweep
  build_ucc_ansatz + myQLM simulator (energies)     Pauli-rotation tile kernel + expectation kernel
  scipy expm_multiply per ansatz operator           exact generator exponential on the device
  dense eigh + sample loop (fidelity)               host eigh (n <= 14) / device Lanczos above + device overlap

The 2^n x 2^n scipy matrices (``hamiltonian_sparse``, ``cluster_ops_sparse``) of the main entry point are accepted
but never touched: everything is derived from the Pauli lists ``hamiltonian_sp`` / ``cluster_ops_sp`` (the same
operators, reference molecule_factory_with_sparse.py:339, :615).  The module-level helpers whose reference signature
takes ONLY matrices (``prepare_adapt_state``, ``compute_gradient_i``, ``return_gradient_list``) accept either the
Pauli-list operators or, up to 14 qubits, the reference's matrices -- a matrix is decomposed into its Pauli list once
(``lowering.operator_from_matrix``, cached per object) and then handled by the same kernels.
"""
import numpy as np
import scipy.optimize

from .. import _hotpath
from ..common_files.circuit import CircuitSummary, count, hf_gates, ucc_circuit
from ..common_files.sorted_gradient import abs_sort_desc, corresponding_index, index_without_0, value_without_0
from ..engine import BUF_PSI, get_engine
from ..lowering import as_operator, is_matrix, pack_operator

FIDELITY_MAX_QUBITS = 14  # dense eigh is O(8^n): above this the ground state comes from a Lanczos iteration on the device
LANCZOS_MAX_QUBITS = 31   # four resident vectors (psi, sigma, work, ground state): 4 x 34 GB at 31 qubits


def prepare_adapt_state(reference_ket, spmat_ops, parameters):
    """psi = prod_k exp(theta_k A_k) |ref>, exact exponential of each generator
    (reference fermionic_adapt_vqe.py:12-38).  ``spmat_ops``: the anti-Hermitian generators as Pauli-sum operators
    (cluster_ops_sp entries) or as the reference's scipy matrices (cluster_ops_sparse entries, n <= 14).  Returns the
    state as a dense column vector; it also stays resident on the device."""
    ket = reference_ket.toarray() if hasattr(reference_ket, "toarray") else np.asarray(reference_ket)
    n = int(np.log2(ket.reshape(-1).shape[0]))
    engine = get_engine(n)
    _hotpath.load_reference_ket(engine, ket)
    for k in range(len(parameters)):
        engine.apply_exp(_hotpath.packed(as_operator(spmat_ops[k])), float(parameters[k]))
    return engine.get_state().reshape(-1, 1)


def _gradients_on_device(engine, cluster_ops_sp, hamiltonian_sp):
    ov = _hotpath.pool_overlaps(engine, hamiltonian_sp, cluster_ops_sp)
    return _hotpath.snap_ties((2.0 * ov.real).tolist())


def compute_gradient_i(i, cluster_ops_sparse, v, sig=None, hamiltonian_sp=None):
    """g_i = 2 Re <sig| A_i |v> (reference :41-74) for one pool operator.  ``cluster_ops_sparse``: Pauli-sum operators or
    the reference's matrices; ``v``: state vector; ``sig``: H v as the caller computed it (reference call form), or
    None together with ``hamiltonian_sp`` (Pauli list or matrix) to let the engine form H v."""
    from ..engine import BUF_SIGMA
    op = as_operator(cluster_ops_sparse[i])
    engine = get_engine(op.nbqbits)
    _hotpath.load_reference_ket(engine, v)
    if sig is not None:
        s = sig.toarray() if hasattr(sig, "toarray") else np.asarray(sig)
        engine.set_state(np.asarray(s, dtype=np.complex128).reshape(-1), BUF_SIGMA)
        ov = engine.pool_overlaps(_hotpath.pack_pool([op]), bra=BUF_SIGMA, ket=BUF_PSI)
        gi = 2.0 * complex(ov[0])
        assert np.isclose(gi.imag, 0)  # as the reference (:72)
        return _hotpath.snap_ties([gi.real])[0]
    if hamiltonian_sp is None:
        raise TypeError("compute_gradient_i needs sig (= H v) or hamiltonian_sp")
    return _gradients_on_device(engine, [op], as_operator(hamiltonian_sp))[0]


def return_gradient_list(cluster_ops_sparse, hamiltonian_sparse, curr_state):
    """Whole-pool gradient sweep (reference :77-122): returns ``list_grad`` (|g_k|),
    ``curr_norm`` (sum g_k^2, not yet square-rooted), the signed maximum and its index.  Operators and Hamiltonian
    as Pauli lists (any size) or as the reference's scipy matrices (n <= 14)."""
    ham = as_operator(hamiltonian_sparse)
    if len(cluster_ops_sparse) and is_matrix(cluster_ops_sparse[0]):
        ops = _hotpath.operators_of_matrices(cluster_ops_sparse)
    else:
        ops = cluster_ops_sparse
    engine = get_engine(ham.nbqbits)
    if curr_state is not None:
        _hotpath.load_reference_ket(engine, curr_state)
    return _gradient_summary(_gradients_on_device(engine, ops, ham))


def _gradient_summary(grads):
    list_grad, curr_norm, next_deriv, next_index = [], 0, 0, 0
    for oi, gi in enumerate(grads):
        list_grad.append(abs(gi))
        curr_norm += gi * gi
        if abs(gi) > abs(next_deriv):
            next_deriv = gi
            next_index = oi
    return list_grad, curr_norm, next_deriv, next_index


def ucc_action(hamiltonian_sp, cluster_ops_sp, hf_init_sp, theta_current):
    """E(theta) of the Trotterised ansatz (reference :126-162)."""
    return _hotpath.ucc_energy(theta_current, hamiltonian_sp, cluster_ops_sp, hf_init_sp)


def print_gradient_lists_and_indices(list_grad):
    """reference :165-180"""
    mylist_value_without_0 = value_without_0(list_grad)
    mylist_index_without_0 = index_without_0(list_grad)
    sorted_mylist_value_without_0 = abs_sort_desc(value_without_0(list_grad))
    sorted_index = corresponding_index(mylist_value_without_0, mylist_index_without_0, sorted_mylist_value_without_0)
    return sorted_mylist_value_without_0, sorted_index


def prepare_hf_state(hf_init_sp, cluster_ops_sp):
    """reference :183-213 (binary_repr without width: MSB-first, no zero padding)."""
    n = cluster_ops_sp[0].nbqbits
    return CircuitSummary(n, hf_gates(n, hf_init_sp, padded=False))


def _circuit_state(engine, circuit):
    engine.set_basis_state(0)
    _hotpath.apply_gate_list(engine, circuit.gates)


def hf_energy(hf_state, hamiltonian_sp):
    """<HF|H|HF> (reference :216-238); ``hf_state`` is what prepare_hf_state returned."""
    engine = get_engine(hamiltonian_sp.nbqbits)
    _circuit_state(engine, hf_state)
    return float(engine.expectation(engine.paulisum(hamiltonian_sp)).real)


def commutators_calculations(cluster_ops_sp, hamiltonian_sp):
    """reference :241-270; needs operator products, i.e. real qat Hamiltonians."""
    out = []
    for oi in cluster_ops_sp:
        out.append(-(hamiltonian_sp * oi * (complex(0, 1)) - oi * (complex(0, 1)) * hamiltonian_sp))
    return out


def prepare_state_ansatz(cluster_ops_sp, hf_init_sp, parameters):
    """reference :273-306"""
    return ucc_circuit(cluster_ops_sp[0].nbqbits, cluster_ops_sp, hf_init_sp, parameters)


class _AnsatzCircuit(CircuitSummary):
    """CircuitSummary that also remembers the Pauli-rotation form of the ansatz so that the
    state can be prepared with the fused rotation kernel instead of gate by gate."""

    def __init__(self, summary, cluster_ops_sp, hf_init_sp, parameters):
        self.__dict__.update(summary.__dict__)
        self.rotation_form = (list(cluster_ops_sp), hf_init_sp, list(parameters))


def get_statevector(result, nbqbits):
    """reference :309-328 (kept for API compatibility: builds a dense vector from samples)."""
    statevector = np.zeros((2 ** nbqbits), np.complex128)
    for sample in result:
        statevector[sample.state.int] = sample.amplitude
    return statevector


def fun_fidelity(circ, eigenvalues, eigenvectors, nbqbits):
    """|<gs|psi_circ>|^2 (reference :331-361); the overlap is reduced on the device.  ``eigenvectors`` is the reference's
    dense eigenvector matrix or a ``ground_state.DeviceGroundState`` (vector resident in the engine's aux buffer)."""
    engine = get_engine(nbqbits)
    on_device = hasattr(eigenvectors, "fidelity")
    ee = None if on_device else eigenvectors[:, np.argmin(eigenvalues)]
    form = getattr(circ, "rotation_form", None)
    if form is not None:
        _hotpath.prepare_ucc_state(engine, form[0], form[1], form[2])
    else:
        _circuit_state(engine, circ)
    if on_device:
        return eigenvectors.fidelity(BUF_PSI)
    return abs(engine.overlap_host(ee, BUF_PSI)) ** 2


def fermionic_adapt_vqe(hamiltonian_sparse, cluster_ops_sparse, reference_ket, hamiltonian_sp, cluster_ops_sp,
                        hf_init_sp, n_max_grads, fci, optimizer, tolerance, type_conver, threshold_needed,
                        max_external_iterations=30):
    """Fermionic ADAPT-VQE loop, reference fermionic_adapt_vqe.py:371-593."""
    iterations = {"energies": [], "energies_substracted_from_FCI": [], "norms": [], "Max_gradients": [],
                  "fidelity": [], "CNOTs": [], "Hadamard": [], "RY": [], "RX": []}
    result = {}
    print("threshold needed for convergence", threshold_needed)
    print("Max_external_iterations:", max_external_iterations)
    print("how many maximum gradient are selected", n_max_grads)
    print("The optimizer method used:", optimizer)
    print("Tolerance for reaching convergence", tolerance)
    ansatz_ops, ansatz_gen, op_indices, parameters_ansatz = [], [], [], []
    nbqbits = hamiltonian_sp.nbqbits
    engine = get_engine(nbqbits)
    if nbqbits <= FIDELITY_MAX_QUBITS and hasattr(hamiltonian_sp, "get_matrix"):
        eigenvalues, eigenvectors = np.linalg.eigh(hamiltonian_sp.get_matrix())
    elif nbqbits <= LANCZOS_MAX_QUBITS and not getattr(engine, "n_global", 0):
        # lowest eigenpair of the HF state's symmetry sector by Lanczos on the device (ground_state.py)
        from ..ground_state import lanczos_ground_state
        eigenvectors = lanczos_ground_state(engine, hamiltonian_sp, int(hf_init_sp))
        eigenvalues = np.array([eigenvectors.energy])
    else:
        eigenvalues = eigenvectors = None  # fidelity reported as nan
    hf_state = prepare_hf_state(hf_init_sp, cluster_ops_sp)
    ref_energy = hf_energy(hf_state, hamiltonian_sp)
    print(ref_energy)
    print(" The reference energy of the molecular system is: %12.8f" % ref_energy)
    curr_state = hf_state
    prev_norm = 0.0
    for n_iter in range(0, max_external_iterations):
        print("\n\n\n")
        print(" --------------------------------------------------------------------------")
        print("                     Fermionic_ADAPT-VQE iteration: ", n_iter)
        print(" --------------------------------------------------------------------------")
        print(" Check gradient list chronological order")
        # exact-exponential state on the device, then sigma = H psi and the pool sweep
        _hotpath.load_reference_ket(engine, reference_ket)
        for gen, th in zip(ansatz_gen, parameters_ansatz):
            engine.apply_exp(_hotpath.packed(gen), float(th))
        list_grad, curr_norm, next_deriv, next_index = _gradient_summary(
            _gradients_on_device(engine, cluster_ops_sp, hamiltonian_sp))
        sorted_mylist_value_without_0, sorted_index = print_gradient_lists_and_indices(list_grad)
        curr_norm = np.sqrt(curr_norm)
        print(" Norm of the gradients in current iteration = %12.8f" % curr_norm)
        print(" Max gradient in current iteration= %12.8f" % next_deriv)
        print(" Index of the Max gradient in current iteration= ", next_index)
        if eigenvalues is not None:
            fid = fun_fidelity(curr_state, eigenvalues, eigenvectors, nbqbits)
        else:
            fid = float("nan")
        converged = False
        if type_conver == "norm":
            if curr_norm < threshold_needed:
                converged = True
        else:
            print(" type convergence is not defined")
            exit()
        if converged or (abs(curr_norm - prev_norm) < 10 ** (-8)):
            print("Convergence is done")
            result["indices"] = op_indices
            result["Number_operators"] = len(ansatz_ops)
            result["final_norm"] = curr_norm
            result["parameters"] = parameters_ansatz
            gates = curr_state.ops
            result["Number_CNOT_gates"] = count("CNOT", gates)
            result["Number_Hadamard_gates"] = count("H", gates)
            result["Number_RX_gates"] = count("_2", gates)
            print(" -----------Final ansatz----------- ")
            # as in the reference, opt_result is unbound if this happens at iteration 0 (:531)
            print(" *final converged energy iteration is %20.12f" % opt_result.fun)
            result["final_energy_last_iteration"] = opt_result.fun
            break
        chosen_batch = sorted_mylist_value_without_0
        gamma1, sorted_index1 = [], []
        curr_norm1 = 0
        for z in chosen_batch:
            curr_norm1 += z * z
        curr_norm1 = np.sqrt(curr_norm1)
        for i in range(n_max_grads):
            gamma1.append(chosen_batch[i] / curr_norm1)
            sorted_index1.append(sorted_index[i])
        print("sorted_index1: ", sorted_index1)
        for j in range(len(sorted_index1)):
            parameters_ansatz.append(0.01)
            ansatz_ops.append(complex(0.0, 1.0) * cluster_ops_sp[sorted_index1[j]])
            op_indices.append(sorted_index1[j])
            ansatz_gen.append(cluster_ops_sp[sorted_index1[j]])
        opt_result = scipy.optimize.minimize(
            lambda parameters: ucc_action(hamiltonian_sp, ansatz_ops, hf_init_sp, parameters),
            x0=parameters_ansatz, method=optimizer, tol=tolerance, options={"maxiter": 100000, "disp": True})
        xlist = opt_result.x
        print(" Finished energy iteration_i: %20.12f" % opt_result.fun)
        print(" -----------New ansatz created----------- ")
        print(" %4s \t%s \t%s" % ("#", "Coefficients", "Term"))
        parameters_ansatz = []
        for si in range(len(ansatz_ops)):
            print(" %4i \t%f \t%s" % (si, xlist[si], op_indices[si]))
            parameters_ansatz.append(xlist[si])
        curr_state = _AnsatzCircuit(prepare_state_ansatz(ansatz_ops, hf_init_sp, parameters_ansatz),
                                    ansatz_ops, hf_init_sp, parameters_ansatz)
        prev_norm = curr_norm
        gates = curr_state.ops
        iterations["energies"].append(opt_result.fun)
        iterations["energies_substracted_from_FCI"].append(abs(opt_result.fun - fci))
        iterations["norms"].append(curr_norm1)
        iterations["Max_gradients"].append(sorted_mylist_value_without_0[0])
        iterations["fidelity"].append(fid)
        iterations["CNOTs"].append(count("CNOT", gates))
        iterations["Hadamard"].append(count("H", gates))
        iterations["RY"].append(count("_4", gates))
        iterations["RX"].append(count("_2", gates))
    return iterations, result
