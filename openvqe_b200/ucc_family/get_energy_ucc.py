"""Drop-in for reference ``openvqe/ucc_family/get_energy_ucc.py``.

Same class, method names, arguments and returned dictionaries; the myQLM circuit
construction + CLinalg simulation inside ``ucc_action`` is replaced by the CUDA
engine (Pauli-rotation tile kernel + X-mask-grouped expectation kernel).
"""
import scipy.optimize

from .. import _hotpath
from ..common_files.circuit import count, ucc_circuit


class EnergyUCC:
    def ucc_action(self, theta_current, hamiltonian_sp, cluster_ops_sp, hf_init_sp, energies=[]):
        """E(theta) = <HF| U(theta)^+ H U(theta) |HF>, U = ordered product of
        exp(-i theta_j c_jk P_jk) (reference get_energy_ucc.py:8-50).  The shared
        mutable default ``energies`` is kept on purpose: the reference appends every
        objective value to it."""
        value = _hotpath.ucc_energy(theta_current, hamiltonian_sp, cluster_ops_sp, hf_init_sp)
        energies.append(value)
        return value

    def prepare_state_ansatz(self, hamiltonian_sp, cluster_ops_sp, hf_init_sp, parameters):
        """Gate-level description of the trial circuit (reference :52-90); the
        returned object has ``.ops`` usable by ``count``."""
        return ucc_circuit(hamiltonian_sp.nbqbits, cluster_ops_sp, hf_init_sp, parameters)

    def get_energies(self, hamiltonian_sp, cluster_ops_sp, pool_generator, hf_init_sp,
                     theta_current1, theta_current2, fci):
        """Two BFGS minimisations (fermionic generators, then qubit-pool generators),
        reference get_energy_ucc.py:92-206; same tolerances, options and result keys."""
        iterations = {
            "minimum_energy_result1_guess": [],
            "minimum_energy_result2_guess": [],
            "theta_optimized_result1": [],
            "theta_optimized_result2": [],
        }
        result = {}
        tolerance = 10 ** (-4)
        method = "BFGS"
        print("tolerance= ", tolerance)
        print("method= ", method)
        energies_1, energies_2 = [], []
        # under torchrun the finite-difference evaluations of every BFGS gradient are spread over the ranks
        # (identical trajectory and energies list, see _hotpath.distributed_fd); serially jac is None as in the reference
        fun1, jac1 = _hotpath.distributed_fd(
            lambda theta: _hotpath.ucc_energy(theta, hamiltonian_sp, cluster_ops_sp, hf_init_sp), energies_1)
        fun2, jac2 = _hotpath.distributed_fd(
            lambda theta: _hotpath.ucc_energy(theta, hamiltonian_sp, pool_generator, hf_init_sp), energies_2)
        if _hotpath.adjoint_enabled():  # opt-in: analytic gradients from one adjoint sweep (not the reference's trajectory)
            fun1, jac1 = _hotpath.adjoint_fun_jac(hamiltonian_sp, cluster_ops_sp, hf_init_sp, energies_1)
            fun2, jac2 = _hotpath.adjoint_fun_jac(hamiltonian_sp, pool_generator, hf_init_sp, energies_2)
        opt_result1 = scipy.optimize.minimize(
            fun1, x0=theta_current1, jac=jac1, method=method, tol=tolerance, options={"maxiter": 50000, "disp": True})
        opt_result2 = scipy.optimize.minimize(
            fun2, x0=theta_current2, jac=jac2, method=method, tol=tolerance, options={"maxiter": 50000, "disp": True})
        theta_optimized_result1 = [opt_result1.x[k] for k in range(len(theta_current1))]
        theta_optimized_result2 = [opt_result2.x[k] for k in range(len(theta_current2))]
        circ1 = self.prepare_state_ansatz(hamiltonian_sp, cluster_ops_sp, hf_init_sp, theta_optimized_result1)
        # the reference builds the second circuit from cluster_ops_sp too (:187-189)
        circ2 = self.prepare_state_ansatz(hamiltonian_sp, cluster_ops_sp, hf_init_sp, theta_optimized_result2)
        iterations["minimum_energy_result1_guess"].append(opt_result1.fun)
        iterations["minimum_energy_result2_guess"].append(opt_result2.fun)
        iterations["theta_optimized_result1"].append(theta_optimized_result1)
        iterations["theta_optimized_result2"].append(theta_optimized_result2)
        result["CNOT1"] = count("CNOT", circ1.ops)
        result["CNOT2"] = count("CNOT", circ2.ops)
        result["len_op1"] = len(theta_optimized_result1)
        result["len_op2"] = len(theta_optimized_result2)
        result["energies1_substracted_from_FCI"] = abs(opt_result1.fun - fci)
        result["energies2_substracted_from_FCI"] = abs(opt_result2.fun - fci)
        result["energies_1"] = energies_1
        result["energies_2"] = energies_2
        return iterations, result
