"""Drop-in for reference ``openvqe/ucc_family/get_energy_qucc.py`` (QUCCSD).

The ansatz is gate-defined (Yordanov efficient excitation circuits, reference
openvqe/common_files/circuit.py:13-106); the engine executes exactly those gates
on the GPU, fused tile by tile, and evaluates <H> with the grouped expectation
kernel.  Same class / method names, arguments and result dictionaries.
"""
import scipy.optimize

from .. import _hotpath
from ..common_files.circuit import CircuitSummary, count, hf_gates, quccsd_circuit


class EnergyUCC:
    def action_quccsd(self, theta_0, hamiltonian_sp, cluster_ops, hf_init_sp, energies=[]):
        """E(theta) of the QUCCSD circuit (reference get_energy_qucc.py:11-56).
        ``cluster_ops`` are the FERMIONIC operators: only ``terms[0].qbits`` (2 or 4
        spin-orbital indices) is read, as in the reference (:46-49)."""
        value = _hotpath.quccsd_energy(theta_0, hamiltonian_sp, cluster_ops, hf_init_sp)
        energies.append(value)
        return value

    def prepare_hf_state(self, hf_init_sp, cluster_ops_sp):
        """Hartree-Fock X-gate circuit (reference :58-89)."""
        n = cluster_ops_sp[0].nbqbits
        return CircuitSummary(n, hf_gates(n, hf_init_sp, padded=False))

    def prepare_state_ansatz(self, hamiltonian_sp, hf_init_sp, cluster_ops, theta):
        """Gate-level description of the QUCCSD circuit (reference :91-134)."""
        return quccsd_circuit(hamiltonian_sp.nbqbits, hf_init_sp, cluster_ops, theta)

    def get_energies(self, hamiltonian_sp, cluster_ops, hf_init_sp, theta_current1, theta_current2, FCI):
        """Two BFGS minimisations (theta0 = MP2 guess, theta0 = constant step),
        reference get_energy_qucc.py:136-244."""
        iterations = {
            "minimum_energy_result1_guess": [],
            "minimum_energy_result2_guess": [],
            "theta_optimized_result1": [],
            "theta_optimized_result2": [],
        }
        result = {}
        tolerance = 10 ** (-5)
        method = "BFGS"
        print("tolerance= ", tolerance)
        print("method= ", method)
        energies1, energies2 = [], []
        # under torchrun the finite-difference evaluations of every BFGS gradient are spread over the ranks
        # (identical trajectory and energies list, see _hotpath.distributed_fd); serially jac is None as in the reference
        fun1, jac1 = _hotpath.distributed_fd(
            lambda theta: _hotpath.quccsd_energy(theta, hamiltonian_sp, cluster_ops, hf_init_sp), energies1)
        fun2, jac2 = _hotpath.distributed_fd(
            lambda theta: _hotpath.quccsd_energy(theta, hamiltonian_sp, cluster_ops, hf_init_sp), energies2)
        opt_result1 = scipy.optimize.minimize(
            fun1, x0=theta_current1, jac=jac1, method=method, tol=tolerance, options={"maxiter": 50000, "disp": True})
        opt_result2 = scipy.optimize.minimize(
            fun2, x0=theta_current2, jac=jac2, method=method, tol=tolerance, options={"maxiter": 50000, "disp": True})
        theta_optimized_result1 = [opt_result1.x[k] for k in range(len(theta_current1))]
        theta_optimized_result2 = [opt_result2.x[k] for k in range(len(theta_current2))]
        circ1 = self.prepare_state_ansatz(hamiltonian_sp, hf_init_sp, cluster_ops, theta_optimized_result1)
        circ2 = self.prepare_state_ansatz(hamiltonian_sp, hf_init_sp, cluster_ops, theta_optimized_result2)
        iterations["minimum_energy_result1_guess"].append(opt_result1.fun)
        iterations["minimum_energy_result2_guess"].append(opt_result2.fun)
        iterations["theta_optimized_result1"].append(theta_optimized_result1)
        iterations["theta_optimized_result2"].append(theta_optimized_result2)
        result["CNOT1"] = count("CNOT", circ1.ops)
        result["CNOT2"] = count("CNOT", circ2.ops)
        result["len_op1"] = len(theta_optimized_result1)
        result["len_op2"] = len(theta_optimized_result2)
        result["energies_1"] = energies1
        result["energies_2"] = energies2
        result["energies1_substracted_from_FCI"] = abs(opt_result1.fun - FCI)
        result["energies2_substracted_from_FCI"] = abs(opt_result2.fun - FCI)
        return iterations, result
