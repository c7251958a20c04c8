"""Worker of tests/test_sharded_gpu.py::test_replica_mode_two_ranks -- SPMD replica mode on real GPUs (gloo group,
one process per rank; both ranks share GPU 0 when the box has a single GPU).  The distributed finite-difference
BFGS run and the split pool sweep must reproduce the single-process results bit for bit."""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch.distributed as dist
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    from openvqe_b200 import sharded
    sharded.enable_replica()   # opt-in: every rank runs the same problem in lockstep
    from openvqe_b200.adapt import fermionic_adapt_vqe as fa
    from openvqe_b200.ucc_family.get_energy_ucc import EnergyUCC
    from oracle import statevector_oracle as orc
    from tests.helpers import ham_from_json, load_golden, pool_from_json
    fx = load_golden("h2_631g.json.gz")
    ham = ham_from_json(fx["hamiltonian"])
    ops = pool_from_json(8, fx["supccgsd_ansatz"])[:6]
    theta0 = [0.01] * len(ops)

    def run():
        with contextlib.redirect_stdout(io.StringIO()):
            return EnergyUCC().get_energies(ham, ops, ops, fx["hf_init_sp"], theta0, theta0, fx.get("fci", -1.15))

    it_par, res_par = run()
    os.environ["VQE_B200_REPLICA_FD"] = "0"
    it_ser, res_ser = run()
    os.environ.pop("VQE_B200_REPLICA_FD")
    assert res_par["energies_1"] == res_ser["energies_1"] and len(res_par["energies_1"]) > 10
    assert it_par["theta_optimized_result1"] == it_ser["theta_optimized_result1"]
    # pool sweep split over the ranks == unsplit
    pool = pool_from_json(8, fx["spin_complement_gsd"])
    psi = orc.basis_state(8, fx["hf_init_sp"])
    split = fa.return_gradient_list(pool, ham, psi)
    os.environ["VQE_B200_REPLICA_POOL"] = "0"
    whole = fa.return_gradient_list(pool, ham, psi)
    os.environ.pop("VQE_B200_REPLICA_POOL")
    assert split[0] == whole[0] and split[3] == whole[3] == 38
    dist.barrier()
    if rank == 0:
        print("replica worker ok: %d energies, pool %d" % (len(res_par["energies_1"]), len(pool)), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
