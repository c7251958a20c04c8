"""Worker of tests/test_sharding_cpu.py::test_gloo_world2_host_logic (CPU, gloo, world size 2)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class StubEngine:
    """Stands in for the CUDA engine so that the SPLIT/GATHER logic can run without a GPU: the overlaps come from
    the CPU oracle (this is a test double; the product path never does this)."""
    n_global = 0

    def __init__(self, psi, sigma, ops):
        self.psi, self.sigma, self.ops, self.calls = psi, sigma, ops, []

    def pool_overlaps(self, pool, bra=1, ket=0):
        from oracle import statevector_oracle as orc
        n_ops = len(pool.offsets) - 1
        self.calls.append(n_ops)
        out = np.zeros(n_ops, dtype=np.complex128)
        for k in range(n_ops):
            v = np.zeros_like(self.psi)
            for t in range(pool.offsets[k], pool.offsets[k + 1]):
                v += complex(pool.cre[t], pool.cim[t]) * orc.apply_pauli(self.psi, int(pool.x[t]), int(pool.z[t]), int(pool.ny[t]))
            out[k] = np.vdot(self.sigma, v)
        return out


def main():
    import torch.distributed as dist
    from openvqe_b200 import sharded
    from openvqe_b200.lowering import pack_pool
    from tests.helpers import random_antihermitian, random_state
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    assert world == 2
    # all-gather + fixed-order sum: identical bits on every rank
    mine = np.array([1e16 if rank == 0 else -1e16 + 2.0, 0.1 * (rank + 1)])
    rows = sharded.allgather_f64(mine)
    assert rows.shape == (2, 2) and rows[rank].tolist() == mine.tolist()
    tot = sharded.sum_in_rank_order(rows)
    both = sharded.allgather_f64(tot)
    assert both[0].tolist() == both[1].tolist()
    # replica-mode pool sweep
    n = 6
    rng = np.random.default_rng(9)  # same on both ranks (SPMD)
    psi, sigma = random_state(rng, n), random_state(rng, n)
    ops = [random_antihermitian(rng, n, int(rng.integers(1, 5)), max_weight=4) for _ in range(11)]
    pool = pack_pool(ops)
    eng = StubEngine(psi, sigma, ops)
    split = sharded.replica_pool_overlaps(eng, pool)
    assert eng.calls == [6 if rank == 0 else 5]              # 11 operators over 2 ranks
    full = StubEngine(psi, sigma, ops).pool_overlaps(pool)
    assert np.array_equal(split, full)
    # the hot-path switch sees the initialised group
    from openvqe_b200 import _hotpath
    assert not _hotpath.replica_split_active(eng)           # replica mode is opt-in
    assert _hotpath.distributed_fd(lambda x: 0.0, [])[1] is None
    sharded.enable_replica()
    assert _hotpath.replica_split_active(eng)
    os.environ["VQE_B200_REPLICA_POOL"] = "0"
    assert not _hotpath.replica_split_active(eng)
    # ranks that disagree on the inputs are detected (and fall back to the unsplit path together)
    assert sharded.ranks_agree(b"same")
    assert not sharded.ranks_agree(b"rank%d" % rank)
    other = pack_pool(ops[:7] if rank == 0 else ops[1:8])
    eng2 = StubEngine(psi, sigma, ops)
    sharded.replica_pool_overlaps(eng2, other)
    assert eng2.calls == [7]                                # not split: every rank swept its own 7 operators
    # finite-difference gradients of a BFGS run spread over the ranks: same trajectory, same energies list
    import scipy.optimize
    os.environ.pop("VQE_B200_REPLICA_POOL", None)
    w = np.linspace(0.5, 1.5, 7)

    def action(x):
        return float(np.sum(w * np.cos(x - 0.3 * w)) + 0.05 * np.sum(x ** 4))

    x0 = np.linspace(-0.4, 0.6, 7)
    e_par, e_ser = [], []
    fun, jac = _hotpath.distributed_fd(action, e_par)
    assert jac is not None
    r_par = scipy.optimize.minimize(fun, x0=x0, jac=jac, method="BFGS", tol=1e-6)
    os.environ["VQE_B200_REPLICA_FD"] = "0"
    fun_s, jac_s = _hotpath.distributed_fd(action, e_ser)
    assert jac_s is None
    r_ser = scipy.optimize.minimize(fun_s, x0=x0, jac=None, method="BFGS", tol=1e-6)
    os.environ.pop("VQE_B200_REPLICA_FD")
    assert np.array_equal(r_par.x, r_ser.x) and r_par.fun == r_ser.fun and r_par.nit == r_ser.nit
    assert e_par == e_ser and len(e_par) > 20
    dist.barrier()
    if rank == 0:
        print("gloo worker ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
