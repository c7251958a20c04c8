"""Worker of tests/test_sharded_gpu.py::test_two_process_ipc -- one rank of a 2-rank sharded state per process.
Checks the one-process-per-GPU plumbing (IPC handles over gloo, device-side flag barriers around peer passes,
rank-ordered partial sums) against the CPU oracle."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, default=1)
    args = ap.parse_args()
    import torch.distributed as dist
    from openvqe_b200.lowering import term_masks
    from openvqe_b200.sharded import ShardedEngine
    from oracle import statevector_oracle as orc
    from tests.helpers import random_hermitian, random_pauli, random_state
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    device = rank % args.devices
    n = 14
    eng = ShardedEngine(n, device)
    rng = np.random.default_rng(4242)  # same stream on every rank (SPMD)
    psi = random_state(rng, n)
    eng.set_state(psi)
    xs, zs, nys, angs = [], [], [], []
    ref = psi.copy()
    for k in range(40):
        op, qb = random_pauli(rng, n, max_weight=6)
        if k % 3 == 1:
            qb = sorted(set(qb) | {0})
            op = "".join(rng.choice(list("XYZ"), size=len(qb)))
        x, z, ny = term_masks(op, qb, n)
        a = float(rng.uniform(-0.5, 0.5))
        xs.append(x); zs.append(z); nys.append(ny); angs.append(a)
        ref = orc.pauli_rotation(ref, x, z, ny, a)
    eng.apply_rotations(xs, zs, nys, angs)
    eng.barrier()
    mine = eng.get_local_state()
    nl = n - 1
    err = np.max(np.abs(mine - ref[rank << nl:(rank + 1) << nl]))
    assert err < 1e-12, err
    assert np.max(np.abs(eng.get_state() - ref)) < 1e-12   # get_state all-gathers the full vector (reference-shaped helpers)
    assert abs(eng.norm2() - 1.0) < 1e-12
    ham = random_hermitian(rng, n, 200, max_weight=6, const=0.5)
    e = eng.expectation(eng.paulisum(ham))
    e_ref = orc.expectation(ref, ham)
    assert abs(e.real - e_ref) < 1e-11 and abs(e.imag) < 1e-11, (e, e_ref)
    # every rank holds bit-identical totals
    import openvqe_b200.sharded as sh
    rows = sh.allgather_f64([e.real])
    assert rows[0, 0] == rows[1, 0]
    # drop-in: after sharded.enable() the reference-shaped entry points run on the sharded state
    from openvqe_b200 import engine as engine_mod
    from openvqe_b200.ucc_family.get_energy_ucc import EnergyUCC
    from tests.helpers import jw_excitation
    del eng
    os.environ["VQE_B200_DEVICE"] = str(device)
    sh.enable(min_qubits=12)
    n2 = 12
    gens = []
    for _ in range(6):
        p_, q_, r_, s_ = sorted(rng.choice(n2, size=4, replace=False).tolist())
        gens.append(jw_excitation(n2, [r_, s_], [p_, q_]))
    gens.append(jw_excitation(n2, [7], [0]))  # a single that flips the global qubit
    ham2 = random_hermitian(rng, n2, 120, max_weight=6, const=-0.25)
    th = rng.uniform(-0.3, 0.3, size=len(gens)).tolist()
    hf2 = 0b111111000000
    e_api = EnergyUCC().ucc_action(th, ham2, gens, hf2, [])
    assert isinstance(engine_mod.get_engine(n2), sh.ShardedEngine)
    e_orc = orc.ucc_action(th, ham2, gens, hf2)
    assert abs(e_api - e_orc) < 1e-10, (e_api, e_orc)
    # exact exponential of a non-commuting generator on the sharded state (sigma and work vectors attached over IPC)
    from openvqe_b200.lowering import pack_operator
    from tests.helpers import random_antihermitian
    eng2 = engine_mod.get_engine(n2)
    psi2 = random_state(rng, n2)
    eng2.set_state(psi2)
    gen_nc = random_antihermitian(rng, n2, 5, max_weight=4)
    eng2.apply_exp(pack_operator(gen_nc), 0.9)
    ref2 = orc.fermionic_adapt_state(psi2, [gen_nc], [0.9])
    assert np.max(np.abs(eng2.get_state() - ref2)) < 1e-11
    sh.disable()
    dist.barrier()
    if rank == 0:
        print("sharded worker ok: |dpsi| = %.2e, E = %.12f (oracle %.12f)" % (err, e.real, e_ref), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
