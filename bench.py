#!/usr/bin/env python
"""bench.py -- UCC energy evaluations per second on the 24-qubit (C4-scale) workload, ADAPT pool-gradient sweep
time vs qubits, and the fraction of the HBM roofline the two hot kernels reach.

A "step" is ONE energy evaluation E(theta) of the Trotterised UCCSD ansatz on the 24-qubit instance BASELINE.json names
for config C4 -- H2O / 6-31G, O 1s frozen, (8e,12o): |HF> -> 11 008 Pauli rotations (1 424 generators) -> <H> over 8 921
Pauli terms in 1 567 X-mask groups -- on a 2^24 complex128 state (268 MB, larger than the 126 MB L2, so nothing is
L2-resident between sweeps).  theta changes every step.  `--molecule h12` selects the round-1 stand-in instead (H12
chain / STO-3G: 14 112 rotations, 1 818 generators, 14 905 terms in 2 767 groups); the default N = 1 run reports it under
`h12_standin` for continuity with round 1.

  value     evaluations/s with the Hamiltonian and rotation program resident in HBM (only the angles and the
            16-byte result cross PCIe), timed with CUDA events on the engine's stream, per-launch profiling OFF.
  e2e       the same metric through the reference-facing call EnergyUCC.ucc_action(theta, H, generators, hf)
            with host objects: lowering (cached), H2D of the operation descriptors, D2H of the energy, every step.
  roofline  dominant kernel class of the step (a separate profiled pass of the same steps, CUDA events around every
            launch): `achieved` = PHYSICAL bytes one launch moves (2*S for a rotation pass, S for an expectation pass;
            equal to ncu's dram__bytes, `traffic`) / average launch time; `frac` = achieved / measured HBM peak.
            The SURVEY 8d algorithmic figure (2*S per rotation, S per X-mask group) is reported beside it as
            `algorithmic_gbs` with the fusion factor (rotations or groups per pass) -- it exceeds the peak by
            construction and is NOT a bandwidth.
  N > 1     below 33 qubits the path shards as independent energy evaluations: every rank
            evaluates its own theta on its own GPU, no data-path collective ("weak" scaling).  The same line then also
            carries `sharded_c5`: the synthetic C5 program on a state SHARDED over the N ranks (33 + log2 N qubits =
            137 GB per GPU; one energy + one forward-difference gradient component; sharded == unsharded check at 30 q).

Beside the headline: `adapt_pool_sweep` (sigma = H psi + <sigma|A_k|psi> for the pool of the workload's generators at 24 qubits,
with the CPU port timed on a sample beside it), `adapt_pool_sweep_12q` (the reference-shaped return_gradient_list on
the H6 fixture against the oracle's scipy restatement of the reference code), `qubit_sweep` (energy evaluation and pool
sweep at 12...30 qubits, GPU and CPU port) and `quccsd` (the same excitations through EnergyUCC.action_quccsd).

`--impl reference` times the CPU port of the reference path (oracle/c, OpenMP over all host cores, thread count set
explicitly because torchrun exports OMP_NUM_THREADS=1): every timed step is ONE FULL evaluation of the same workload
(all rotations and all Hamiltonian terms, one 2^24 sweep each -- the cost structure of the reference's simulator).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# 24-qubit workloads (packed fixtures, tests/golden): the molecule BASELINE.json names for config C4 and the round-1 stand-in
MOLECULES = {
    "h2o": ("h2o_631g_24q.npz", "C4: H2O/6-31G, O 1s frozen, (8e,12o) active space, 24-qubit UCCSD energy evaluation"),
    "h12": ("h12_sto3g_24q.npz", "C4-scale stand-in of round 1: H12 chain/STO-3G, 24-qubit UCCSD energy evaluation"),
}
NCU_TRAFFIC = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")                      # ncu default: caches flushed before every kernel
NCU_TRAFFIC_IN_STREAM = os.path.join(ROOT, "profiles", "r2_ncu_traffic_in_stream.json")  # ncu --cache-control none


def load_workload(molecule=None):
    """The packed 24-qubit problem instance: default = the molecule BASELINE.json names (VQE_BENCH_MOLECULE / --molecule
    select the round-1 H12 stand-in instead)."""
    molecule = (molecule or os.environ.get("VQE_BENCH_MOLECULE", "h2o")).lower()
    fname, label = MOLECULES[molecule]
    z = np.load(os.path.join(ROOT, "tests", "golden", fname))
    w = {k: z[k] for k in z.files}
    w["n"] = int(w["n"])
    w["hf_init_sp"] = int(w["hf_init_sp"])
    w["meta"] = json.loads(str(w["meta"]))
    w["molecule"] = molecule
    w["label"] = "%s: %d Pauli rotations (%d generators) + <H> over %d terms / %d X-mask groups" % (
        label, len(w["rot_x"]), int(w["rot_owner"].max()) + 1, len(w["ham_x"]), len(set(w["ham_x"].tolist())))
    return w


def config_for(w, world):
    """The `config` object of the JSON line -- identical for the repo arm and the reference arm."""
    n = w["n"]
    return {"workload": w["label"], "molecule": w["molecule"], "qubits": n, "state_bytes": 16.0 * (1 << n),
            "l2_policy": "state larger than L2 (126 MB): 268 MB interleaved, 134 MB in the real layout a UCC evaluation keeps; every step "
                         "rewrites it (|HF>) first; consecutive passes walk their tiles in alternating directions, so a pass starts with "
                         "the lines the previous pass left in L2",
            "parallelism": "replicas: independent energy evaluations per GPU" if world > 1 else "1 GPU"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.proc = None
        for interval in ("50", "100"):  # 100 ms is the proven setting; 50 gives more samples inside a short timed region
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.FIELDS,
                                              "--format=csv,noheader,nounits", "-lms", interval],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except OSError:
                self.proc = None
                break
            time.sleep(0.15)
            if self.proc.poll() is None:
                break  # still looping: accepted
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            p = [v.strip() for v in line.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class Packed:
    pass


def packed_from(w, prefix):
    p = Packed()
    p.x = np.ascontiguousarray(w[prefix + "_x"], dtype=np.uint64)
    p.z = np.ascontiguousarray(w[prefix + "_z"], dtype=np.uint64)
    p.ny = np.ascontiguousarray(w[prefix + "_ny"], dtype=np.int32)
    if prefix == "ham":
        p.cre = np.ascontiguousarray(w["ham_cre"], dtype=np.float64)
        p.cim = np.zeros_like(p.cre)
    return p


def thetas_for(w, n_steps, rank):
    """A different parameter vector every step (MP2 amplitudes, scaled): nothing can be cached."""
    base = np.asarray(w["theta_mp2"], dtype=np.float64)
    n_gen = int(w["rot_owner"].max()) + 1
    base = np.resize(base, n_gen)
    base = np.where(base == 0.0, 0.01, base)
    return [base * (1.0 + 0.003 * (s + 1) + 0.0007 * rank) for s in range(n_steps)]


def pool_of(n, rot, owner, rc):
    """The workload's 1 818 generators as an ADAPT pool of anti-Hermitian operators T - T^dagger (i * real * P)."""
    from openvqe_b200.lowering import PackedTerms
    n_gen = int(owner.max()) + 1
    offs = np.zeros(n_gen + 1, dtype=np.int32)
    np.add.at(offs, owner + 1, 1)
    offs = np.cumsum(offs).astype(np.int32)
    return PackedTerms(n, rot.x, rot.z, rot.ny, np.zeros_like(rc), rc, offs)


# ------------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle/c): bounded samples and full evaluations
# ------------------------------------------------------------------------------------------------
def cpu_sample(w, n_rot_sample=48, n_term_sample=96):
    """Time the CPU port on a bounded sample of the energy evaluation and scale to one full evaluation."""
    from oracle import c_oracle
    cores = c_oracle.set_threads()
    n = w["n"]
    rot, ham = packed_from(w, "rot"), packed_from(w, "ham")
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[w["hf_init_sp"]] = 1.0
    th = thetas_for(w, 1, 0)[0]
    angles = th[w["rot_owner"]] * w["rot_c"]
    # spread the sample over the program so that it sees the same mix of X-masks
    ridx = np.linspace(0, len(angles) - 1, n_rot_sample).astype(int)
    tidx = np.linspace(0, len(ham.x) - 1, n_term_sample).astype(int)
    c_oracle.apply_rotations(psi, n, rot.x[ridx[:2]], rot.z[ridx[:2]], rot.ny[ridx[:2]], angles[ridx[:2]])  # warm-up
    t0 = time.perf_counter()
    c_oracle.apply_rotations(psi, n, rot.x[ridx], rot.z[ridx], rot.ny[ridx], angles[ridx])
    t_rot = (time.perf_counter() - t0) / n_rot_sample
    t0 = time.perf_counter()
    c_oracle.expectation(psi, n, ham.x[tidx], ham.z[tidx], ham.ny[tidx], ham.cre[tidx], ham.cim[tidx])
    t_term = (time.perf_counter() - t0) / n_term_sample
    est = len(angles) * t_rot + len(ham.x) * t_term
    return {"seconds_per_eval": est, "t_rotation_s": t_rot, "t_term_s": t_term, "cores": cores,
            "sample": "%d of %d rotations + %d of %d Hamiltonian terms at 24 qubits (one 2^24 sweep each), scaled "
                      "linearly to one evaluation" % (n_rot_sample, len(angles), n_term_sample, len(ham.x))}


def cpu_pool_sample(n, psi, ham, pool, n_term_sample=64, n_op_sample=24):
    """CPU port of the ADAPT sweep (reference fermionic_adapt_vqe.py:77-122: sigma = H psi once, then one overlap per
    pool operator) on a bounded sample, scaled to the whole Hamiltonian / pool."""
    from oracle import c_oracle
    from openvqe_b200.lowering import PackedTerms
    cores = c_oracle.set_threads()
    tidx = np.linspace(0, len(ham.x) - 1, min(n_term_sample, len(ham.x))).astype(int)
    sub = Packed()
    sub.x, sub.z, sub.ny, sub.cre, sub.cim = (np.ascontiguousarray(a[tidx]) for a in (ham.x, ham.z, ham.ny, ham.cre, ham.cim))
    t0 = time.perf_counter()
    sig = c_oracle.apply_paulisum(psi, n, sub)
    t_sigma = (time.perf_counter() - t0) / len(tidx) * len(ham.x)
    n_ops = len(pool.offsets) - 1
    pick = np.unique(np.linspace(0, n_ops - 1, min(n_op_sample, n_ops)).astype(int))
    lo, hi = pool.offsets[pick], pool.offsets[pick + 1]
    idx = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)])
    soffs = np.concatenate([[0], np.cumsum(hi - lo)]).astype(np.int32)
    subpool = PackedTerms(n, pool.x[idx], pool.z[idx], pool.ny[idx], pool.cre[idx], pool.cim[idx], soffs)
    t0 = time.perf_counter()
    c_oracle.pool_overlaps(sig, psi, n, subpool)
    t_pool = (time.perf_counter() - t0) / len(idx) * len(pool.x)
    return {"seconds": t_sigma + t_pool, "sigma_s": t_sigma, "overlaps_s": t_pool, "cores": cores, "kind": "port",
            "sample": "%d of %d Hamiltonian terms (sigma = H psi) + %d of %d pool operators, one 2^%d sweep per Pauli string, "
                      "scaled linearly" % (len(tidx), len(ham.x), len(pick), n_ops, n)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import c_oracle
    cores = c_oracle.set_threads()
    w = load_workload(args.molecule)
    n = w["n"]
    rot, ham = packed_from(w, "rot"), packed_from(w, "ham")
    owner, rc = w["rot_owner"], np.asarray(w["rot_c"], dtype=np.float64)
    ths = thetas_for(w, args.warmup + args.steps, 0)
    psi = np.empty(1 << n, dtype=np.complex128)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(w, 8, 8)  # warm-up: page in the state, spin up the OpenMP team
    budget = float(os.environ.get("VQE_REF_BUDGET_S", "420"))
    secs, energies = [], []
    t_start = time.perf_counter()
    for k in range(args.steps):
        if secs and (time.perf_counter() - t_start) + max(secs) > budget:
            break  # keep the whole run within a few minutes; the steps actually timed are reported
        t0 = time.perf_counter()
        e = c_oracle.ucc_energy(n, w["hf_init_sp"], rot, ths[args.warmup + k][owner] * rc, ham, psi)
        secs.append(time.perf_counter() - t0)
        energies.append(e)
    sec = float(np.mean(secs))
    val = 1.0 / sec
    line = {"impl": "reference", "metric": "ucc_energy_evals_per_s", "value": val, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": len(secs), "steps_requested": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128 state)",
            "data": "synthetic", "config": config_for(w, world), "energy_first_step": energies[0],
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": cores, "kind": "port",
                             "sample": "every timed step is one FULL evaluation (%d rotations + %d terms, one 2^%d sweep each); "
                                       "CPU port = oracle/c/vqe_oracle.c (OpenMP, %d threads set explicitly), the reference's own myQLM "
                                       "simulator is not installable here" % (len(rot.x), len(ham.x), n, cores),
                             "seconds_per_step": secs, "wall_s": time.perf_counter() - t_start},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def build_host_objects(w):
    """Duck-typed qat objects (what the reference hands to EnergyUCC.ucc_action) from the packed fixture."""
    n = w["n"]

    class Term:
        __slots__ = ("coeff", "op", "qbits")

        def __init__(self, c, op, qb):
            self.coeff, self.op, self.qbits = c, op, qb

    class Ham:
        def __init__(self, terms, const=0.0):
            self.nbqbits, self.terms, self.constant_coeff = n, terms, const

    def term(x, z, c):
        op, qb = [], []
        for q in range(n):
            b = n - 1 - q
            xb, zb = (int(x) >> b) & 1, (int(z) >> b) & 1
            if xb or zb:
                op.append("Y" if xb and zb else ("X" if xb else "Z"))
                qb.append(q)
        return Term(float(c), "".join(op), qb)

    hterms, const = [], 0.0
    for x, z, c in zip(w["ham_x"], w["ham_z"], w["ham_cre"]):
        if x == 0 and z == 0:
            const += float(c)
        else:
            hterms.append(term(x, z, c))
    ham = Ham(hterms, const)
    gens = [[] for _ in range(int(w["rot_owner"].max()) + 1)]
    for x, z, c, o in zip(w["rot_x"], w["rot_z"], w["rot_c"], w["rot_owner"]):
        gens[int(o)].append(term(x, z, c))
    return ham, [Ham(t) for t in gens]


def pool_sweep_12q():
    """BASELINE.md section 3 item 1 at the size the reference can run: the reference-shaped return_gradient_list on the
    12-qubit H6 fixture (spin-complement GSD pool) against the oracle's restatement of the reference's scipy code."""
    from tests.helpers import ham_from_json, load_golden, pool_from_json
    from oracle import statevector_oracle as orc
    from openvqe_b200.adapt import fermionic_adapt_vqe as fa
    fx = load_golden("h6_sto3g.json.gz")
    ham = ham_from_json(fx["hamiltonian"])
    key = "spin_complement_gsd" if "spin_complement_gsd" in fx else "spin_complement_gsd_subset"
    pool = pool_from_json(12, fx[key])
    psi = np.array(fx["state"]["state_re"]) + 1j * np.array(fx["state"]["state_im"])
    fa.return_gradient_list(pool, ham, psi)  # warm-up: lowering + upload (cached afterwards)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        lg, nrm, nd, ni = fa.return_gradient_list(pool, ham, psi)
    t_gpu = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    g_cpu = orc.fermionic_pool_gradients(psi, ham, pool)
    t_cpu = time.perf_counter() - t0
    return {"qubits": 12, "pool": key, "pool_size": len(pool), "api": "openvqe_b200.adapt.fermionic_adapt_vqe.return_gradient_list",
            "seconds": t_gpu, "max_abs_diff_vs_oracle": float(np.max(np.abs(np.array(lg) - np.abs(g_cpu)))),
            "cpu_baseline": {"seconds": t_cpu, "cores": 1, "kind": "port",
                             "sample": "oracle/statevector_oracle.fermionic_pool_gradients: numpy restatement of reference "
                                       "fermionic_adapt_vqe.py:77-122 (sigma = H psi, one operator application + dot per pool operator), whole pool"}}


def qubit_sweep(device, sizes, with_cpu=True):
    """Energy evaluation and ADAPT pool sweep vs register size on the synthetic C5 program (tools/c5_synthetic.py:
    256 JW generators = 1 664 Pauli rotations, 64-group Hamiltonian): GPU times, and the CPU port on a bounded sample."""
    from tools import c5_synthetic as c5
    from openvqe_b200.engine import BUF_PSI, BUF_SIGMA, Engine
    from openvqe_b200.lowering import PackedTerms
    out = []
    for n in sizes:
        gen, ham = c5.generators(n), c5.hamiltonian(n)
        S = 16.0 * (1 << n)
        eng = Engine(n, device=device)
        hp = PackedTerms(n, ham["x"], ham["z"], ham["ny"], ham["cre"], np.zeros_like(ham["cre"]))
        ps = eng.paulisum(hp)
        owner, coeff = gen["owner"], gen["coeff"]
        ang = gen["theta"][owner] * coeff
        hf = c5.hf_index(n)
        offs = np.zeros(gen["n_generators"] + 1, dtype=np.int32)
        np.add.at(offs, owner + 1, 1)
        offs = np.cumsum(offs).astype(np.int32)
        pool = PackedTerms(n, gen["x"], gen["z"], gen["ny"], np.zeros(len(coeff)), np.asarray(coeff, dtype=np.float64), offs)

        def energy(scale):
            eng.set_basis_state(hf)
            eng.apply_rotations(gen["x"], gen["z"], gen["ny"], ang * scale)
            return eng.expectation(ps).real

        def sweep():
            eng.apply_paulisum(ps, dst=BUF_SIGMA, src=BUF_PSI)
            return eng.pool_overlaps(pool, bra=BUF_SIGMA, ket=BUF_PSI)

        energy(1.0)
        sweep()
        eng.synchronize()
        reps = 3 if n <= 28 else 2
        eng.timer_begin()
        for k in range(reps):
            e = energy(1.0 + 0.01 * (k + 1))
        t_e = eng.timer_end() / reps / 1e3
        eng.timer_begin()
        for k in range(reps):
            ov = sweep()
        t_s = eng.timer_end() / reps / 1e3
        row = {"qubits": n, "state_bytes": S, "energy_eval_s": t_e, "energy_evals_per_s": 1.0 / t_e, "pool_sweep_s": t_s,
               "rotations": int(np.count_nonzero(ang)), "hamiltonian_terms": int(len(ham["x"])), "hamiltonian_groups": int(ham["n_groups"]),
               "pool_size": int(gen["n_generators"]), "energy": e, "max_abs_gradient": float(np.max(np.abs(2.0 * ov.real)))}
        if with_cpu:
            from oracle import c_oracle
            cores = c_oracle.set_threads()
            psi = np.zeros(1 << n, dtype=np.complex128)
            psi[hf] = 1.0
            k_r = 8 if n <= 24 else 3
            ridx = np.linspace(0, len(ang) - 1, k_r).astype(int)
            t0 = time.perf_counter()
            c_oracle.apply_rotations(psi, n, gen["x"][ridx], gen["z"][ridx], gen["ny"][ridx], ang[ridx])
            t_rot = (time.perf_counter() - t0) / k_r
            hpk = Packed()
            hpk.x, hpk.z, hpk.ny, hpk.cre, hpk.cim = hp.x, hp.z, hp.ny, hp.cre, hp.cim
            k_t = 8 if n <= 24 else 3
            tidx = np.linspace(0, len(hp.x) - 1, k_t).astype(int)
            t0 = time.perf_counter()
            c_oracle.expectation(psi, n, hp.x[tidx], hp.z[tidx], hp.ny[tidx], hp.cre[tidx], hp.cim[tidx])
            t_term = (time.perf_counter() - t0) / k_t
            cp = cpu_pool_sample(n, psi, hpk, pool, n_term_sample=k_t, n_op_sample=max(2, k_t // 2))
            row["cpu_port"] = {"energy_eval_s": int(np.count_nonzero(ang)) * t_rot + len(hp.x) * t_term, "pool_sweep_s": cp["seconds"],
                               "cores": cores, "sample": "%d rotations + %d terms + pool sample, one 2^%d sweep each, scaled linearly" % (k_r, k_t, n)}
            del psi
        out.append(row)
        del eng, ps
    return out


def quick_value(molecule, device, steps, warmup):
    """Device-timed evaluations/s of another 24-qubit instance (same step as the headline, resident inputs)."""
    from openvqe_b200.engine import Engine
    from openvqe_b200.lowering import PackedTerms
    w = load_workload(molecule)
    n = w["n"]
    eng = Engine(n, device=device)
    hp, rot = packed_from(w, "ham"), packed_from(w, "rot")
    ps = eng.paulisum(PackedTerms(n, hp.x, hp.z, hp.ny, hp.cre, hp.cim))
    owner, rc = w["rot_owner"], np.asarray(w["rot_c"], dtype=np.float64)
    ths = thetas_for(w, warmup + steps, 0)
    energies = []

    def step(theta):
        eng.set_basis_state(w["hf_init_sp"])
        eng.apply_rotations(rot.x, rot.z, rot.ny, theta[owner] * rc)
        return eng.expectation(ps).real

    for th in ths[:warmup]:
        step(th)
    eng.synchronize()
    l0 = eng.launch_count
    eng.timer_begin()
    for th in ths[warmup:]:
        energies.append(step(th))
    ms = eng.timer_end()
    return {"workload": w["label"], "ms_per_step": ms / steps, "evals_per_s": steps / (ms / 1e3),
            "gpu_launches": int(eng.launch_count - l0), "energy_first_step": energies[0]}


def run_ours(args, rank, world, local_rank):
    from openvqe_b200.engine import BUF_PSI, BUF_SIGMA, Engine
    from openvqe_b200.lowering import PackedTerms
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = load_workload(args.molecule)
    n = w["n"]
    S = 16.0 * (1 << n)
    eng = Engine(n, device=local_rank)
    hp = packed_from(w, "ham")
    ps = eng.paulisum(PackedTerms(n, hp.x, hp.z, hp.ny, hp.cre, hp.cim))
    rot = packed_from(w, "rot")
    owner, rc = w["rot_owner"], np.asarray(w["rot_c"], dtype=np.float64)
    hf = w["hf_init_sp"]

    def step(theta):
        eng.set_basis_state(hf)
        eng.apply_rotations(rot.x, rot.z, rot.ny, theta[owner] * rc)
        return eng.expectation(ps).real

    def barrier():
        eng.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ths = thetas_for(w, args.warmup + args.steps, rank)
    for th in ths[:args.warmup]:
        step(th)
    # layout of the state during the step: a purely real state is kept as 2^n doubles (half the bytes per pass)
    eng.set_basis_state(hf)
    eng.apply_rotations(rot.x, rot.z, rot.ny, ths[0][owner] * rc)
    real_layout = eng.real_layout
    S_phys = S / 2.0 if real_layout else S
    # ---- timed: resident inputs, per-launch profiling off ---------------------------------------------
    eng.profile(False)
    eng.transfer_bytes(reset=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    l0 = eng.launch_count
    eng.timer_begin()
    energies = [step(th) for th in ths[args.warmup:]]
    ms = eng.timer_end()
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count - l0
    ms = max_over_ranks(ms)
    value = world * args.steps / (ms / 1e3)
    # ---- the same steps once more with CUDA events around every launch: per-kernel-class time for the roofline ----
    for k in range(6):
        eng.profile_read(k, reset=True)
    eng.profile(True)
    eng.timer_begin()
    for th in ths[args.warmup:]:
        step(th)
    ms_prof = eng.timer_end()
    prep_ms, prep_n = eng.profile_read(0)
    exp_ms, exp_n = eng.profile_read(1)
    eng.profile(False)
    # ---- second half of the BASELINE metric: ADAPT pool-gradient sweep time ------------------------
    # sigma = H psi once, then <sigma|A_k|psi> for every operator of the pool in one batched sweep (reference
    # fermionic_adapt_vqe.py:77-122).  Pool = the 1 818 UCCSD generators of the workload (as T - T^dagger);
    # psi = the state of the last evaluation.  Not part of the timed steps above.
    pool_sweep = None
    if rank == 0 and world == 1 and not args.no_pool:
        n_gen = int(owner.max()) + 1
        pool = pool_of(n, rot, owner, rc)
        eng.apply_paulisum(ps, dst=BUF_SIGMA, src=BUF_PSI)
        eng.pool_overlaps(pool, bra=BUF_SIGMA, ket=BUF_PSI)  # warm-up
        eng.synchronize()
        for k in range(6):
            eng.profile_read(k, reset=True)
        eng.profile(True)
        t0 = time.perf_counter()
        eng.apply_paulisum(ps, dst=BUF_SIGMA, src=BUF_PSI)
        ov = eng.pool_overlaps(pool, bra=BUF_SIGMA, ket=BUF_PSI)
        sweep_s = time.perf_counter() - t0
        ap_ms, ap_n = eng.profile_read(2)
        po_ms, po_n = eng.profile_read(3)
        eng.profile(False)
        grads = 2.0 * ov.real
        pool_groups = len(set(zip(owner.tolist(), rot.x.tolist())))
        peak_p, _ = peaks()
        pool_sweep = {"qubits": n, "pool_size": n_gen, "pool_strings": int(len(rot.x)), "seconds": sweep_s,
                      "sigma_ms": ap_ms, "sigma_passes": ap_n, "sweep_ms": po_ms, "sweep_passes": po_n,
                      # physical: a sigma pass reads psi and reads + writes sigma (2*S on the first pass), a sweep pass reads both
                      "sigma_physical_gbs": (3.0 * ap_n - 1.0) * S / max(ap_ms / 1e3, 1e-9) / 1e9,
                      "sweep_physical_gbs": 2.0 * po_n * S / max(po_ms / 1e3, 1e-9) / 1e9,
                      "sigma_frac_of_hbm_peak": (3.0 * ap_n - 1.0) * S / max(ap_ms / 1e3, 1e-9) / 1e9 / peak_p,
                      "sweep_frac_of_hbm_peak": 2.0 * po_n * S / max(po_ms / 1e3, 1e-9) / 1e9 / peak_p,
                      "algorithmic_gbs": (2.0 * pool_groups * S + (ps.n_groups + 1) * S) / max(sweep_s, 1e-9) / 1e9,
                      "max_abs_gradient": float(np.max(np.abs(grads))), "gradient_norm": float(np.sqrt(np.sum(grads ** 2)))}
        if not args.no_cpu:
            psi_host = eng.get_state()
            pool_sweep["cpu_baseline"] = cpu_pool_sample(n, psi_host, hp, pool)
            del psi_host
    # ---- e2e through the reference-facing API ------------------------------------------------
    from openvqe_b200.ucc_family.get_energy_ucc import EnergyUCC
    from openvqe_b200 import engine as engine_mod
    engine_mod._ENGINES[(n, local_rank)] = eng  # the API uses the process-wide engine of this device
    ham_obj, gen_objs = build_host_objects(w)
    api = EnergyUCC()
    import openvqe_b200._hotpath as hot
    hot_energy = lambda th: hot.ucc_energy(th, ham_obj, gen_objs, hf, device=local_rank)
    hot_energy(ths[0])  # first call lowers + uploads H (cached afterwards, like the reference's one-time set-up)
    e_check = hot_energy(ths[args.warmup])
    assert abs(e_check - energies[0]) < 1e-9, (e_check, energies[0])
    eng.transfer_bytes(reset=True)
    barrier()
    t0 = time.perf_counter()
    eng.timer_begin()
    log = []
    for th in ths[args.warmup:]:
        if local_rank == 0 and world == 1:
            api.ucc_action(th, ham_obj, gen_objs, hf, log)
        else:
            hot_energy(th)
    ms_e2e = eng.timer_end()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    barrier()
    h2d, d2h = eng.transfer_bytes()
    ms_e2e = max_over_ranks(max(ms_e2e, wall_e2e))
    e2e_value = world * args.steps / (ms_e2e / 1e3)
    # ---- BASELINE config 4 names QUCCSD: the same excitations through EnergyUCC.action_quccsd ------------------
    # (gate-defined ansatz of reference get_energy_qucc.py:11-56; every excitation template is applied as one
    # tabulated plane rotation).  One gate-by-gate evaluation is timed beside it for comparison.
    quccsd = None
    if rank == 0 and world == 1 and not args.no_pool:
        from openvqe_b200.ucc_family.get_energy_qucc import EnergyUCC as EnergyQUCC

        class _Exc:
            def __init__(self, qbits):
                self.nbqbits, self.terms = n, [type("T", (), {"qbits": qbits})()]

        exc = []
        for o in range(int(owner.max()) + 1):
            xm = int(rot.x[np.argmax(owner == o)])
            qs = sorted(n - 1 - b for b in range(n) if (xm >> b) & 1)
            exc.append(_Exc([qs[2], qs[3], qs[0], qs[1]] if len(qs) == 4 else [qs[1], qs[0]]))
        qapi = EnergyQUCC()
        qth = ths[args.warmup]
        qapi.action_quccsd(qth, ham_obj, exc, hf, [])  # warm-up
        eng.synchronize()
        t0 = time.perf_counter()
        reps = max(1, args.steps)
        for k in range(reps):
            e_q = qapi.action_quccsd(ths[args.warmup + k % args.steps], ham_obj, exc, hf, [])
        t_tab = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        hot.prepare_quccsd_state(eng, n, hf, exc, qth, use_tables=False)
        e_gate = float(eng.expectation(eng.paulisum(ham_obj)).real)
        t_gate = time.perf_counter() - t0
        e_tab = qapi.action_quccsd(qth, ham_obj, exc, hf, [])
        quccsd = {"api": "openvqe_b200.ucc_family.get_energy_qucc.EnergyUCC.action_quccsd", "excitations": len(exc),
                  "evals_per_s": 1.0 / t_tab, "ms_per_eval": t_tab * 1e3, "gate_by_gate_ms_per_eval": t_gate * 1e3,
                  "abs_diff_vs_gate_by_gate": abs(e_tab - e_gate), "energy": e_tab}
    # ---- pool sweep at the reference's own size and the qubit sweep (rank 0 of a 1-GPU run only) ----------------
    sweep12, qsweep = None, None
    if rank == 0 and world == 1 and not args.no_pool:
        engine_mod._ENGINES.pop((n, local_rank), None)
        sweep12 = pool_sweep_12q()
        if not args.no_sweep:
            del eng, ps
            engine_mod.release_engines()
            qsweep = qubit_sweep(local_rank, [12, 16, 20, 24, 28, 30], with_cpu=not args.no_cpu)
    # ---- N > 1: the sharded C5 program on the same ranks (BASELINE config 5), reported under `sharded_c5` --------
    sharded = None
    if world > 1 and not args.no_sharded:
        try:
            del eng, ps
        except NameError:
            pass
        engine_mod.release_engines()
        import gc
        gc.collect()
        sharded = run_c5(args, rank, world, local_rank, dist=dist, as_dict=True)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel class -------------------------------------------------------
    peak, peak_src = peaks()
    n_rot, n_groups = len(rot.x), ps_groups(w)
    traffic, traffic_warm = {}, {}
    if os.path.exists(NCU_TRAFFIC):
        with open(NCU_TRAFFIC) as f:
            traffic = json.load(f)
    if os.path.exists(NCU_TRAFFIC_IN_STREAM):
        with open(NCU_TRAFFIC_IN_STREAM) as f:
            traffic_warm = json.load(f)

    def kernel_entry(name, label, ms_k, n_k, phys_per_launch, alg_total, fusion_key, fusion_units):
        ent = {"kernel": name, "what": label, "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
               "launches_per_step": n_k / args.steps, "ms_per_step": ms_k / args.steps,
               "avg_launch_us": ms_k * 1e3 / max(n_k, 1), "physical_bytes_per_launch": phys_per_launch,
               "achieved": phys_per_launch * n_k / max(ms_k / 1e3, 1e-9) / 1e9,
               "algorithmic_bytes_per_launch": alg_total / max(n_k, 1),
               "algorithmic_gbs": alg_total / max(ms_k / 1e3, 1e-9) / 1e9, fusion_key: fusion_units / max(n_k, 1),
               "traffic": None, "share_of_step": ms_k / max(ms_prof, 1e-9)}
        ent["frac"] = ent["achieved"] / peak
        tr = traffic.get(name)
        if tr:
            ent["traffic"] = tr["dram_read_bytes"] + tr["dram_write_bytes"]
            ent["traffic_source"] = ("ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, caches flushed before every kernel "
                                     "(the tail of a pass's write-back drains after the kernel ends), " + os.path.relpath(NCU_TRAFFIC, ROOT))
        tw = traffic_warm.get(name)
        if tw:  # what the kernel pulls from / pushes to DRAM inside the stream: consecutive passes share lines through L2
            ent["traffic_in_stream"] = tw["dram_read_bytes"] + tw["dram_write_bytes"]
            ent["traffic_in_stream_source"] = "the same with ncu --cache-control none, " + os.path.relpath(NCU_TRAFFIC_IN_STREAM, ROOT)
        return ent

    # kernel names of the default path: the real layout runs the item-table rotation kernel and the pair-mode expectation kernel
    rot_kernel = "k_col_stab" if real_layout and os.environ.get("VQE_COL_TAB", "2") == "2" else "k_tile_col"
    exp_kernel = "k_expect_rlp" if real_layout and os.environ.get("VQE_EXP_RL2", "1") != "0" else "k_expect_lean"
    prep = kernel_entry(rot_kernel, "rotation passes (every consecutive Pauli rotation whose X-mask fits the tile bits, collapsed runs)",
                        prep_ms, prep_n, 2.0 * S_phys, n_rot * 2.0 * S * args.steps, "rotations_per_pass", n_rot * args.steps)
    expk = kernel_entry(exp_kernel, "expectation passes (every X-mask group whose X-mask fits the tile bits)",
                        exp_ms, exp_n, S_phys, n_groups * S * args.steps, "groups_per_pass", n_groups * args.steps)
    for ent in (prep, expk):
        ent["state_layout"] = "real (2^n doubles: the state of a UCC evaluation is purely real)" if real_layout else "interleaved complex128"
    # what keeps the fraction below 1: a pass does r rotations / g groups of arithmetic on the tile while it sits in shared memory
    prep["limited_by"] = ("HBM for passes with few runs (7-9 JW doubles: ~0.8 of the peak), shared-memory wavefronts for passes with 25-28 "
                          "runs (2 loads + 2 stores of 8 bytes per rotated pair)")
    expk["limited_by"] = ("shared-memory bandwidth and issue: 16 bytes of tile reads per visited pair, ~230 pair visits per amplitude and "
                          "evaluation against 8 bytes of HBM traffic per amplitude and pass")
    dom, other = (prep, expk) if prep_ms >= exp_ms else (expk, prep)
    line = {"metric": "ucc_energy_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 (complex128 state)", "data": "synthetic",
            "config": config_for(w, world), "energy_first_step": energies[0],
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps, "ms_per_step": ms_e2e / args.steps,
                    "api": "openvqe_b200.ucc_family.get_energy_ucc.EnergyUCC.ucc_action"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": dom, "roofline_other": other,
            "profiled_ms_per_step": ms_prof / args.steps,
            "adapt_pool_sweep": pool_sweep, "adapt_pool_sweep_12q": sweep12, "qubit_sweep": qsweep, "quccsd": quccsd}
    if sharded is not None:
        line["sharded_c5"] = sharded
    if world == 1 and not args.no_pool and w["molecule"] != "h12":
        line["h12_standin"] = quick_value("h12", local_rank, args.steps, args.warmup)
    if world == 1 and not args.no_cpu:
        cb = cpu_sample(w)
        line["cpu_baseline"] = {"value": 1.0 / cb["seconds_per_eval"], "unit": "evals/s", "cores": cb["cores"],
                                "kind": "port", "sample": cb["sample"] + "; CPU port = oracle/c/vqe_oracle.c (OpenMP, thread count set explicitly)"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def ps_groups(w):
    return len(set(int(x) for x in w["ham_x"]))


# ------------------------------------------------------------------------------------------------
# C5: synthetic 30-36 qubit UCC rotations + expectation; sharded over the ranks when world > 1
# ------------------------------------------------------------------------------------------------
C5_DEFAULT_QUBITS = {1: 33, 2: 34, 4: 35, 8: 36}   # 2^33 amplitudes = 137 GB per GPU at every size (weak scaling)


def run_c5(args, rank, world, local_rank, dist=None, as_dict=False):
    import torch
    from tools import c5_synthetic as c5
    from openvqe_b200.engine import Engine
    from openvqe_b200.lowering import PackedTerms
    from openvqe_b200 import sharded
    torch.cuda.set_device(local_rank)
    own_group = False
    if world > 1 and dist is None:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        own_group = True
    steps, warmup = (1, 1) if as_dict else (args.steps, args.warmup)
    grad_components = 1 if as_dict else args.grad_components
    g = sharded.n_global_for(world)

    def barrier(eng):
        eng.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- sharded == unsharded at 30 qubits (driver-visible parity of the multi-GPU path) ------------------------
    verify = None
    if (as_dict or args.verify) and world > 1:
        nv = 30
        gen, ham = c5.generators(nv), c5.hamiltonian(nv)
        ang = gen["theta"][gen["owner"]] * gen["coeff"]
        hpv = PackedTerms(nv, ham["x"], ham["z"], ham["ny"], ham["cre"], np.zeros_like(ham["cre"]))
        se = sharded.ShardedEngine(nv, local_rank)
        se.set_basis_state(c5.hf_index(nv))
        se.apply_rotations(gen["x"], gen["z"], gen["ny"], ang)
        e_sh = se.expectation(se.paulisum(hpv)).real
        barrier(se)
        del se
        e_one = None
        if rank == 0:
            one = Engine(nv, device=local_rank)
            one.set_basis_state(c5.hf_index(nv))
            one.apply_rotations(gen["x"], gen["z"], gen["ny"], ang)
            e_one = one.expectation(one.paulisum(hpv)).real
            del one
            verify = {"qubits": nv, "energy_sharded": e_sh, "energy_one_gpu": e_one, "abs_err": abs(e_sh - e_one)}
        if dist is not None:
            dist.barrier()
    n = (0 if as_dict else args.qubits) or C5_DEFAULT_QUBITS.get(world, 33)
    nl = n - g
    S_local = 16.0 * (1 << nl)
    eng = sharded.ShardedEngine(n, local_rank) if world > 1 else Engine(n, device=local_rank)
    gen, ham = c5.generators(n), c5.hamiltonian(n)
    ps = eng.paulisum(PackedTerms(n, ham["x"], ham["z"], ham["ny"], ham["cre"], np.zeros_like(ham["cre"])))
    hf = c5.hf_index(n)
    owner, coeff = gen["owner"], gen["coeff"]
    plan = sharded.plan_rotations(n, g, gen["x"], gen["z"], gen["ny"], gen["theta"][owner] * coeff)
    n_local_pass = sum(1 for p in plan if p[0] == 0)
    n_peer_pass = len(plan) - n_local_pass
    n_gather_pass = sum(1 for p in plan if p[0] == 2)

    layout = {"real": False}

    def energy(theta):
        eng.set_basis_state(hf)
        eng.apply_rotations(gen["x"], gen["z"], gen["ny"], theta[owner] * coeff)
        layout["real"] = eng.real_layout   # a purely real state is kept as 2^nl doubles through the rotation passes
        return eng.expectation(ps).real

    fd_h = 1.4901161193847656e-08  # scipy's 2-point step
    # gradient components: generators that excite occupied -> virtual orbitals of the reference determinant first
    # (the others have an exactly vanishing first-order effect on |HF>)
    occ_mask = ((1 << (n // 2)) - 1) << (n - n // 2)
    first = {}
    for r_, o_ in enumerate(owner.tolist()):
        first.setdefault(o_, int(gen["x"][r_]))
    acting = [j for j, xm in sorted(first.items()) if 2 * bin(xm & occ_mask).count("1") == bin(xm).count("1")]
    grad_idx = (acting + [j for j in sorted(first) if j not in acting])[:grad_components]

    def step(theta):
        e0 = energy(theta)
        grad = []
        for j in grad_idx:
            tj = theta.copy()
            tj[j] += fd_h
            grad.append((energy(tj) - e0) / fd_h)
        return e0, grad

    ths = [gen["theta"] * (1.0 + 0.01 * s) for s in range(warmup + steps)]
    for th in ths[:warmup]:
        step(th)
    norm = eng.norm2()
    for k in range(6):
        eng.profile_read(k, reset=True)
    eng.profile(True)
    eng.gather_bytes(reset=True)
    eng.relabel_stats(reset=True)
    sampler = ClockSampler(local_rank) if rank == 0 and not as_dict else None
    barrier(eng)
    l0 = eng.launch_count
    eng.timer_begin()
    t0 = time.perf_counter()
    results = [step(th) for th in ths[warmup:]]
    ms = eng.timer_end()
    wall = (time.perf_counter() - t0) * 1e3
    barrier(eng)
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count - l0
    ms = max_over_ranks(max(ms, wall))
    prof = [eng.profile_read(k) for k in range(6)]
    eng.profile(False)
    prof = [(max_over_ranks(m), c) for m, c in prof]
    gather_bytes = eng.gather_bytes()
    n_swaps, swap_bytes = eng.relabel_stats()
    # ---- e2e: the same evaluations through the reference-facing call EnergyUCC.ucc_action with host objects ------
    # (every rank makes the same call; the sharded engine is the process-wide engine of this register size)
    e2e = None
    if not as_dict:
        from openvqe_b200 import engine as engine_mod
        from openvqe_b200.ucc_family.get_energy_ucc import EnergyUCC

        class _Term:
            __slots__ = ("coeff", "op", "qbits")

            def __init__(self, c, op, qb):
                self.coeff, self.op, self.qbits = c, op, qb

        class _Ham:
            def __init__(self, terms, const=0.0):
                self.nbqbits, self.terms, self.constant_coeff = n, terms, const

        hterms, hconst = [], 0.0
        for cf, op, qb in c5.to_terms(n, ham, "cre"):
            if op:
                hterms.append(_Term(cf, op, qb))
            else:
                hconst += cf
        ham_obj = _Ham(hterms, hconst)
        gens = [[] for _ in range(gen["n_generators"])]
        for (cf, op, qb), o in zip(c5.to_terms(n, gen, "coeff"), owner.tolist()):
            gens[o].append(_Term(cf, op, qb))
        gen_objs = [_Ham(t) for t in gens]
        os.environ["VQE_B200_DEVICE"] = str(local_rank)
        engine_mod._ENGINES[(n, engine_mod.default_device())] = eng
        api = EnergyUCC()
        e_api = api.ucc_action(ths[warmup], ham_obj, gen_objs, hf, [])  # lowers + uploads H once (cached afterwards)
        assert abs(e_api - results[0][0]) < 1e-9, (e_api, results[0][0])
        eng.transfer_bytes(reset=True)
        barrier(eng)
        t0 = time.perf_counter()
        for th in ths[warmup:]:
            api.ucc_action(th, ham_obj, gen_objs, hf, [])
        eng.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        h2d, d2h = eng.transfer_bytes()
        barrier(eng)
        e2e = {"value": steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": h2d / steps,
               "d2h_bytes_per_step": d2h / steps, "ms_per_step": e2e_s * 1e3 / steps,
               "api": "openvqe_b200.ucc_family.get_energy_ucc.EnergyUCC.ucc_action (sharded engine)"}
        engine_mod._ENGINES.pop((n, engine_mod.default_device()), None)
    del eng, ps
    if rank != 0:
        if own_group:
            dist.destroy_process_group()
        return None
    peak, peak_src = peaks()
    evals = steps * (1 + grad_components)
    n_rot = int(np.count_nonzero(gen["theta"][owner] * coeff))
    # local rotation pass: 2*S_local of HBM per rank
    loc_ms, loc_n = prof[0]
    peer_ms, peer_n = prof[4]
    exl_ms, exl_n = prof[1]
    exp_ms, exp_n = prof[5]
    roofline = {"kernel": "rotation passes (local)", "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                "launches_per_eval": loc_n / evals, "avg_launch_ms": loc_ms / max(loc_n, 1),
                "physical_bytes_per_launch": 2.0 * S_local * (0.5 if layout["real"] else 1.0),
                "state_layout": "real (2^nl doubles per rank during the rotation passes)" if layout["real"] else "interleaved complex128",
                "achieved": 2.0 * S_local * (0.5 if layout["real"] else 1.0) * loc_n / max(loc_ms / 1e3, 1e-9) / 1e9,
                "algorithmic_gbs": n_rot * 2.0 * S_local * evals / max((loc_ms + peer_ms) / 1e3, 1e-9) / 1e9,
                "traffic": None, "share_of_step": (loc_ms + peer_ms) / ms}
    roofline["frac"] = roofline["achieved"] / peak
    nvlink = None
    if world > 1:
        # Exchange-form peer pass: per rank S_local/2 in + S_local/2 out for the loads and the same again for the stores
        # = S_local per direction per pass.  Gather-form peer pass: only the partner amplitudes the planner's dependency
        # closure names are read (inbound only, nothing is written remotely): bytes counted by the library per launch of
        # k_gather_need.  The figure below is inbound bytes per rank over the time of ALL peer-pass launches (gather
        # kernels and the local passes that follow them included), i.e. a lower bound of the link rate while busy.
        # With qubit relabelling (default) a rotation never runs as a peer pass: a qubit the program flips is moved out of
        # its global index bit by ONE exchange of half a shard (each rank reads a quarter of a shard from its partner and
        # writes a quarter into it); the planner's peer-pass counts then describe the program WITHOUT relabelling.
        relabelled = n_swaps > 0 or os.environ.get("VQE_RELABEL", "1") != "0"
        planned = {"local": n_local_pass, "peer": n_peer_pass, "peer_in_gather_form": n_gather_pass}
        if relabelled:
            n_peer_pass = n_gather_pass = 0
            n_local_pass = loc_n / evals
        n_exch = n_peer_pass - n_gather_pass
        gather_in = float(gather_bytes) if gather_bytes is not None else None
        inbound = (gather_in if gather_in is not None else 0.0) + S_local * n_exch * evals + float(swap_bytes)
        nvlink = {"qubit_swaps_per_eval": n_swaps / evals, "swap_bytes_read_per_eval_per_rank": swap_bytes / evals,
                  "rotation_passes_without_relabelling": planned,
                  "peer_passes_per_eval": n_peer_pass, "gather_form_passes_per_eval": n_gather_pass,
                  "exchange_form_passes_per_eval": n_exch, "peer_launches_per_eval": peer_n / evals,
                  "peer_ms_per_eval": peer_ms / evals,
                  "gathered_bytes_per_eval_per_rank": gather_in / evals if gather_in is not None else None,
                  "inbound_bytes_per_eval_per_rank": inbound / evals,
                  "achieved_gbs_inbound": inbound / max(peer_ms / 1e3, 1e-9) / 1e9,
                  "peak_gbs_per_direction": 770.0, "peak_source": "measured peer copy per direction (B200_PROFILING.md); 900 nominal",
                  "expectation_peer_pass": {"launches_per_eval": exp_n / evals, "ms_per_launch": exp_ms / max(exp_n, 1),
                                            "achieved_gbs_in": 0.5 * S_local * exp_n / max(exp_ms / 1e3, 1e-9) / 1e9}}
        nvlink["frac"] = nvlink["achieved_gbs_inbound"] / 770.0
    cfg = {"workload": "C5 synthetic %d-qubit UCC energy%s: %d Pauli rotations (256 generators) + <H> over %d terms / "
                       "%d X-mask groups, state sharded over %d GPU(s) (top %d qubits global)"
                       % (n, " + %d FD gradient component(s)" % grad_components if grad_components else "",
                          n_rot, len(ham["x"]), ham["n_groups"], world, g),
           "qubits": n, "shard_bytes": S_local, "l2_policy": "shard (%.0f GB) larger than L2" % (S_local / 1e9),
           "parallelism": ("state sharded; qubit swaps (half a shard over NVLink each) keep the rotation passes local, "
                           "expectation peer passes over NVLink") if world > 1 else "1 GPU",
           "rotation_passes": {"local": n_local_pass, "peer": n_peer_pass, "peer_in_gather_form": n_gather_pass},
           "expectation_passes": {"local": exl_n / evals, "peer": exp_n / evals}}
    body = {"value": evals / (ms / 1e3), "unit": "evals/s", "seconds_per_energy_evaluation": ms / 1e3 / evals,
            "seconds_per_step": ms / 1e3 / steps, "steps": steps, "warmup": warmup,
            "energy_first_step": results[0][0], "gradient_first_step": results[0][1], "gradient_components": grad_idx,
            "norm2_after_warmup": norm, "verify_vs_unsharded": verify, "gpu_launches": int(launches),
            "roofline": roofline, "nvlink": nvlink,
            "expectation": {"local_ms_per_eval": exl_ms / evals, "peer_ms_per_eval": exp_ms / evals}}
    if as_dict:
        body["config"] = cfg
        return body
    line = {"metric": "ucc_energy_evals_per_s", "n_gpus": world, "ms_per_step": ms / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128 state)", "data": "synthetic", "config": cfg,
            "e2e": e2e, "clocks": clocks}
    line.update(body)
    print(json.dumps(line), flush=True)
    if own_group:
        dist.destroy_process_group()
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-pool", action="store_true", help="skip the ADAPT pool-gradient sweep / QUCCSD / qubit-sweep legs")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 12...30-qubit sweep")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the sharded C5 leg")
    ap.add_argument("--molecule", default=None, choices=sorted(MOLECULES),
                    help="24-qubit instance of the c4 workload (default: VQE_BENCH_MOLECULE or h2o, the molecule BASELINE names)")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4: 24-qubit UCCSD energy (headline; replicas when N > 1, plus the sharded C5 leg).  c5: only the "
                         "synthetic 30-36 qubit state SHARDED over the N GPUs")
    ap.add_argument("--qubits", type=int, default=0, help="c5 only: register size (default 33/34/35/36 for 1/2/4/8 GPUs)")
    ap.add_argument("--grad-components", type=int, default=0, help="c5 only: forward-difference gradient components per step")
    ap.add_argument("--verify", action="store_true", help="c5 only: first compare sharded with unsharded at 30 qubits")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.workload == "c5":
        run_c5(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
