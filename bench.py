#!/usr/bin/env python
"""bench.py -- UCC energy evaluations per second on the 24-qubit (C4-scale) workload.

A "step" is ONE energy evaluation E(theta) of the Trotterised UCCSD ansatz: |HF> -> 14 112 Pauli rotations
(1 818 generators) -> <H> over 14 905 Pauli terms in 2 767 X-mask groups, on a 2^24 complex128 state (268 MB,
larger than the 126 MB L2, so nothing is L2-resident between sweeps).  theta changes every step.

  value   evaluations/s with the Hamiltonian and rotation program resident in HBM (only the angles and the
          16-byte result cross PCIe), timed with CUDA events on the engine's stream.
  e2e     the same metric through the reference-facing call EnergyUCC.ucc_action(theta, H, generators, hf)
          with host objects: lowering (cached), H2D of the operation descriptors, D2H of the energy, every step.
  N > 1   below 33 qubits the path shards as independent energy evaluations: every rank
          evaluates its own theta on its own GPU, no data-path collective ("weak" scaling).

The line also carries `adapt_pool_sweep` (sigma = H psi plus <sigma|A_k|psi> for the 1 818-operator pool of the
workload, the second half of BASELINE's metric) and `quccsd` (the same excitations through the gate-defined
EnergyUCC.action_quccsd, BASELINE config 4), both outside the timed steps.

`--workload c5 [--qubits n] [--grad-components k] [--verify]`: the synthetic C5 program (tools/c5_synthetic.py) on a
state SHARDED over the N ranks (33/34/35/36 qubits for 1/2/4/8 GPUs = 137 GB per GPU): local-pass HBM GB/s, peer-pass
NVLink figures, pass counts by form, e2e through EnergyUCC.ucc_action on the sharded engine.

`--impl reference` times the CPU port of the reference path (oracle/c, OpenMP over all host cores) on a bounded
sample of the same workload and scales it to one evaluation.
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

WORKLOAD = "C4-scale 24-qubit UCCSD energy evaluation (H12/STO-3G stand-in for H2O/6-31G active space): " \
           "14112 Pauli rotations + <H> over 14905 terms / 2767 X-mask groups"


def load_workload():
    z = np.load(os.path.join(ROOT, "tests", "golden", "h12_sto3g_24q.npz"))
    w = {k: z[k] for k in z.files}
    w["n"] = int(w["n"])
    w["hf_init_sp"] = int(w["hf_init_sp"])
    w["meta"] = json.loads(str(w["meta"]))
    return w


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


# ------------------------------------------------------------------------------------------------
def cpu_sample(w, n_rot_sample=48, n_term_sample=96):
    """Time the CPU port on a bounded sample and scale to one full evaluation."""
    from oracle import c_oracle
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
    return {"seconds_per_eval": est, "t_rotation_s": t_rot, "t_term_s": t_term, "cores": c_oracle.threads(),
            "sample": "%d of %d rotations + %d of %d Hamiltonian terms at 24 qubits (one 2^24 sweep each), scaled "
                      "linearly to one evaluation" % (n_rot_sample, len(angles), n_term_sample, len(ham.x))}


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = load_workload()
    for _ in range(min(args.warmup, 1)):
        cpu_sample(w, 8, 8)
    ests, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_sample(w, 24, 48)
        ests.append(last["seconds_per_eval"])
    sec = float(np.mean(ests))
    val = 1.0 / sec
    line = {"impl": "reference", "metric": "ucc_energy_evals_per_s", "value": val, "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128 state)", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": last["cores"], "kind": "port",
                             "sample": "per step: " + last["sample"] + "; CPU port = oracle/c/vqe_oracle.c (OpenMP), the reference's own myQLM simulator is "
                               "not installable here", "wall_s": time.perf_counter() - t0},
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


def run_ours(args, rank, world, local_rank):
    from openvqe_b200.engine import Engine
    from openvqe_b200.lowering import PackedTerms
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w = load_workload()
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
    # ---- timed: resident inputs -------------------------------------------------------------
    for k in range(6):
        eng.profile_read(k, reset=True)
    eng.profile(True)
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
    prep_ms, prep_n = eng.profile_read(0)
    exp_ms, exp_n = eng.profile_read(1)
    eng.profile(False)
    value = world * args.steps / (ms / 1e3)
    # ---- second half of the BASELINE metric: ADAPT pool-gradient sweep time ------------------------
    # sigma = H psi once, then <sigma|A_k|psi> for every operator of the pool in one batched sweep (reference
    # fermionic_adapt_vqe.py:77-122).  Pool = the 1 818 UCCSD generators of the workload (as T - T^dagger);
    # psi = the state of the last evaluation.  Not part of the timed steps above.
    pool_sweep = None
    if rank == 0 and world == 1 and not args.no_pool:
        from openvqe_b200.engine import BUF_PSI, BUF_SIGMA
        n_gen = int(owner.max()) + 1
        offs = np.zeros(n_gen + 1, dtype=np.int32)
        np.add.at(offs, owner + 1, 1)
        offs = np.cumsum(offs).astype(np.int32)
        pool = PackedTerms(n, rot.x, rot.z, rot.ny, np.zeros_like(rc), rc, offs)  # i * (real coefficient) * P: anti-Hermitian
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
        pool_sweep = {"qubits": n, "pool_size": n_gen, "pool_strings": int(len(rot.x)), "seconds": sweep_s,
                      "sigma_ms": ap_ms, "sigma_passes": ap_n, "sweep_ms": po_ms, "sweep_passes": po_n,
                      "algorithmic_gbs": (2.0 * pool_groups * S + (ps.n_groups + 1) * S) / max(sweep_s, 1e-9) / 1e9,
                      "max_abs_gradient": float(np.max(np.abs(grads))), "gradient_norm": float(np.sqrt(np.sum(grads ** 2)))}
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
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel -------------------------------------------------------
    peak, peak_src = peaks()
    n_rot, n_groups = len(rot.x), ps.n_groups
    prep_alg = n_rot * 2.0 * S * args.steps      # 2*S per rotation
    exp_alg = n_groups * S * args.steps          # S per X-mask group
    prep = {"kernel": "k_tile_rot", "bound": "hbm", "achieved": prep_alg / (prep_ms / 1e3) / 1e9, "peak": peak,
            "unit": "GB/s", "launches_per_step": prep_n / args.steps, "ms_per_step": prep_ms / args.steps,
            "algorithmic_bytes_per_launch": prep_alg / max(prep_n, 1), "physical_bytes_per_launch": 2.0 * S,
            "physical_gbs": prep_n * 2.0 * S / (prep_ms / 1e3) / 1e9, "rotations_per_pass": n_rot * args.steps / max(prep_n, 1)}
    expk = {"kernel": "k_tile_expect", "bound": "hbm", "achieved": exp_alg / (exp_ms / 1e3) / 1e9, "peak": peak,
            "unit": "GB/s", "launches_per_step": exp_n / args.steps, "ms_per_step": exp_ms / args.steps,
            "algorithmic_bytes_per_launch": exp_alg / max(exp_n, 1), "physical_bytes_per_launch": S,
            "physical_gbs": exp_n * S / (exp_ms / 1e3) / 1e9, "groups_per_pass": n_groups * args.steps / max(exp_n, 1)}
    dom, other = (prep, expk) if prep_ms >= exp_ms else (expk, prep)
    roofline = dict(dom)
    roofline["frac"] = dom["achieved"] / peak
    # DRAM bytes per launch of that kernel from the committed ncu capture of this same command (profiles/)
    roofline["traffic"] = None
    tpath = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tr = json.load(f).get(dom["kernel"])
        if tr:
            roofline["traffic"] = tr["dram_read_bytes"] + tr["dram_write_bytes"]
            roofline["traffic_source"] = "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r1_ncu_traffic.json"
    roofline["peak_source"] = peak_src
    roofline["share_of_step"] = dom["ms_per_step"] / (ms / args.steps)
    other = dict(other)
    other["frac"] = other["achieved"] / peak
    line = {"metric": "ucc_energy_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 (complex128 state)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "qubits": n, "state_bytes": S, "l2_policy": "state (268 MB) larger than L2 (126 MB)",
                       "parallelism": "replicas: independent energy evaluations per GPU" if world > 1 else "1 GPU",
                       "energy_first_step": energies[0]},
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps, "ms_per_step": ms_e2e / args.steps,
                    "api": "openvqe_b200.ucc_family.get_energy_ucc.EnergyUCC.ucc_action"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_other": other,
            "adapt_pool_sweep": pool_sweep, "quccsd": quccsd}
    if world == 1 and not args.no_cpu:
        cb = cpu_sample(w)
        line["cpu_baseline"] = {"value": 1.0 / cb["seconds_per_eval"], "unit": "evals/s", "cores": cb["cores"],
                                "kind": "port", "sample": cb["sample"] + "; CPU port = oracle/c/vqe_oracle.c (OpenMP)"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# C5: synthetic 30-36 qubit UCC rotations + expectation; sharded over the ranks when world > 1
# ------------------------------------------------------------------------------------------------
C5_DEFAULT_QUBITS = {1: 33, 2: 34, 4: 35, 8: 36}   # 2^33 amplitudes = 137 GB per GPU at every size (weak scaling)


def run_c5(args, rank, world, local_rank):
    import torch
    from tools import c5_synthetic as c5
    from openvqe_b200.engine import Engine
    from openvqe_b200.lowering import PackedTerms
    from openvqe_b200 import sharded
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.qubits or C5_DEFAULT_QUBITS.get(world, 33)
    g = sharded.n_global_for(world)
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

    def energy(theta):
        eng.set_basis_state(hf)
        eng.apply_rotations(gen["x"], gen["z"], gen["ny"], theta[owner] * coeff)
        return eng.expectation(ps).real

    fd_h = 1.4901161193847656e-08  # scipy's 2-point step
    # gradient components: generators that excite occupied -> virtual orbitals of the reference determinant first
    # (the others have an exactly vanishing first-order effect on |HF>)
    occ_mask = ((1 << (n // 2)) - 1) << (n - n // 2)
    first = {}
    for r_, o_ in enumerate(owner.tolist()):
        first.setdefault(o_, int(gen["x"][r_]))
    acting = [j for j, xm in sorted(first.items()) if 2 * bin(xm & occ_mask).count("1") == bin(xm).count("1")]
    grad_idx = (acting + [j for j in sorted(first) if j not in acting])[:args.grad_components]

    def step(theta):
        e0 = energy(theta)
        grad = []
        for j in grad_idx:
            tj = theta.copy()
            tj[j] += fd_h
            grad.append((energy(tj) - e0) / fd_h)
        return e0, grad

    def barrier():
        eng.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ths = [gen["theta"] * (1.0 + 0.01 * s) for s in range(args.warmup + args.steps)]
    for th in ths[:args.warmup]:
        step(th)
    norm = eng.norm2()
    for k in range(6):
        eng.profile_read(k, reset=True)
    eng.profile(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    l0 = eng.launch_count
    eng.timer_begin()
    t0 = time.perf_counter()
    results = [step(th) for th in ths[args.warmup:]]
    ms = eng.timer_end()
    wall = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count - l0
    ms = max_over_ranks(max(ms, wall))
    prof = [eng.profile_read(k) for k in range(6)]
    eng.profile(False)
    prof = [(max_over_ranks(m), c) for m, c in prof]
    # ---- e2e: the same evaluations through the reference-facing call EnergyUCC.ucc_action with host objects ------
    # (every rank makes the same call; the sharded engine is the process-wide engine of this register size)
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
    e_api = api.ucc_action(ths[args.warmup], ham_obj, gen_objs, hf, [])  # lowers + uploads H once (cached afterwards)
    assert abs(e_api - results[0][0]) < 1e-9, (e_api, results[0][0])
    eng.transfer_bytes(reset=True)
    barrier()
    t0 = time.perf_counter()
    for th in ths[args.warmup:]:
        api.ucc_action(th, ham_obj, gen_objs, hf, [])
    eng.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d, d2h = eng.transfer_bytes()
    barrier()
    verify = None
    if args.verify and n <= 31:
        # the same evaluation on ONE unsharded context (rank 0 only; needs 2^n amplitudes next to the shard)
        if rank == 0:
            one = Engine(n, device=local_rank)
            ps1 = one.paulisum(PackedTerms(n, ham["x"], ham["z"], ham["ny"], ham["cre"], np.zeros_like(ham["cre"])))
            one.set_basis_state(hf)
            th = ths[args.warmup]
            one.apply_rotations(gen["x"], gen["z"], gen["ny"], th[owner] * coeff)
            verify = abs(one.expectation(ps1).real - results[0][0])
            del one
        barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    evals = args.steps * (1 + args.grad_components)
    n_rot = int(np.count_nonzero(gen["theta"][owner] * coeff))
    # local rotation pass: 2*S_local of HBM per rank; r rotations fused per pass are credited r*2*S
    loc_ms, loc_n = prof[0]
    peer_ms, peer_n = prof[4]
    exl_ms, exl_n = prof[1]
    exp_ms, exp_n = prof[5]
    rot_alg = n_rot * 2.0 * S_local * evals
    roofline = {"kernel": "k_tile_rot (local + peer passes)", "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                "achieved": rot_alg / max((loc_ms + peer_ms) / 1e3, 1e-9) / 1e9,
                "algorithmic_bytes_per_launch": rot_alg / max(loc_n + peer_n, 1),
                "local_pass": {"launches_per_eval": loc_n / evals, "ms_per_launch": loc_ms / max(loc_n, 1),
                               "physical_gbs": 2.0 * S_local * loc_n / max(loc_ms / 1e3, 1e-9) / 1e9},
                "traffic": None, "share_of_step": (loc_ms + peer_ms) / ms}
    roofline["frac"] = roofline["achieved"] / peak
    nvlink = None
    if world > 1:
        # a peer pass moves, per rank, S_local/2 in and S_local/2 out over NVLink for the loads and the same again
        # for the stores (the partner mirrors it): S_local per direction per pass
        nvlink = {"kernel": "k_tile_rot peer pass (tiles of ranks r and r^m staged through peer memory)",
                  "launches_per_eval": peer_n / evals, "ms_per_launch": peer_ms / max(peer_n, 1),
                  "bytes_per_direction_per_launch": S_local,
                  "achieved_gbs_per_direction": S_local * peer_n / max(peer_ms / 1e3, 1e-9) / 1e9,
                  "peak_gbs_per_direction": 900.0, "peak_source": "NVLink 5 nominal per direction per GPU",
                  "expectation_peer_pass": {"launches_per_eval": exp_n / evals, "ms_per_launch": exp_ms / max(exp_n, 1),
                                            "achieved_gbs_in": 0.5 * S_local * exp_n / max(exp_ms / 1e3, 1e-9) / 1e9}}
        nvlink["frac"] = nvlink["achieved_gbs_per_direction"] / 900.0
    line = {"metric": "ucc_energy_evals_per_s", "value": evals / (ms / 1e3), "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128 state)", "data": "synthetic",
            "config": {"workload": "C5 synthetic %d-qubit UCC energy%s: %d Pauli rotations (256 generators) + <H> over %d terms / "
                                   "%d X-mask groups, state sharded over %d GPU(s) (top %d qubits global)"
                                   % (n, " + %d FD gradient component(s)" % args.grad_components if args.grad_components else "",
                                      n_rot, len(ham["x"]), ham["n_groups"], world, g),
                       "qubits": n, "shard_bytes": S_local, "l2_policy": "shard (%.0f GB) larger than L2" % (S_local / 1e9),
                       "parallelism": "state sharded, peer passes over NVLink" if world > 1 else "1 GPU",
                       "rotation_passes": {"local": n_local_pass, "peer": n_peer_pass, "peer_in_gather_form": n_gather_pass},
                       "expectation_passes": {"local": exl_n / evals, "peer": exp_n / evals},
                       "energy_first_step": results[0][0], "gradient_first_step": results[0][1],
                       "gradient_components": grad_idx, "norm2_after_warmup": norm,
                       "verify_vs_unsharded_abs_err": verify},
            "e2e": {"value": args.steps / e2e_s, "unit": "evals/s", "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps, "ms_per_step": e2e_s * 1e3 / args.steps,
                    "api": "openvqe_b200.ucc_family.get_energy_ucc.EnergyUCC.ucc_action (sharded engine)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "nvlink": nvlink,
            "expectation": {"local_ms_per_eval": exl_ms / evals, "peer_ms_per_eval": exp_ms / evals}}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-pool", action="store_true", help="skip the ADAPT pool-gradient sweep timing")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4: 24-qubit UCCSD energy (headline; replicas when N > 1).  c5: synthetic 30-36 qubit state SHARDED over the N GPUs")
    ap.add_argument("--qubits", type=int, default=0, help="c5 only: register size (default 33/34/35/36 for 1/2/4/8 GPUs)")
    ap.add_argument("--grad-components", type=int, default=0, help="c5 only: forward-difference gradient components per step")
    ap.add_argument("--verify", action="store_true", help="c5 only, n <= 31: compare with one unsharded context on rank 0")
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
