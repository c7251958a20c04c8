"""Time the opt-in adjoint gradient on the 24-qubit bench workload (1 818 generators): python tools/time_adjoint.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from openvqe_b200 import _hotpath  # noqa: E402

w = bench.load_workload()
ham, gens = bench.build_host_objects(w)
theta = bench.thetas_for(w, 1, 0)[0]
hf = w["hf_init_sp"]
e0 = _hotpath.ucc_energy(theta, ham, gens, hf)
t0 = time.perf_counter()
e, g = _hotpath.ucc_energy_and_gradient(theta, ham, gens, hf)
t_adj = time.perf_counter() - t0
t0 = time.perf_counter()
for _ in range(3):
    _hotpath.ucc_energy(theta, ham, gens, hf)
t_e = (time.perf_counter() - t0) / 3
j = int(np.argmax(np.abs(g)))
h = 1e-5
tp, tm = theta.copy(), theta.copy()
tp[j] += h
tm[j] -= h
fd = (_hotpath.ucc_energy(tp, ham, gens, hf) - _hotpath.ucc_energy(tm, ham, gens, hf)) / (2 * h)
print(json.dumps({"qubits": w["n"], "generators": len(gens), "adjoint_gradient_s": t_adj, "energy_eval_s": t_e,
                  "finite_difference_gradient_s_estimate": (len(gens) + 1) * t_e, "abs_energy_diff": abs(e - e0),
                  "largest_component": float(g[j]), "its_central_difference": fd, "grad_norm": float(np.linalg.norm(g))}))
