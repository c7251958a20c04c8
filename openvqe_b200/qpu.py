"""The engine as a generic myQLM QPU.

``B200QPU().submit(circuit.to_job(observable=H))`` evaluates any gate circuit the reference builds with
``qat.lang.AQASM`` (X, Y, Z, H, S, T, PH, RX, RY, RZ, CNOT) on the CUDA engine, so that code written against
``qat.qpus.get_default_qpu()`` -- e.g. the weighted subspace-search driver, reference
openvqe/common_files/get_energy_WSSVQE.py:151-178, or fun_fidelity, adapt/fermionic_adapt_vqe.py:331-361 -- gets the
B200 back-end by swapping the QPU object.  Duck-typed like the rest of the boundary:

  job.circuit     ``.nbqbits`` and either ``iterate_simple()`` -> (name, params, qbits) (myQLM's public flattening) or
                  ``.ops`` whose items carry ``.name`` / ``.angle`` / ``.qbits``
  job.observable  ``.terms`` (Term: ``.coeff``, ``.op``, ``.qbits``), ``.constant_coeff``, ``.nbqbits``  -> OBS job:
                  ``Result.value`` = <psi|O|psi> (exact, no shots)
  no observable   SAMPLE job: the Result iterates over Samples with ``.state.int``, ``.amplitude``, ``.probability``
                  (every non-zero amplitude, the form get_statevector consumes; n <= 30)

When myQLM is installed the class derives from ``qat.core.qpu.QPUHandler`` (``submit_job`` is the plugin hook and
``submit`` / batches / plugin stacking come from myQLM); without it ``submit`` is provided here.  Gates outside the
engine's native set are rewritten as Z rotations plus a tracked global phase (the phase is applied to the state
before sampling; an observable does not see it).
"""
from __future__ import annotations

import cmath
import math

import numpy as np

from .engine import BUF_PSI, GATE_KINDS, get_engine

try:  # pragma: no cover - myQLM is not installable in the build container
    from qat.core.qpu import QPUHandler as _Base
except Exception:  # noqa: BLE001
    _Base = object

# name -> (native gate, angle, global phase) for the fixed gates rewritten onto the native set
_REWRITE = {
    "Z": ("RZ", math.pi, cmath.exp(0.5j * math.pi)),
    "S": ("RZ", 0.5 * math.pi, cmath.exp(0.25j * math.pi)),
    "T": ("RZ", 0.25 * math.pi, cmath.exp(0.125j * math.pi)),
    "Y": ("RY", math.pi, 1j),          # RY(pi) = -i Y
}


class _State:
    __slots__ = ("int",)

    def __init__(self, i):
        self.int = int(i)


class Sample:
    __slots__ = ("state", "amplitude", "probability")

    def __init__(self, i, amp):
        self.state = _State(i)
        self.amplitude = complex(amp)
        self.probability = abs(amp) ** 2


class Result:
    def __init__(self, value=None, samples=None):
        self.value = value
        self.raw_data = samples if samples is not None else []

    def __iter__(self):
        return iter(self.raw_data)

    def __len__(self):
        return len(self.raw_data)


def circuit_gates(circuit):
    """[(name, [qbits], angle or None)] of a flattened circuit, plus the global phase factor the rewriting introduced."""
    if hasattr(circuit, "iterate_simple"):
        items = [(name, list(params or []), list(qbits)) for name, params, qbits in circuit.iterate_simple()]
    else:
        items = [(getattr(op, "name", None) or op.gate, [] if getattr(op, "angle", None) is None else [op.angle], list(op.qbits))
                 for op in circuit.ops]
    gates, phase = [], 1.0 + 0.0j
    for name, params, qbits in items:
        name = str(name)
        if name == "I":
            continue
        if name in ("X", "H"):
            gates.append((name, qbits, None))
        elif name in ("RX", "RY", "RZ"):
            gates.append((name, qbits, float(params[0])))
        elif name == "PH":
            gates.append(("RZ", qbits, float(params[0])))
            phase *= cmath.exp(0.5j * float(params[0]))
        elif name in _REWRITE:
            native, angle, ph = _REWRITE[name]
            gates.append((native, qbits, angle))
            phase *= ph
        elif name in ("CNOT", "C-X"):
            gates.append(("CNOT", qbits, None))
        else:
            raise NotImplementedError("B200QPU: gate %r is not in the supported set (X Y Z H S T PH RX RY RZ CNOT)" % name)
    return gates, phase


class B200QPU(_Base):
    """Drop-in for ``qat.qpus.get_default_qpu()`` backed by the CUDA engine."""

    def __init__(self, device=None, zero_tol: float = 0.0):
        if _Base is not object:
            super().__init__()
        self.device = device
        self.zero_tol = float(zero_tol)

    def _prepare(self, circuit):
        engine = get_engine(circuit.nbqbits, self.device)
        gates, phase = circuit_gates(circuit)
        engine.set_basis_state(0)
        if gates:
            kinds = [GATE_KINDS[g[0]] for g in gates]
            q0 = [g[1][0] for g in gates]
            q1 = [g[1][1] if len(g[1]) > 1 else 0 for g in gates]
            ang = [0.0 if g[2] is None else g[2] for g in gates]
            engine.apply_gates(kinds, q0, q1, ang)
        return engine, phase

    def submit_job(self, job):
        circuit = job.circuit
        observable = getattr(job, "observable", None)
        engine, phase = self._prepare(circuit)
        if observable is not None:
            return Result(value=float(engine.expectation(engine.paulisum(observable)).real))
        if circuit.nbqbits > 30:
            raise ValueError("SAMPLE job on %d qubits: the amplitude list does not fit a host array" % circuit.nbqbits)
        if phase != 1.0:
            engine.scale_state(phase)
        psi = engine.get_state(BUF_PSI)
        keep = np.nonzero(np.abs(psi) > self.zero_tol)[0]
        return Result(samples=[Sample(int(i), psi[i]) for i in keep])

    if _Base is object:
        def submit(self, job):
            if hasattr(job, "jobs"):  # a Batch
                return [self.submit_job(j) for j in job.jobs]
            return self.submit_job(job)


def get_b200_qpu(device=None):
    """Counterpart of ``qat.qpus.get_default_qpu``."""
    return B200QPU(device)
