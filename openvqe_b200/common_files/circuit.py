"""Gate-level description of the circuits the reference WOULD build.

Two jobs:
  * the QUCCSD excitation templates (reference openvqe/common_files/circuit.py:13-106)
    as flat gate tuples ``(name, qubits, angle)`` that ``Engine.apply_gates``
    executes on the GPU;
  * gate bookkeeping: the reference returns CNOT / H / "RX" / "RY" counts obtained
    by grepping ``str(op)`` for ``gate='NAME'`` (circuit.py:186-205).  ``CircuitSummary``
    reproduces those strings, including myQLM's ``_0, _1, ...`` dictionary keys for
    parametrised gates, so ``count`` gives the same
    numbers without building a myQLM circuit.
"""
from __future__ import annotations

from math import pi


class GateOp:
    __slots__ = ("gate", "qbits", "name", "angle")

    def __init__(self, gate, qbits, name, angle):
        self.gate, self.qbits, self.name, self.angle = gate, list(qbits), name, angle

    def __repr__(self):
        return "Op(gate='%s', qbits=%s)" % (self.gate, self.qbits)


class CircuitSummary:
    """What ``prepare_state_ansatz`` returns: ``.ops`` for ``count`` and the raw
    gate tuples for the engine."""

    def __init__(self, nbqbits, gates):
        self.nbqbits = nbqbits
        self.gates = list(gates)
        keys = {}
        self.ops = []
        for name, qbits, angle in self.gates:
            if angle is None:
                key = name
            else:
                ident = (name, float(angle))
                if ident not in keys:
                    keys[ident] = "_%d" % len(keys)
                key = keys[ident]
            self.ops.append(GateOp(key, qbits, name, angle))


def count(gate, mylist):
    """Number of ops whose string contains ``gate='<GATE>'`` (lower-case names
    are upper-cased), as reference circuit.py:186-205."""
    gate = str(gate)
    if gate == gate.lower():
        gate = gate.upper()
    needle = "gate='%s'" % gate
    return sum(1 for op in mylist if needle in str(op))


def hf_gates(nbqbits, hf_init_sp, padded):
    """X gates of the Hartree-Fock determinant.  ``padded=False`` reproduces the
    reference's ``binary_repr`` without width (get_energy_qucc.py:40-45): the bit
    string is read MSB-first from qubit 0 with no zero padding."""
    bits = format(int(hf_init_sp), "b")
    if padded:
        bits = bits.zfill(nbqbits)
    return [("X", [j], None) for j in range(min(nbqbits, len(bits))) if bits[j] == "1"]


def pauli_rotation_gates(op, qbits, angle):
    """Staircase circuit of exp(-i angle P) as myQLM's Trotterisation emits it."""
    act = [(l, q) for l, q in zip(op, qbits) if l != "I"]
    if not act:
        return []
    g = []
    for l, q in act:
        if l == "X":
            g.append(("H", [q], None))
        elif l == "Y":
            g.append(("RX", [q], pi / 2))
    qs = [q for _, q in act]
    for a, b in zip(qs[:-1], qs[1:]):
        g.append(("CNOT", [a, b], None))
    g.append(("RZ", [qs[-1]], 2.0 * angle))
    for a, b in reversed(list(zip(qs[:-1], qs[1:]))):
        g.append(("CNOT", [a, b], None))
    for l, q in act:
        if l == "X":
            g.append(("H", [q], None))
        elif l == "Y":
            g.append(("RX", [q], -pi / 2))
    return g


def ucc_circuit(nbqbits, cluster_ops_sp, hf_init_sp, parameters):
    """Gate list of the circuit built at reference get_energy_ucc.py:79-89."""
    gates = []
    for n_term, (op, th) in enumerate(zip(cluster_ops_sp, parameters)):
        if n_term == 0:
            gates += hf_gates(nbqbits, hf_init_sp, padded=True)
        for t in op.terms:
            c = complex(t.coeff)
            if c == 0:
                continue
            gates += pauli_rotation_gates(t.op, t.qbits, float(th) * c.real)
    return CircuitSummary(nbqbits, gates)


# ---- QUCCSD templates (Yordanov efficient excitation circuits) ------------------------------
def single_excitation_gates(exci, theta):
    """reference circuit.py:13-38"""
    a, b = exci
    g = [("CNOT", [i, i + 1], None) for i in range(a + 1, b - 1)]
    g += [("RZ", [a], pi / 2), ("RY", [b], -pi / 2), ("RZ", [b], -pi / 2), ("CNOT", [a, b], None),
          ("RY", [a], theta), ("RZ", [b], -pi / 2), ("CNOT", [a, b], None), ("RY", [a], -theta),
          ("H", [b], None), ("CNOT", [a, b], None)]
    g += [("CNOT", [b - 2 - i, b - 1 - i], None) for i in range(max(0, b - a - 2))]
    return g


def double_excitation_gates(exci, theta):
    """reference circuit.py:40-93"""
    e0, e1, e2, e3 = exci
    ry = lambda t: ("RY", [e0], t)
    h = lambda q: ("H", [q], None)
    cx = lambda c, t: ("CNOT", [c, t], None)
    g = [cx(e0, e1), cx(e2, e3)]
    g += [cx(i, i + 1) for i in range(e0 + 1, e1 - 1)]
    g += [cx(i, i + 1) for i in range(e2 + 1, e3 - 1)]
    g += [cx(e0, e2),
          ry(theta), h(e1), cx(e0, e1), ry(-theta), h(e3), cx(e0, e3), ry(theta), cx(e0, e1),
          ry(-theta), h(e2), cx(e0, e2), ry(theta), cx(e0, e1), ry(-theta), cx(e0, e3),
          ry(theta), h(e3), cx(e0, e1), ry(-2 * theta), h(e1), cx(e0, e2), h(e2), cx(e0, e2)]
    g += [cx(e1 - 2 - i, e1 - 1 - i) for i in range(max(0, e1 - e0 - 2))]
    g += [cx(e3 - 2 - i, e3 - 1 - i) for i in range(max(0, e3 - e2 - 2))]
    g += [cx(e0, e1), cx(e2, e3)]
    return g


def efficient_fermionic_ansatz_gates(list_exci, list_theta):
    """reference circuit.py:95-106 (indexing ``list_theta[i]`` raises IndexError when
    there are fewer parameters than excitations, as the reference does)."""
    g = []
    for i in range(len(list_exci)):
        if len(list_exci[i]) == 4:
            g += double_excitation_gates(list_exci[i], float(list_theta[i]))
        else:
            g += single_excitation_gates(list_exci[i], float(list_theta[i]))
    return g


def quccsd_circuit(nbqbits, hf_init_sp, cluster_ops, theta):
    """Gate list of the circuit built at reference get_energy_qucc.py:38-51."""
    list_exci = [list(op.terms[0].qbits) for op in cluster_ops]
    gates = hf_gates(nbqbits, hf_init_sp, padded=False)
    gates += efficient_fermionic_ansatz_gates(list_exci, theta)
    return CircuitSummary(nbqbits, gates)
