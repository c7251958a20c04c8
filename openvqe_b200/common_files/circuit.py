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

    def to_job(self, job_type="SAMPLE", observable=None, **kw):
        """myQLM's ``Circuit.to_job`` shape, for ``openvqe_b200.qpu.B200QPU().submit(...)``."""
        return CircuitJob(self, job_type, observable)


class CircuitJob:
    __slots__ = ("circuit", "job_type", "observable")

    def __init__(self, circuit, job_type, observable):
        self.circuit, self.job_type, self.observable = circuit, job_type, observable


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


# ---- the QUCCSD templates as tabulated plane rotations ------------------------------------------
# Both templates only couple basis states that differ on ALL their core qubits (2 for a single, 4 for a double):
# their unitary is a set of plane rotations on the pairs (p, p ^ 1..1), one angle per occupation pattern, times a
# global phase.  The CNOT ladders of reference circuit.py:21-23/36-38 and :51-60/82-90 act only on the qubits strictly
# between the core qubits, which the core never touches, and are undone gate by gate at the end of the template: they
# cancel (the reference's "fermionic" excitation is a qubit excitation).  The angle of every pattern is linear in
# theta; the coefficients are derived ONCE from the template's own 2^k x 2^k unitary, built here from the gate
# matrices (myQLM conventions: RY(t) = exp(-i t Y / 2), RZ(t) = diag(e^{-it/2}, e^{it/2}), CNOT(control, target)),
# so the engine applies exactly the unitary of the gate list -- in one sweep per excitation instead of ~30.
def _gate_matrix(name, angle):
    import numpy as np
    if name == "X":
        return np.array([[0, 1], [1, 0]], dtype=complex)
    if name == "H":
        return np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2.0)
    c, s = np.cos(angle / 2.0), np.sin(angle / 2.0)
    if name == "RX":
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)
    if name == "RY":
        return np.array([[c, -s], [s, c]], dtype=complex)
    if name == "RZ":
        return np.array([[np.exp(-0.5j * angle), 0], [0, np.exp(0.5j * angle)]], dtype=complex)
    raise ValueError(name)


def _template_unitary(k, gates):
    """2^k x 2^k unitary of a gate list on k role qubits (role 0 = most significant bit)."""
    import numpy as np
    dim = 1 << k
    u = np.eye(dim, dtype=complex)
    idx = np.arange(dim)
    for name, qb, angle in gates:
        if name == "CNOT":
            cbit, tbit = 1 << (k - 1 - qb[0]), 1 << (k - 1 - qb[1])
            perm = np.where(idx & cbit, idx ^ tbit, idx)
            u = u[perm, :]
        else:
            m = _gate_matrix(name, angle)
            bit = 1 << (k - 1 - qb[0])
            lo, hi = idx[(idx & bit) == 0], idx[(idx & bit) == 0] | bit
            new = u.copy()
            new[lo, :] = m[0, 0] * u[lo, :] + m[0, 1] * u[hi, :]
            new[hi, :] = m[1, 0] * u[lo, :] + m[1, 1] * u[hi, :]
            u = new
    return u


def _template_table(k, builder):
    """-> (coef[2^(k-1)], phase_per_theta...) : angle of the pair (p, p ^ (2^k - 1)), p < 2^(k-1), is coef[p] * theta
    with a' = cos a - sin b, b' = sin a + cos b (a at p); ``phase`` is the template's global phase (theta-independent).
    Returns None when the template is not of that form (then the gate list is executed as is)."""
    import numpy as np
    dim, full = 1 << k, (1 << k) - 1
    out = []
    for theta in (0.05, 0.1):
        u = _template_unitary(k, builder(list(range(k)), theta))
        phase = u[0, 0] / abs(u[0, 0])
        v = u / phase
        if np.abs(v.imag).max() > 1e-13:
            return None
        v = v.real
        mask = np.ones((dim, dim), dtype=bool)
        ang = np.zeros(dim // 2)
        for p in range(dim // 2):
            q = p ^ full
            mask[p, p] = mask[q, q] = mask[p, q] = mask[q, p] = False
            if abs(v[p, p] - v[q, q]) > 1e-13 or abs(v[p, q] + v[q, p]) > 1e-13:
                return None
            ang[p] = np.arctan2(v[q, p], v[p, p])
        if np.abs(v[mask]).max() > 1e-13:
            return None
        out.append((ang / theta, phase))
    if np.abs(out[0][0] - out[1][0]).max() > 1e-9 or abs(out[0][1] - out[1][1]) > 1e-13:
        return None
    return np.round(out[1][0] * 2.0) / 2.0, complex(out[1][1])   # coefficients are multiples of 1/2


_TEMPLATES = {}


def template_tables():
    """{2: (coef, phase), 4: (coef, phase)} of the single / double excitation templates (None if not tabulable)."""
    if not _TEMPLATES:
        _TEMPLATES[2] = _template_table(2, single_excitation_gates)
        _TEMPLATES[4] = _template_table(4, double_excitation_gates)
    return _TEMPLATES


def _ladders_cancel(exci):
    """The ladder qubits (strictly between e0, e1 and between e2, e3) must not be core qubits."""
    if len(exci) == 2:
        return exci[0] != exci[1]
    e0, e1, e2, e3 = exci
    if len({e0, e1, e2, e3}) != 4:
        return False
    core = (e0, e1, e2, e3)
    return not any(e0 < q < e1 or e2 < q < e3 for q in core)


_PLANE_PROGRAMS = {}


def _plane_program(nbqbits, list_exci):
    """Theta-independent part of ``quccsd_plane_ops`` (cached per excitation list): X-masks, table offsets, a-side
    patterns, the angle coefficient and owner of every entry, and the number of single excitations (global phase)."""
    import numpy as np
    key = (nbqbits, tuple(tuple(int(q) for q in e) for e in list_exci))
    hit = _PLANE_PROGRAMS.get(key)
    if hit is not None:
        return hit
    tabs = template_tables()
    xs, offs, pats, coefs, owner = [], [0], [], [], []
    phase_counts = {}
    prog = None
    ok = True
    for i, exci in enumerate(key[1]):
        k = len(exci)
        tab = tabs.get(k)
        if tab is None or not _ladders_cancel(exci) or min(exci) < 0 or max(exci) >= nbqbits:
            ok = False
            break
        coef, ph = tab
        bits = [1 << (nbqbits - 1 - q) for q in exci]   # index bit of every role qubit
        x = 0
        for b in bits:
            x |= b
        top = 1 << (x.bit_length() - 1)
        for p in range(1 << (k - 1)):
            if coef[p] == 0.0:
                continue
            pat = 0
            for r in range(k):
                if (p >> (k - 1 - r)) & 1:
                    pat |= bits[r]
            c = float(coef[p])
            if pat & top:           # the engine's a-side has the highest X bit clear: swap the roles of a and b
                pat ^= x
                c = -c
            pats.append(pat)
            coefs.append(c)
            owner.append(i)
        xs.append(x)
        offs.append(len(pats))
        phase_counts[ph] = phase_counts.get(ph, 0) + 1
    if ok:
        prog = (np.array(xs, dtype=np.uint64), np.array(offs, dtype=np.int32), np.array(pats, dtype=np.uint64),
                np.array(coefs, dtype=np.float64), np.array(owner, dtype=np.int64), phase_counts)
    if len(_PLANE_PROGRAMS) > 64:
        _PLANE_PROGRAMS.clear()
    _PLANE_PROGRAMS[key] = prog
    return prog


def quccsd_plane_ops(nbqbits, list_exci, list_theta):
    """The QUCCSD ansatz (reference circuit.py:95-106) as tabulated plane rotations for
    ``Engine.apply_plane_rotations``: -> (xmask, offsets, pattern, cos, sin, global_phase), or None when some
    excitation is not of the tabulable form (the caller then executes the gate list)."""
    import numpy as np
    prog = _plane_program(nbqbits, list_exci)
    if prog is None:
        return None
    xs, offs, pats, coefs, owner, phase_counts = prog
    theta = np.asarray([float(t) for t in list_theta[:len(list_exci)]], dtype=np.float64)
    ang = coefs * theta[owner]
    phase = 1.0 + 0.0j
    for ph, cnt in phase_counts.items():
        phase *= ph ** cnt
    return xs, offs, pats, np.cos(ang), np.sin(ang), phase


def hf_index(nbqbits, hf_init_sp):
    """Basis index prepared by the reference's unpadded X gates (get_energy_qucc.py:40-45)."""
    idx = 0
    for name, qb, _ in hf_gates(nbqbits, hf_init_sp, padded=False):
        idx |= 1 << (nbqbits - 1 - qb[0])
    return idx
