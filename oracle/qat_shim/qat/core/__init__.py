"""TEST INFRASTRUCTURE ONLY -- minimal pure-Python stand-in for ``qat.core``.

myQLM (qat-core 1.8.5) is a closed binary that is not installable in this
sandbox.  This stand-in exposes just enough of ``Term`` / ``Observable`` /
``Circuit`` for the *unmodified* reference modules under
``/root/reference/openvqe`` to import and run, so that their scipy code paths
can be executed verbatim when generating golden vectors
(``oracle/make_golden.py``) and when timing the CPU baseline.

Nothing in the product package ``openvqe_b200`` imports this.
"""
from __future__ import annotations


class _Part:
    __slots__ = ("re", "im")

    def __init__(self, z):
        z = complex(z)
        self.re = z.real
        self.im = z.imag


class _CoeffBox:
    """Mimics ``term._coeff.complex_p.re / .im`` (used at
    reference openvqe/common_files/qubit_pool.py:729-732)."""

    __slots__ = ("complex_p",)

    def __init__(self, z):
        self.complex_p = _Part(z)


class Term:
    """``Term(coefficient, pauli_op, qbits)``: one Pauli / ladder-operator
    string.  ``op`` is a string over I,X,Y,Z (spin) or C,c (fermionic,
    C = creation), ``qbits`` the list of qubits each letter acts on."""

    def __init__(self, coefficient=1.0, pauli_op="", qbits=None, **kw):
        if "coeff" in kw:
            coefficient = kw["coeff"]
        self.coeff = coefficient
        self.op = str(pauli_op)
        self.qbits = list(qbits) if qbits is not None else []
        if len(self.op) != len(self.qbits):
            raise ValueError("Term: len(op) != len(qbits): %r %r" % (self.op, self.qbits))

    @property
    def _coeff(self):
        return _CoeffBox(self.coeff)

    def copy(self):
        return Term(self.coeff, self.op, list(self.qbits))

    def key(self):
        return (self.op, tuple(self.qbits))

    def __mul__(self, other):
        if isinstance(other, (int, float, complex)):
            return Term(self.coeff * other, self.op, list(self.qbits))
        return NotImplemented

    __rmul__ = __mul__

    def __repr__(self):
        return "%s * (%s|%s)" % (self.coeff, self.op, self.qbits)


class Op:
    """One gate of a flattened circuit.  ``str(op)`` contains ``gate='NAME'``
    because the reference's gate counter greps for that substring
    (reference openvqe/common_files/circuit.py:186-205)."""

    __slots__ = ("gate", "qbits", "name", "angle")

    def __init__(self, gate, qbits, name=None, angle=None):
        self.gate = gate          # key in the gate dictionary ('CNOT', 'H', 'X', '_0', ...)
        self.qbits = list(qbits)
        self.name = name or gate  # abstract gate name ('RX', 'RZ', ...)
        self.angle = angle

    def __repr__(self):
        return "Op(gate='%s', qbits=%s, type=GATETYPE)" % (self.gate, self.qbits)


class Circuit:
    """Flattened gate list.  Parametrised gates receive dictionary keys
    ``_0, _1, ...`` in order of first appearance of each distinct
    (name, angle), as observed in the stored notebook outputs (SURVEY
    Appendix B item 10)."""

    def __init__(self, nbqbits, gates):
        self.nbqbits = nbqbits
        self.ops = []
        self._gate_keys = {}
        for (name, qbits, angle) in gates:
            if angle is None:
                key = name
            else:
                ident = (name, float(angle))
                if ident not in self._gate_keys:
                    self._gate_keys[ident] = "_%d" % len(self._gate_keys)
                key = self._gate_keys[ident]
            self.ops.append(Op(key, qbits, name, angle))

    def to_job(self, job_type="SAMPLE", observable=None, **kw):
        from ..qpus import Job
        return Job(self, job_type, observable)


class Observable:
    pass
