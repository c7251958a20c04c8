"""TEST INFRASTRUCTURE ONLY -- stand-in for ``qat.lang.AQASM`` (closed qat-lang 3.0.4).

Gate conventions: RX(t)=exp(-i t X/2), RY(t)=exp(-i t Y/2),
RZ(t)=diag(e^{-it/2}, e^{+it/2}), CNOT(control, target)."""
from ...core import Circuit


class _GateInst:
    __slots__ = ("name", "angle", "arity")

    def __init__(self, name, angle=None, arity=1):
        self.name, self.angle, self.arity = name, angle, arity


class _ParamGate:
    def __init__(self, name):
        self.name = name

    def __call__(self, angle):
        return _GateInst(self.name, float(angle))


X = _GateInst("X")
Y = _GateInst("Y")
Z = _GateInst("Z")
H = _GateInst("H")
I = _GateInst("I")
CNOT = _GateInst("CNOT", arity=2)
RX = _ParamGate("RX")
RY = _ParamGate("RY")
RZ = _ParamGate("RZ")


def _flat(args):
    out = []
    for a in args:
        if isinstance(a, (list, tuple)):
            out.extend(_flat(a))
        else:
            out.append(int(a))
    return out


class QRoutine:
    def __init__(self):
        self.gates = []  # (name, [local qubits], angle)

    def apply(self, gate, *qbits):
        q = _flat(qbits)
        if isinstance(gate, QRoutine):
            for (name, lq, angle) in gate.gates:
                self.gates.append((name, [q[i] for i in lq], angle))
        else:
            self.gates.append((gate.name, q, gate.angle))
        return self

    @property
    def arity(self):
        return 1 + max((max(q) for (_, q, _) in self.gates), default=-1)


class Program(QRoutine):
    def __init__(self):
        super().__init__()
        self.nbqbits = 0

    def qalloc(self, n):
        reg = list(range(self.nbqbits, self.nbqbits + int(n)))
        self.nbqbits += int(n)
        return reg

    def to_circ(self, **kw):
        return Circuit(self.nbqbits, self.gates)
