"""TEST INFRASTRUCTURE ONLY -- numpy state-vector stand-in for
``qat.qpus.get_default_qpu()`` (myQLM CLinalg / PyLinalg).  Gate-by-gate full
state sweeps: this mirrors the cost structure of the reference's simulator and
is what ``bench.py --impl reference`` times."""
import numpy as np

_S2 = 1.0 / np.sqrt(2.0)


def _matrix(name, angle):
    if name == "X":
        return np.array([[0, 1], [1, 0]], dtype=np.complex128)
    if name == "Y":
        return np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
    if name == "Z":
        return np.array([[1, 0], [0, -1]], dtype=np.complex128)
    if name == "H":
        return np.array([[_S2, _S2], [_S2, -_S2]], dtype=np.complex128)
    if name == "I":
        return np.eye(2, dtype=np.complex128)
    c, s = np.cos(angle / 2.0), np.sin(angle / 2.0)
    if name == "RX":
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)
    if name == "RY":
        return np.array([[c, -s], [s, c]], dtype=np.complex128)
    if name == "RZ":
        return np.array([[c - 1j * s, 0], [0, c + 1j * s]], dtype=np.complex128)
    raise ValueError("qat shim: unknown gate %r" % name)


def simulate(circuit):
    n = circuit.nbqbits
    psi = np.zeros((2,) * n, dtype=np.complex128)
    psi[(0,) * n] = 1.0
    for op in circuit.ops:
        if op.name == "CNOT":
            c, t = op.qbits
            sl = [slice(None)] * n
            sl[c] = 1
            sub = psi[tuple(sl)]
            ax = t - 1 if t > c else t
            psi[tuple(sl)] = np.flip(sub, axis=ax)
        else:
            (q,) = op.qbits
            m = _matrix(op.name, op.angle)
            psi = np.moveaxis(np.tensordot(m, psi, axes=([1], [q])), 0, q)
    return np.ascontiguousarray(psi).reshape(-1)


def pauli_expectation(psi, term, n):
    xm = zm = ny = 0
    for letter, q in zip(term.op, term.qbits):
        bit = n - 1 - q
        if letter in "XY":
            xm |= 1 << bit
        if letter in "YZ":
            zm |= 1 << bit
        if letter == "Y":
            ny += 1
    idx = np.arange(psi.shape[0], dtype=np.int64)
    par = np.zeros_like(idx)
    v = idx & zm
    while v.any():
        par ^= v & 1
        v >>= 1
    ppsi = np.empty_like(psi)
    ppsi[idx ^ xm] = (1j ** ny) * (1 - 2 * par) * psi
    return np.vdot(psi, ppsi)


class _State:
    def __init__(self, i):
        self.int = int(i)


class Sample:
    def __init__(self, i, amp):
        self.state = _State(i)
        self.amplitude = complex(amp)
        self.probability = abs(amp) ** 2


class Result:
    def __init__(self, value=None, samples=None):
        self.value = value
        self.raw_data = samples or []

    def __iter__(self):
        return iter(self.raw_data)


class Job:
    def __init__(self, circuit, job_type, observable):
        self.circuit, self.job_type, self.observable = circuit, job_type, observable


class NumpyQPU:
    def submit(self, job):
        psi = simulate(job.circuit)
        if job.job_type == "OBS":
            obs = job.observable
            val = complex(obs.constant_coeff) * np.vdot(psi, psi)
            for t in obs.terms:
                if t.coeff == 0:
                    continue
                val += complex(t.coeff) * pauli_expectation(psi, t, obs.nbqbits)
            return Result(value=float(val.real))
        nz = np.nonzero(psi)[0]
        return Result(samples=[Sample(i, psi[i]) for i in nz])


def get_default_qpu():
    return NumpyQPU()
