"""TEST INFRASTRUCTURE ONLY -- stand-in for ``qat.fermion.transforms``.
Only the Jordan-Wigner code is restated (the path BASELINE.json names uses
'JW' throughout); Bravyi-Kitaev / parity raise NotImplementedError."""
import numpy as np

from . import FermionHamiltonian, SpinHamiltonian


def transform_to_jw_basis(fermion_hamiltonian):
    if isinstance(fermion_hamiltonian, SpinHamiltonian):
        return fermion_hamiltonian
    return fermion_hamiltonian.to_spin()


def transform_to_bk_basis(h):
    raise NotImplementedError("qat shim: Bravyi-Kitaev transform not restated")


def transform_to_parity_basis(h):
    raise NotImplementedError("qat shim: parity transform not restated")


def get_jw_code(nbqbits):
    return np.identity(nbqbits, dtype=int)


def get_bk_code(nbqbits):
    raise NotImplementedError("qat shim: Bravyi-Kitaev code not restated")


def get_parity_code(nbqbits):
    raise NotImplementedError("qat shim: parity code not restated")


def recode_integer(integer, code):
    """Occupation integer (qubit 0 = MSB) -> encoded integer; identity for JW."""
    code = np.asarray(code)
    n = code.shape[0]
    bits = np.array([(int(integer) >> (n - 1 - q)) & 1 for q in range(n)], dtype=int)
    new_bits = code.dot(bits) % 2
    out = 0
    for q in range(n):
        out |= int(new_bits[q]) << (n - 1 - q)
    return out
