"""CPU checks of the C-ABI library: it loads, exports every symbol include/vqe_b200.h declares, and the
product path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "vqe_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vqe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from openvqe_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
    assert sorted(_lib.SYMBOLS) == declared  # the ctypes binding covers the whole header


def test_no_cpu_fallback():
    from openvqe_b200 import _lib
    lib = _lib.load()
    if lib.vqe_device_count() > 0:
        pytest.skip("a GPU is present")
    from openvqe_b200.engine import Engine
    with pytest.raises(_lib.VQEError, match="no CPU fallback"):
        Engine(4)
    from openvqe_b200.ucc_family.get_energy_ucc import EnergyUCC
    from tests.helpers import Ham, T
    with pytest.raises(_lib.VQEError):
        EnergyUCC().ucc_action([0.1], Ham(2, [T(1.0, "Z", [0])]), [Ham(2, [T(1.0, "XY", [0, 1])])], 2, [])


def test_lowering_bit_convention():
    from openvqe_b200.lowering import pack_operator, term_masks
    # qubit 0 is the most significant index bit
    assert term_masks("X", [0], 4) == (8, 0, 0)
    assert term_masks("ZY", [1, 3], 4) == (1, 5, 1)
    from tests.helpers import Ham, T
    p = pack_operator(Ham(3, [T(0.5, "XZ", [0, 2]), T(-1j, "Y", [1])], 0.25), with_constant=True)
    assert p.x.tolist() == [4, 2, 0] and p.z.tolist() == [1, 2, 0] and p.ny.tolist() == [0, 1, 0]
    assert p.cre.tolist() == [0.5, 0.0, 0.25] and p.cim.tolist() == [0.0, -1.0, 0.0]
