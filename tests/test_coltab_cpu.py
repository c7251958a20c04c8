"""CPU checks of the item table the real-layout rotation kernel k_col_tab walks on the GPU (one 32-bit word per item of
a collapsed run: shared-memory byte offset, Z parity, pattern number).  The host interpreter vqe_debug_coltab_host plans
the program exactly as vqe_apply_pauli_rotations does, builds the table with the library's own routine and applies it
tile by tile with the kernel's per-item arithmetic, so a wrong deposit, pattern, swizzle or sign is caught here, without a
GPU, against the string-by-string numpy oracle (reference: one state sweep per Pauli rotation,
openvqe/ucc_family/get_energy_ucc.py:8-50 through myQLM's simulator)."""
import ctypes as C

import numpy as np
import pytest

from oracle import statevector_oracle as orc
from tests.helpers import jw_excitation


def _coltab(n, xs, zs, nys, angs, psi_re, tile_bits=13, low_bits=4, form=1):
    from openvqe_b200 import _lib
    lib = _lib.load()
    x = np.ascontiguousarray(xs, dtype=np.uint64)
    z = np.ascontiguousarray(zs, dtype=np.uint64)
    ny = np.ascontiguousarray(nys, dtype=np.int32)
    a = np.ascontiguousarray(angs, dtype=np.float64)
    psi = np.ascontiguousarray(psi_re, dtype=np.float64).copy()
    n_pass, n_words = C.c_int32(), C.c_int32()
    p = lambda v: v.ctypes.data_as(C.c_void_p)
    _lib.check(lib.vqe_debug_coltab_host(n, tile_bits, low_bits, form, len(x), p(x), p(z), p(ny), p(a), p(psi),
                                         C.byref(n_pass), C.byref(n_words)))
    return psi, n_pass.value, n_words.value


def _jw_program(n, n_gen, seed):
    from openvqe_b200.lowering import pack_operator
    rng = np.random.default_rng(seed)
    xs, zs, nys, angs = [], [], [], []
    for g in range(n_gen):
        if g % 4 == 3:
            p, q = sorted(rng.choice(n, size=2, replace=False).tolist())
            pk = pack_operator(jw_excitation(n, [q], [p]))
        else:
            p, q, r, s = sorted(rng.choice(n, size=4, replace=False).tolist())
            pk = pack_operator(jw_excitation(n, [r, s], [p, q]))
        th = float(rng.uniform(-0.8, 0.8))
        for k in range(len(pk)):
            xs.append(int(pk.x[k])); zs.append(int(pk.z[k])); nys.append(int(pk.ny[k])); angs.append(th * float(pk.cre[k]))
    return xs, zs, nys, angs


def _oracle_state(n, start, xs, zs, nys, angs):
    ref = start.astype(np.complex128)
    for x, z, ny, a in zip(xs, zs, nys, angs):
        ref = orc.pauli_rotation(ref, x, z, ny, a)
    assert np.max(np.abs(ref.imag)) < 1e-14  # odd-ny rotations keep a real state real
    return ref.real


@pytest.mark.parametrize("form", [1, 0])
@pytest.mark.parametrize("n,tile_bits,low_bits", [(14, 13, 4), (16, 13, 4), (15, 12, 4), (15, 13, 5), (10, 13, 4)])
def test_item_table_jw_excitations_match_oracle(n, tile_bits, low_bits, form):
    """JW singles and doubles (runs with ONE active occupation pattern) from |HF>: several tiles per pass, several passes,
    outside-tile Z letters; amplitudes outside the reachable sector stay exactly 0.0."""
    xs, zs, nys, angs = _jw_program(n, 14, 4100 + n + tile_bits)
    hf = ((1 << (n // 2)) - 1) << (n - n // 2)
    start = np.zeros(1 << n)
    start[hf] = 1.0
    got, n_pass, n_words = _coltab(n, xs, zs, nys, angs, start, tile_bits, low_bits, form)
    ref = _oracle_state(n, start, xs, zs, nys, angs)
    assert n_pass >= 1 and n_words > 0
    assert np.max(np.abs(got - ref)) < 1e-12
    assert np.all(got[np.abs(ref) < 1e-13] == 0.0)


@pytest.mark.parametrize("form", [1, 0])
def test_item_table_runs_with_several_active_patterns(form):
    """Same-X-mask runs whose angles are unrelated (not the JW image of an excitation): several occupation patterns of the
    run have a non-zero angle, so the items carry a pattern number and the kernel reads (cos, sin) per item."""
    from openvqe_b200.lowering import term_masks
    n = 15
    rng = np.random.default_rng(77)
    xs, zs, nys, angs = [], [], [], []
    for g in range(10):
        q = sorted(rng.choice(n, size=4, replace=False).tolist())
        # 7 or 8 strings: enough for the planner's cost model to prefer ONE tabulated plane rotation over 7-8 sweeps
        for op in ["XXXY", "XXYX", "XYXX", "YXXX", "XYYY", "YXYY", "YYXY", "YYYX"][: 7 + g % 2]:
            x, z, ny = term_masks(op, q, n)
            xs.append(x); zs.append(z); nys.append(ny); angs.append(float(rng.uniform(-0.5, 0.5)))
    start = rng.normal(size=1 << n)
    start /= np.linalg.norm(start)
    got, _, n_words = _coltab(n, xs, zs, nys, angs, start, form=form)
    ref = _oracle_state(n, start, xs, zs, nys, angs)
    assert np.max(np.abs(got - ref)) < 1e-12
    assert n_words > 10 * (1 << 9)  # more than one pattern per run


def test_item_table_refuses_programs_outside_the_real_collapsed_form():
    from openvqe_b200 import _lib
    from openvqe_b200.lowering import term_masks
    n = 14
    x, z, ny = term_masks("XXYY", [1, 4, 6, 9], n)  # even ny: +-i phase, the state does not stay real
    x2, z2, ny2 = term_masks("YYXX", [1, 4, 6, 9], n)
    with pytest.raises(_lib.VQEError):
        _coltab(n, [x, x2], [z, z2], [ny, ny2], [0.3, 0.2], np.ones(1 << n) / 2.0 ** 7)


def _banks(n, xs, zs, nys, angs, tile_bits=13, low_bits=4):
    from openvqe_b200 import _lib
    lib = _lib.load()
    x = np.ascontiguousarray(xs, dtype=np.uint64)
    z = np.ascontiguousarray(zs, dtype=np.uint64)
    ny = np.ascontiguousarray(nys, dtype=np.int32)
    a = np.ascontiguousarray(angs, dtype=np.float64)
    acc, con, una = C.c_int64(), C.c_int64(), C.c_int64()
    p = lambda v: v.ctypes.data_as(C.c_void_p)
    _lib.check(lib.vqe_debug_coltab_banks(n, tile_bits, low_bits, len(x), p(x), p(z), p(ny), p(a), C.byref(acc), C.byref(con),
                                          C.byref(una)))
    return acc.value, con.value, una.value


def test_item_order_leaves_only_the_unavoidable_bank_conflicts(monkeypatch):
    """A 64-bit shared-memory access is served per half-warp.  With the item order the tables are built with, the 16 elements of
    a half-warp sit in 16 different 8-byte bank pairs for every run whose free tile positions can reach all bank pairs at all
    (not when the X-mask fixes tile position 0, or both positions that feed one bank bit under the 128-byte swizzle); the plain
    ascending enumeration also conflicts when the X-mask holds position 2 or 3."""
    n = 16
    xs, zs, nys, angs = _jw_program(n, 60, 9001)
    acc, con, una = _banks(n, xs, zs, nys, angs)
    assert acc > 1000 and con == una
    monkeypatch.setenv("VQE_COL_LANE_ORDER", "0")
    acc0, con0, una0 = _banks(n, xs, zs, nys, angs)
    assert acc0 == acc and una0 == una and con0 > con
