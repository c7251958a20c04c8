"""CPU checks of the "lean" Pauli-sum tables (DevFlat2 entries, additive tables, outside-tile constants) that
k_expect_lean / k_apply_lean walk on the GPU: the host interpreter vqe_debug_lean_host uses the kernels' own decode
routine, so a wrong pattern, deposit mask, sign or table is caught here, without a GPU, against the numpy oracle.

Inputs: the 12-qubit H6/STO-3G Hamiltonian of the golden fixtures (918 Pauli strings: every structure a molecular
Hamiltonian has -- XXYY-type families, number-operator-dressed hopping terms) with its qubits relabelled into a larger
register, so that tiles, outside-tile Z letters and (for sharded contexts) global qubits all occur."""
import ctypes as C

import numpy as np
import pytest

from oracle import statevector_oracle as orc
from tests.helpers import ham_from_json, load_golden


def _lean(n, x, z, ny, cre, psi, n_global=0, rank=0, tile_bits=12, low_bits=5, want_sigma=True):
    from openvqe_b200 import _lib
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.uint64)
    z = np.ascontiguousarray(z, dtype=np.uint64)
    ny = np.ascontiguousarray(ny, dtype=np.int32)
    cre = np.ascontiguousarray(cre, dtype=np.float64)
    cim = np.zeros_like(cre)
    psi = np.ascontiguousarray(psi, dtype=np.complex128)
    sigma = np.zeros_like(psi) if want_sigma else None
    out = C.c_double()
    nlean, nfat = C.c_int32(), C.c_int32()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _lib.check(lib.vqe_debug_lean_host(n, n_global, rank, tile_bits, low_bits, len(x), p(x), p(z), p(ny), p(cre), p(cim),
                                       p(psi), C.byref(out), C.byref(nlean), C.byref(nfat), p(sigma) if want_sigma else None))
    return out.value, nlean.value, nfat.value, sigma


def _relabelled_h6(n, seed):
    """(x, z, ny, c) of the H6 Hamiltonian with qubit q of the 12 mapped to a random qubit of an n-qubit register."""
    from openvqe_b200.lowering import term_masks
    fx = load_golden("h6_sto3g.json.gz")
    ham = ham_from_json(fx["hamiltonian"])
    rng = np.random.default_rng(seed)
    where = np.sort(rng.choice(n, size=12, replace=False))
    perm = rng.permutation(12)
    xs, zs, nys, cs = [], [], [], []
    for t in ham.terms:
        qb = [int(where[perm[q]]) for q in t.qbits]
        x, z, ny = term_masks(t.op, qb, n)
        xs.append(x); zs.append(z); nys.append(ny); cs.append(float(np.real(t.coeff)))
    return np.array(xs, dtype=np.uint64), np.array(zs, dtype=np.uint64), np.array(nys, dtype=np.int32), np.array(cs)


def _oracle(n, x, z, ny, c, psi, sel):
    e = 0.0
    sig = np.zeros_like(psi)
    for k in np.nonzero(sel)[0]:
        ppsi = orc.apply_pauli(psi, int(x[k]), int(z[k]), int(ny[k]))
        e += c[k] * np.vdot(psi, ppsi).real
        sig += c[k] * ppsi
    return e, sig


@pytest.mark.parametrize("n,seed,low_bits,real", [(12, 1, 5, False), (14, 2, 5, True), (15, 3, 4, False), (16, 4, 3, False),
                                                  (16, 5, 5, True)])
def test_lean_entries_reproduce_offdiagonal_expectation_and_sigma(n, seed, low_bits, real):
    x, z, ny, c = _relabelled_h6(n, seed)
    rng = np.random.default_rng(100 + seed)
    psi = rng.normal(size=1 << n) + (0 if real else 1j) * rng.normal(size=1 << n)
    psi = (psi / np.linalg.norm(psi)).astype(np.complex128)
    e, nlean, nfat, sigma = _lean(n, x, z, ny, c, psi, low_bits=low_bits)
    offdiag = x != 0
    assert nlean == int(offdiag.sum())           # every off-diagonal string of a molecular Hamiltonian is tabulated
    assert nfat == int((~offdiag & (z != 0)).sum()) + int(((x == 0) & (z == 0)).sum())
    e_ref, sig_ref = _oracle(n, x, z, ny, c, psi, offdiag)
    assert abs(e - e_ref) < 1e-12
    assert np.max(np.abs(sigma - sig_ref)) < 1e-12


def test_lean_entries_on_a_shard_with_global_z_letters():
    """Sharded context (2 global qubits): groups whose X-mask is local stay lean; Z letters on global qubits enter as
    per-rank signs.  Sum over the four ranks' partial sums = oracle."""
    n, g = 16, 2
    x, z, ny, c = _relabelled_h6(n, 7)
    rng = np.random.default_rng(77)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    nl = n - g
    local = (x != 0) & ((x >> np.uint64(nl)) == 0)
    tot, sig = 0.0, np.zeros_like(psi)
    for r in range(1 << g):
        shard = np.ascontiguousarray(psi[r << nl:(r + 1) << nl])
        e, nlean, nfat, s = _lean(n, x, z, ny, c, shard, n_global=g, rank=r)
        assert nlean == int(local.sum())
        tot += e
        sig[r << nl:(r + 1) << nl] = s
    e_ref, sig_ref = _oracle(n, x, z, ny, c, psi, local)
    assert abs(tot - e_ref) < 1e-12
    assert np.max(np.abs(sig - sig_ref)) < 1e-12


def test_groups_outside_the_lean_form_stay_on_the_general_path():
    """Odd-ny strings, complex coefficients, X-masks wider than 4 letters and Z-variants that differ by two letters
    outside the X positions are not tabulated (n_lean_terms counts only eligible groups)."""
    from openvqe_b200.lowering import term_masks
    n = 13
    terms = [("XXYY", [0, 3, 5, 9], 0.3), ("YYXX", [0, 3, 5, 9], -0.3),          # lean (family on the X positions)
             ("XZZX", [1, 2, 3, 4], 0.2), ("XZZXZ", [1, 2, 3, 4, 8], 0.1),        # lean (one further Z letter)
             ("XY", [6, 7], 0.5),                                                  # odd ny: general path
             ("XXXXXX", [0, 1, 2, 3, 4, 5], 0.1),                                  # 6 X letters: general path
             ("XX", [10, 12], 0.4), ("XXZZ", [10, 12, 2, 5], 0.2)]                 # two further Z letters: general path
    x, z, ny, c = [], [], [], []
    for op, qb, cf in terms:
        a, b, k = term_masks(op, qb, n)
        x.append(a); z.append(b); ny.append(k); c.append(cf)
    rng = np.random.default_rng(3)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    e, nlean, nfat, _ = _lean(n, x, z, ny, c, psi)
    assert (nlean, nfat) == (4, 4)
    x, z, ny, c = (np.array(v) for v in (x, z, ny, c))
    sel = np.zeros(len(c), dtype=bool)
    sel[:4] = True
    e_ref, _ = _oracle(n, x, z, ny, c, psi, sel)
    assert abs(e - e_ref) < 1e-13


@pytest.mark.parametrize("form", ["1", "2"])
@pytest.mark.parametrize("n,seed,low_bits", [(13, 11, 5), (15, 12, 4), (16, 13, 4), (14, 14, 3)])
def test_real_layout_entries_reproduce_the_expectation(n, seed, low_bits, form, monkeypatch):
    """The real-layout forms of the entries (state kept as n_amp doubles, 8-byte tile elements, its own 128-byte swizzle) on a
    tile of doubles = the oracle.  Form 1: lean_entry_to_rl + the kernels' decode routine (k_expect_lean<true, T, true>);
    form 2: the twin lowered for the real layout with its lane table and conflict-free lane order (k_expect_rlp, the default)."""
    x, z, ny, c = _relabelled_h6(n, seed)
    rng = np.random.default_rng(200 + seed)
    psi = rng.normal(size=1 << n)
    psi = (psi / np.linalg.norm(psi)).astype(np.complex128)
    monkeypatch.setenv("VQE_DEBUG_LEAN_RL", form)
    e, nlean, nfat, _ = _lean(n, x, z, ny, c, psi, low_bits=low_bits, want_sigma=False)
    offdiag = x != 0
    assert nlean == int(offdiag.sum())
    e_ref, _ = _oracle(n, x, z, ny, c, psi, offdiag)
    assert abs(e - e_ref) < 1e-12


@pytest.mark.parametrize("rl", [False, True])
def test_tensor_map_shapes_address_the_tiles(rl):
    """Host emulation of the TMA box traversal (vqe_debug_tma_check / _rl): every element of every sampled tile lands where
    the per-segment gather would put it -- interleaved complex and real layout, natural and 128-byte-swizzled shapes."""
    from openvqe_b200 import _lib
    lib = _lib.load()
    fn = lib.vqe_debug_tma_check_rl if rl else lib.vqe_debug_tma_check
    rng = np.random.default_rng(9)
    checked = 0
    for n_local in (12, 16, 20, 24, 30):
        for low_bits in (3, 4, 5):
            for _ in range(6):
                k = int(rng.integers(0, 8))
                hi = rng.choice(np.arange(low_bits, n_local), size=min(k, n_local - low_bits), replace=False) if n_local > low_bits else []
                need = 0
                for b in hi:
                    need |= 1 << int(b)
                for swz in (1, -1):
                    nreq, dims = C.c_int32(), C.c_int32()
                    bad = fn(n_local, need, 12, low_bits, swz * 5, C.byref(nreq), C.byref(dims))
                    if bad == -1:
                        continue  # the plan has no tensor-map form of this kind (the kernels then take another path)
                    assert bad == 0, (rl, n_local, low_bits, hex(need), swz, bad)
                    checked += 1
    assert checked > 60


@pytest.mark.parametrize("n,n_global,rank", [(12, 0, 0), (14, 0, 0), (16, 0, 0), (15, 2, 3), (13, 1, 0)])
def test_diagonal_part_as_a_quadratic_form(n, n_global, rank):
    """k_expect_diag2_rl evaluates the X-mask-0 strings of a molecular Hamiltonian (Z_p, Z_p Z_q) as a quadratic form in the Z
    letters, chunk by chunk (K(o) + A[t & 31] + B[t >> 5] + T(t)); the host interpreter builds the form with the library's own
    routine and walks the same decomposition -- also on a shard, where the rank bits enter as outside-chunk signs."""
    from openvqe_b200 import _lib
    lib = _lib.load()
    x, z, ny, c = _relabelled_h6(n, 40 + n)
    nl = n - n_global
    rng = np.random.default_rng(700 + n)
    psi_full = np.zeros(1 << n)
    lo, hi = rank << nl, (rank + 1) << nl
    psi_full[lo:hi] = rng.normal(size=1 << nl)
    shard = np.ascontiguousarray(psi_full[lo:hi])
    out, form = C.c_double(), C.c_int32()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    cim = np.zeros_like(c)
    _lib.check(lib.vqe_debug_diag2_host(n, n_global, rank, len(x), p(x), p(z), p(ny.astype(np.int32)), p(c), p(cim), p(shard),
                                        C.byref(out), C.byref(form)))
    assert form.value == 1
    diag = x == 0
    e_ref, _ = _oracle(n, x, z, ny, c, psi_full.astype(np.complex128), diag)
    assert abs(out.value - e_ref) < 1e-11 * max(1.0, abs(e_ref))
    # a three-Z string is not a quadratic form: the GPU path keeps the general pass
    x2, z2 = np.append(x, np.uint64(0)), np.append(z, np.uint64(0b111))
    ny2, c2 = np.append(ny, 0).astype(np.int32), np.append(c, 0.25)
    _lib.check(lib.vqe_debug_diag2_host(n, n_global, rank, len(x2), p(x2), p(z2), p(ny2), p(c2), p(np.zeros_like(c2)), p(shard),
                                        C.byref(out), C.byref(form)))
    assert form.value == 0
