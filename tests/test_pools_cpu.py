"""Pool generation on Pauli lists (openvqe_b200/common_files/pools.py) against the reference's own
generators (openvqe/common_files/generator_excitations.py:83-156, 403-465, 555-609; qubit_pool.py:278-465) executed
unmodified through oracle/qat_shim -- their outputs are committed as packed Pauli lists in tests/golden/h6_full_pools.npz
(oracle/make_golden_r2.py).  Bit-exact: sizes (3 159 / 714 / 60 / 285), operator order, term order, coefficients."""
import os

import numpy as np
import pytest

from openvqe_b200.common_files import pools
from openvqe_b200.lowering import pack_pool

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "h6_full_pools.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _same(packed, gold, name):
    for key in ("x", "z", "ny", "cre", "cim", "offsets"):
        a, b = getattr(packed, key), gold[name + "_" + key]
        assert a.shape == b.shape, (name, key, a.shape, b.shape)
        assert np.array_equal(a, b), (name, key)


@pytest.mark.parametrize("name,make,size", [
    ("uccgsd", lambda: pools.uccgsd(6, 6, "JW"), 3159),
    ("spin_complement_gsd", lambda: pools.spin_complement_gsd(6, 6, "JW"), 714),
    ("singlet_upccgsd_k2", lambda: pools.singlet_upccgsd(6, "JW", 1), 60),
])
def test_fermionic_pools_equal_the_reference_pools(gold, name, make, size):
    pool_size, cluster_ops, cluster_ops_sp = make()
    assert pool_size == size == len(cluster_ops) == len(cluster_ops_sp)
    _same(pack_pool(cluster_ops_sp), gold, name)


def test_pools_keep_identically_zero_operators():
    """nothing is filtered; 4-orbital spin_complement_gsd has 175 operators, 64 of them zero."""
    size, _, sp = pools.spin_complement_gsd(2, 4, "JW")
    assert size == 175
    zero = [k for k, op in enumerate(sp) if all(complex(t.coeff) == 0 for t in op.terms)]
    assert len(zero) == 64
    assert pools.spin_complement_gsd(2, 3, "JW")[0] == 69 and pools.singlet_upccgsd(4, "JW", 0)[0] == 12


def test_qubit_pools(gold):
    size, yxxx = pools.generate_yxxx_pool(12)
    assert size == 285
    _same(pack_pool(yxxx), gold, "yxxx")
    for letters in ("YXXX", "XYXX", "XXYX", "XXXY"):
        n_ops, ops = getattr(pools, "generate_%s_pool" % letters.lower())(10)
        pk, ref = pools.packed_qubit_pool(10, letters), pack_pool(ops)
        for key in ("x", "z", "ny", "cre", "cim", "offsets"):
            assert np.array_equal(getattr(pk, key), getattr(ref, key))
    assert pools.generate_yxxx_pool(8)[0] == 50
    # the 'random' pool draws one of the four variants per position with numpy's global generator, as the reference does
    four = [getattr(pools, "generate_%s_pool" % l)(8)[1] for l in ("yxxx", "xyxx", "xxyx", "xxxy")]
    np.random.seed(7)
    n1, r1 = pools.generate_random_pool(*four)
    np.random.seed(7)
    picks = [np.random.randint(0, 4) for _ in range(50)]
    assert n1 == 50 and all(r1[i] is four[picks[i]][i] for i in range(50))


def test_dispatcher_wires_uccgsd_and_rejects_other_encodings():
    assert pools.generate_cluster_ops("uccgsd", 2, 2)[0] == 65
    assert pools.generate_cluster_ops("spin_complement_gsd", 2, 2)[0] == 21
    assert pools.generate_cluster_ops("no_such_pool", 2, 2) is None
    with pytest.raises(NotImplementedError):
        pools.uccgsd(2, 2, "Bravyi-Kitaev")


def test_normal_ordering_rules():
    """Anticommutation algebra of the normal ordering (creators left / ascending, annihilators right / ascending)."""
    # C_2 c_0 C_1 c_3 = - C_2 C_1 c_0 c_3 = + C_1 C_2 c_0 c_3
    assert pools.order_fermionic_term(1.0, "CcCc", [2, 0, 1, 3]) == [(1.0, "CCcc", [1, 2, 0, 3])]
    # a repeated creator kills the term: C_1 c_0 C_1 c_3 = - C_1 C_1 c_0 c_3 = 0
    assert pools.order_fermionic_term(1.0, "CcCc", [1, 0, 1, 3]) == []
    # contraction: C_2 c_1 C_1 c_3 = C_2 (1 - C_1 c_1) c_3 = C_2 c_3 + C_1 C_2 c_1 c_3
    assert pools.order_fermionic_term(1.0, "CcCc", [2, 1, 1, 3]) == [(1.0, "Cc", [2, 3]), (1.0, "CCcc", [1, 2, 1, 3])]
    # already ordered terms pass through; annihilators are sorted with the sign of the permutation
    assert pools.order_fermionic_term(0.5, "CCcc", [0, 1, 3, 2]) == [(-0.5, "CCcc", [0, 1, 2, 3])]


def test_snap_ties_documented_deviation():
    """The one place where the drop-in's selection can differ from the reference's float-equality semantics
    (DESIGN.md section 4 item 1), pinned: what snap_ties does to exact zeros, to reduction noise, to true ties and -- the
    deviation -- to two DISTINCT gradients closer than 1e-12 relative."""
    from openvqe_b200._hotpath import snap_ties
    from openvqe_b200.common_files.sorted_gradient import abs_sort_desc, corresponding_index, index_without_0, value_without_0

    def selection(grads):   # the order in which the ADAPT loops pick operators (fermionic_adapt_vqe.py:165-180)
        mags = [abs(v) for v in grads]
        vals, idx = value_without_0(mags), index_without_0(mags)
        return corresponding_index(vals, idx, abs_sort_desc(list(vals)))
    # reduction noise below 1e-14 Ha becomes the exact zero the reference's scipy path produces
    assert snap_ties([0.3, 3e-17, -0.1, 0.0]) == [0.3, 0.0, -0.1, 0.0]
    # spin-complement partners (+g, -g) that differ in the last bit: magnitudes equalised to the lowest index, signs kept
    g = 0.27328246013490346
    out = snap_ties([g, 0.5, -g * (1 + 2e-16)])
    assert out[0] == g and out[2] == -g and out[1] == 0.5
    # ... so the reference's helpers resolve the tie to the lowest pool index, as with scipy's exact ties
    assert selection(out) == [1, 0, 2]
    # well-separated gradients (1e-9 relative) are left alone, order preserved
    sep = [0.2, 0.2 * (1 + 1e-9), 0.1]
    assert snap_ties(sep) == sep
    # DEVIATION: two genuinely different gradients 5e-13 apart (relative) are merged; the reference would have picked index 1
    # (the strictly larger one), the drop-in reports a tie and picks index 0.  Physical gradients of distinct operators this
    # close do not occur in the fixtures: in the reference's own sweep over the 3 159-operator H6 pool the closest genuinely distinct
    # magnitudes are 9.5e-5 apart (relative), while exact ties show up split in the last bit (2e-16) -- the case the snap is for.
    near = [0.2, 0.2 * (1 + 5e-13), 0.1]
    merged = snap_ties(near)
    assert merged[0] == merged[1] == 0.2
    assert selection(near)[0] == 1 and selection(merged)[0] == 0
