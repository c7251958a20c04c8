"""ctypes binding of the C ABI declared in include/vqe_b200.h.

The shared library is built in-tree (``openvqe_b200/csrc/build.sh`` or
``__graft_entry__.build()``).  If it is missing, or no CUDA device is present,
every numeric call raises: there is deliberately no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VQE_B200_LIB") or os.path.join(_HERE, "csrc", "libvqe_b200.so")  # env: A/B testing of builds

# every symbol include/vqe_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "vqe_last_error", "vqe_version", "vqe_device_count", "vqe_create", "vqe_destroy", "vqe_n_qubits",
    "vqe_launch_count", "vqe_profile_enable", "vqe_profile_read", "vqe_timer_begin", "vqe_timer_end",
    "vqe_transfer_bytes", "vqe_set_basis_state", "vqe_set_state",
    "vqe_get_state", "vqe_copy_buffer", "vqe_apply_pauli_rotations", "vqe_apply_gates",
    "vqe_paulisum_create", "vqe_paulisum_destroy", "vqe_paulisum_groups", "vqe_paulisum_passes",
    "vqe_expectation", "vqe_apply_paulisum", "vqe_pool_overlaps", "vqe_apply_exp_paulisum",
    "vqe_overlap_host", "vqe_norm2", "vqe_inner", "vqe_buffer_ptr", "vqe_synchronize",
    # sharded state
    "vqe_create_shard", "vqe_shard_info", "vqe_shard_export", "vqe_shard_attach_ipc", "vqe_shard_attach_local",
    "vqe_shard_barrier", "vqe_shard_status", "vqe_group_apply_pauli_rotations", "vqe_group_apply_gates",
    "vqe_group_expectation", "vqe_group_apply_paulisum", "vqe_group_pool_overlaps", "vqe_plan_rotations",
    "vqe_apply_plane_rotations", "vqe_scale_state", "vqe_apply_pauli_rotations_buf", "vqe_plan_paulisum", "vqe_debug_lean_host", "vqe_peer_bytes", "vqe_debug_tma_check",
    "vqe_axpby", "vqe_debug_tma_check_rl", "vqe_state_layout", "vqe_relabel_stats", "vqe_debug_coltab_host", "vqe_debug_diag2_host", "vqe_debug_coltab_banks",
]
IPC_HANDLE_BYTES = 64
SHARD_FLAGS = 3


class VQEError(RuntimeError):
    pass


_lib = None


def load():
    """Load (once) and return the ctypes handle; raises VQEError if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VQEError(
            "CUDA extension %s not found: build it with openvqe_b200/csrc/build.sh "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, u64, i32, dbl = C.c_void_p, C.c_uint64, C.c_int32, C.c_double
    P = C.POINTER
    sig = {
        "vqe_last_error": (C.c_char_p, []),
        "vqe_version": (C.c_int, []),
        "vqe_device_count": (C.c_int, []),
        "vqe_create": (C.c_int, [P(vp), C.c_int, C.c_int]),
        "vqe_destroy": (None, [vp]),
        "vqe_n_qubits": (C.c_int, [vp]),
        "vqe_state_layout": (C.c_int, [vp]),
        "vqe_launch_count": (u64, [vp]),
        "vqe_profile_enable": (C.c_int, [vp, C.c_int]),
        "vqe_profile_read": (C.c_int, [vp, C.c_int, P(dbl), P(u64), C.c_int]),
        "vqe_timer_begin": (C.c_int, [vp]),
        "vqe_timer_end": (C.c_int, [vp, P(dbl)]),
        "vqe_transfer_bytes": (C.c_int, [vp, P(u64), P(u64), C.c_int]),
        "vqe_set_basis_state": (C.c_int, [vp, u64]),
        "vqe_set_state": (C.c_int, [vp, C.c_int, vp]),
        "vqe_get_state": (C.c_int, [vp, C.c_int, vp]),
        "vqe_copy_buffer": (C.c_int, [vp, C.c_int, C.c_int]),
        "vqe_apply_pauli_rotations": (C.c_int, [vp, C.c_int, vp, vp, vp, vp]),
        "vqe_apply_gates": (C.c_int, [vp, C.c_int, vp, vp, vp, vp]),
        "vqe_apply_pauli_rotations_buf": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp]),
        "vqe_paulisum_create": (C.c_int, [vp, P(vp), C.c_int, vp, vp, vp, vp, vp]),
        "vqe_paulisum_destroy": (None, [vp]),
        "vqe_paulisum_groups": (C.c_int, [vp]),
        "vqe_paulisum_passes": (C.c_int, [vp]),
        "vqe_expectation": (C.c_int, [vp, C.c_int, vp, P(dbl)]),
        "vqe_apply_paulisum": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "vqe_pool_overlaps": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
        "vqe_apply_exp_paulisum": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp, dbl]),
        "vqe_overlap_host": (C.c_int, [vp, C.c_int, vp, P(dbl)]),
        "vqe_norm2": (C.c_int, [vp, C.c_int, P(dbl)]),
        "vqe_inner": (C.c_int, [vp, C.c_int, C.c_int, P(dbl)]),
        "vqe_buffer_ptr": (C.c_int, [vp, C.c_int, P(vp), P(u64)]),
        "vqe_synchronize": (C.c_int, [vp]),
        "vqe_create_shard": (C.c_int, [P(vp), C.c_int, C.c_int, C.c_int, C.c_int]),
        "vqe_shard_info": (C.c_int, [vp, P(C.c_int), P(C.c_int), P(C.c_int)]),
        "vqe_shard_export": (C.c_int, [vp, C.c_int, vp]),
        "vqe_shard_attach_ipc": (C.c_int, [vp, C.c_int, C.c_int, vp]),
        "vqe_shard_attach_local": (C.c_int, [vp, vp]),
        "vqe_shard_barrier": (C.c_int, [vp]),
        "vqe_shard_status": (C.c_int, [vp]),
        "vqe_group_apply_pauli_rotations": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp]),
        "vqe_group_apply_gates": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp]),
        "vqe_group_expectation": (C.c_int, [vp, C.c_int, C.c_int, vp, P(dbl)]),
        "vqe_group_apply_paulisum": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp]),
        "vqe_group_pool_overlaps": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
        "vqe_apply_plane_rotations": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp]),
        "vqe_scale_state": (C.c_int, [vp, C.c_int, dbl, dbl]),
        "vqe_axpby": (C.c_int, [vp, C.c_int, C.c_int, dbl, dbl, dbl, dbl]),
        "vqe_plan_rotations": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int,
                                         vp, vp, vp, vp, vp]),
        "vqe_plan_paulisum": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp,
                                        C.c_int, vp, vp, vp]),
        "vqe_debug_tma_check": (C.c_int, [C.c_int, u64, C.c_int, C.c_int, C.c_int, P(i32), P(i32)]),
        "vqe_debug_tma_check_rl": (C.c_int, [C.c_int, u64, C.c_int, C.c_int, C.c_int, P(i32), P(i32)]),
        "vqe_peer_bytes": (C.c_int, [vp, P(u64), C.c_int]),
        "vqe_relabel_stats": (C.c_int, [vp, P(u64), P(u64), C.c_int]),
        "vqe_debug_coltab_host": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, P(i32), P(i32)]),
        "vqe_debug_coltab_banks": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
        "vqe_debug_diag2_host": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, P(dbl), P(i32)]),
        "vqe_debug_lean_host": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                          P(dbl), P(i32), P(i32), vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().vqe_last_error()
        raise VQEError("vqe_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
