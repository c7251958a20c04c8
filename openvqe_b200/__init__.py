"""B200-native VQE energy-and-gradient engine behind the Python entry points of
OpenVQE's ``openvqe.ucc_family`` and ``openvqe.adapt``.

Layout (only what the hot path needs):
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/vqe_b200.h)
  _lib.py          ctypes binding of the C ABI (fails loudly when the library is missing)
  lowering.py      qat-style Pauli operators -> packed X/Z bit masks
  engine.py        thin object wrapper over the C ABI (one context = one state vector)
  ucc_family/      mirrors openvqe.ucc_family (EnergyUCC for UCC and QUCCSD)
  adapt/           mirrors openvqe.adapt (fermionic_adapt_vqe, qubit_adapt_vqe)
  common_files/    host helpers with reference-identical semantics (sorted_gradient, circuit)
  sharded.py       multi-GPU layer: states sharded over 2/4/8 GPUs (one process per GPU, CUDA IPC + device-side
                   barriers), opt-in SPMD replica mode below that

There is NO CPU fallback: every numeric entry point needs the CUDA library and a GPU.
"""
__version__ = "0.1.0"
