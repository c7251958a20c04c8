"""Sharded state vectors: 34-36 qubits over 2/4/8 B200, and the host plumbing for the
replica regime below that.

The top ``n_global`` index bits -- reference qubits 0 .. n_global-1, myQLM's qubit 0 being the most significant
bit -- are the rank.  All arithmetic and all cross-GPU data movement happens in the CUDA library
(``include/vqe_b200.h``, "Sharded state"): operations whose X-mask flips a global bit run as *peer passes*, one
kernel that stages the same tile of ranks r and r^m in shared memory through peer memory over NVLink.  This
module only

* exchanges the CUDA IPC handles of the shards once (``torch.distributed`` object all-gather),
* adds the per-rank partial sums of reductions in rank order (fixed order -> the same bits on every rank),
* and, for states that fit one GPU, splits the ADAPT pool sweep over the ranks (``split_range``).

Two front ends:

``ShardedEngine``  one rank, one process per GPU (``torchrun``); same methods as ``Engine``, every rank issues
                   the same calls (SPMD) and gets the same energies back.
``ShardGroup``     all ranks driven by one process -- used by the 1-GPU tests (several "virtual ranks" on one
                   device exercise exactly the kernels and the planner of the multi-GPU path) and usable for a
                   single-process multi-GPU run.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import engine as engine_mod
from .engine import BUF_PSI, BUF_SIGMA, Engine, PauliSum, _ptr
from .lowering import PackedTerms, pack_operator


# ---- host-side collectives (plumbing) ----------------------------------------------------------------
def _dist():
    import torch.distributed as dist
    return dist


def dist_ready() -> bool:
    try:
        dist = _dist()
        return dist.is_available() and dist.is_initialized()
    except Exception:
        return False


def allgather_f64(values, group=None) -> np.ndarray:
    """[world, len(values)] array of every rank's ``values`` (float64), identical on all ranks."""
    import torch
    dist = _dist()
    v = np.ascontiguousarray(values, dtype=np.float64).reshape(-1)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if "nccl" in backend else torch.device("cpu")
    mine = torch.from_numpy(v.copy()).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return torch.stack(parts).cpu().numpy().reshape(world, -1)


def sum_in_rank_order(rows: np.ndarray) -> np.ndarray:
    """Fixed-order sum over the rank axis: every rank computes bit-identical totals."""
    acc = np.zeros(rows.shape[1], dtype=np.float64)
    for r in range(rows.shape[0]):
        acc = acc + rows[r]
    return acc


def split_range(n_items: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of ``n_items`` owned by ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def n_global_for(world: int) -> int:
    g = int(world).bit_length() - 1
    if world < 1 or (1 << g) != world:
        raise ValueError("the number of ranks must be a power of two, got %d" % world)
    return g


# ---- replica mode (states that fit one GPU): OPT-IN ---------------------------------------------------------
_REPLICA = {"on": False, "group": None}


def enable_replica(group=None):
    """Opt in to SPMD replica mode: every rank of ``group`` (default: the world group) runs the SAME problem in
    lockstep; the ADAPT pool sweep and the finite-difference evaluations of every BFGS gradient are then split over
    the ranks.  Off by default: ranks of a torchrun job often work on different problems."""
    _REPLICA["on"], _REPLICA["group"] = True, group


def disable_replica():
    _REPLICA["on"], _REPLICA["group"] = False, None


def replica_enabled() -> bool:
    return bool(_REPLICA["on"])


def replica_group():
    return _REPLICA["group"]


def ranks_agree(payload: bytes, group=None) -> bool:
    """True when every rank of the group passed the same bytes (CRC-32 + length all-gathered): the guard of every
    replica-mode split -- ranks that turn out to work on different inputs fall back to the serial path together."""
    import zlib
    rows = allgather_f64([float(zlib.crc32(payload)), float(len(payload))], group)
    return bool(np.all(rows == rows[0]))


def replica_pool_overlaps(engine, pool: PackedTerms, bra=BUF_SIGMA, ket=BUF_PSI, group=None):
    """Pool sweep of a state that fits one GPU, split over the ranks (n <= 33): every rank
    holds the same psi and sigma (SPMD: same calls, deterministic kernels), evaluates a contiguous slice of the
    pool and the slices are all-gathered.  No state traffic."""
    dist = _dist()
    if group is None:
        group = replica_group()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_ops = int(pool.offsets.shape[0]) - 1
    sig = b"".join(np.ascontiguousarray(a).tobytes() for a in (pool.offsets, pool.x, pool.z, pool.cre, pool.cim))
    if not ranks_agree(sig, group):  # different pools on different ranks: not SPMD, every rank sweeps its own pool
        return engine.pool_overlaps(pool, bra=bra, ket=ket)
    lo, hi = split_range(n_ops, world, rank)
    width = max(split_range(n_ops, world, r)[1] - split_range(n_ops, world, r)[0] for r in range(world))
    mine = np.zeros(2 * width, dtype=np.float64)
    if hi > lo:
        part = engine.pool_overlaps(slice_packed(pool, lo, hi), bra=bra, ket=ket)
        mine[:2 * (hi - lo)] = np.ascontiguousarray(part).view(np.float64)
    rows = allgather_f64(mine, group)
    out = np.zeros(n_ops, dtype=np.complex128)
    for r in range(world):
        a, b = split_range(n_ops, world, r)
        out[a:b] = rows[r, :2 * (b - a)].view(np.complex128)
    return out


def slice_packed(p: PackedTerms, lo: int, hi: int) -> PackedTerms:
    """Operators [lo, hi) of a packed pool."""
    a, b = int(p.offsets[lo]), int(p.offsets[hi])
    return PackedTerms(p.n, p.x[a:b], p.z[a:b], p.ny[a:b], p.cre[a:b], p.cim[a:b], p.offsets[lo:hi + 1] - a)


# ---- exact generator exponential on a sharded state -------------------------------------------------------
def taylor_exp(engine, packed: PackedTerms, theta: float):
    """psi <- exp(theta A) psi for an anti-Hermitian Pauli sum A on a SHARDED state (reference prepare_adapt_state,
    adapt/fermionic_adapt_vqe.py:12-38, scipy expm_multiply).  Commuting strings are rotations; otherwise the scaled
    Taylor series of vqe_apply_exp_paulisum is driven from the host, because its convergence test needs the norm of the
    current term summed over all ranks:  psi <- (sum_m (theta A / s)^m / m!)^s psi  with  term <- (theta / (s m)) A term.
    ``engine``: a ShardedEngine (every rank runs this in lockstep) or a ShardGroup; three state vectors per rank."""
    import math
    from ._hotpath import _strings_commute
    keep = (packed.cre != 0) | (packed.cim != 0)
    x, z, ny, cim = packed.x[keep], packed.z[keep], packed.ny[keep], packed.cim[keep]
    if len(x) == 0 or theta == 0.0:
        return
    if np.any(packed.cre[keep] != 0.0):
        raise _lib.VQEError("the generator is not anti-Hermitian (real coefficient parts)")
    if _strings_commute(x, z):
        engine.apply_rotations(x, z, ny, -float(theta) * cim)
        return
    from .engine import BUF_WORK
    ps = engine.paulisum(PackedTerms(packed.n, x, z, ny, np.zeros_like(cim), cim))
    scale = max(1, int(math.ceil(abs(theta) * float(np.abs(cim).sum()) / 0.5)))
    for _ in range(scale):
        engine.copy_buffer(BUF_SIGMA, BUF_PSI)
        for m in range(1, 41):
            engine.apply_paulisum(ps, dst=BUF_WORK, src=BUF_SIGMA)                 # next = A term
            engine.axpby(BUF_SIGMA, BUF_WORK, float(theta) / (scale * m), 0.0)    # term = (theta / (s m)) next
            engine.axpby(BUF_PSI, BUF_SIGMA, 1.0, 1.0)                            # psi += term
            if engine.norm2(BUF_SIGMA) < 1e-34:
                break


# ---- one rank per process ------------------------------------------------------------------------------
class ShardedEngine(Engine):
    """One rank of a sharded state (one process per GPU).  Every rank must issue the same sequence of calls."""

    def __init__(self, n_qubits: int, device: int, group=None, attach_scratch: bool = False):
        dist = _dist()
        if not dist_ready():
            raise _lib.VQEError("ShardedEngine needs an initialised torch.distributed process group")
        self.group = group
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        super().__init__(n_qubits, device, n_global=n_global_for(world), rank=rank)
        self.world = world
        from .engine import BUF_WORK
        # sigma (ADAPT sweeps) when two shards fit one GPU; the work vector too (Taylor exponential of a non-commuting
        # generator: three vectors) when three do -- 2^31 amplitudes per rank, i.e. 34 qubits on 8 GPUs
        self._attached = [BUF_PSI] + ([BUF_SIGMA] if attach_scratch else []) + ([BUF_WORK] if attach_scratch and self.n_local <= 31 else [])
        self._connect(self._attached)

    def _connect(self, bufs):
        """Export this rank's shard(s) and flag array as CUDA IPC handles, all-gather them, map the peers'."""
        dist = _dist()
        mine = {}
        for what in list(bufs) + [_lib.SHARD_FLAGS]:
            h = (C.c_ubyte * _lib.IPC_HANDLE_BYTES)()
            _lib.check(self._lib.vqe_shard_export(self.handle, what, h))
            mine[what] = bytes(h)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        for peer, handles in enumerate(everyone):
            if peer == self.rank:
                continue
            for what, raw in handles.items():
                buf = (C.c_ubyte * _lib.IPC_HANDLE_BYTES).from_buffer_copy(raw)
                _lib.check(self._lib.vqe_shard_attach_ipc(self.handle, peer, what, buf))
        dist.barrier(group=self.group)

    # -- state: the host side only ever sees this rank's shard ----------------------------------------
    def set_state(self, vec, buf=BUF_PSI):
        v = np.asarray(vec, dtype=np.complex128).reshape(-1)
        if v.shape[0] == 1 << self.n:  # a full vector: keep my slice
            v = v[self.rank << self.n_local:(self.rank + 1) << self.n_local]
        super().set_state(v, buf)

    def get_local_state(self, buf=BUF_PSI):
        """This rank's shard (2^n_local amplitudes)."""
        return super().get_state(buf)

    def get_state(self, buf=BUF_PSI):
        """The FULL state vector, all-gathered over the ranks (what the reference-shaped helpers expect from
        ``get_state``).  Refused above 30 qubits (16 GiB per host copy): use ``get_local_state``."""
        if self.n > 30:
            raise _lib.VQEError("sharded state of %d qubits: the full vector does not fit a host array; "
                                "use get_local_state() for this rank's shard" % self.n)
        mine = super().get_state(buf)
        rows = allgather_f64(mine.view(np.float64), self.group)
        return rows.reshape(-1).view(np.complex128)

    def _need_scratch(self, what):
        if BUF_SIGMA not in self._attached:
            raise _lib.VQEError("%s needs a second state buffer on every rank, which is not attached for this sharded state "
                                "(shards of 2^%d amplitudes: two of them do not fit one GPU); ADAPT sweeps are available on "
                                "sharded states with n_local <= 32" % (what, self.n_local))

    def apply_paulisum(self, ps, dst=BUF_SIGMA, src=BUF_PSI):
        self._need_scratch("sigma = H psi")
        return super().apply_paulisum(ps, dst, src)

    def apply_exp(self, packed: PackedTerms, theta: float):
        """exp(theta A) on the sharded state (commuting strings: rotations; else the host-driven Taylor series)."""
        from .engine import BUF_WORK
        if BUF_WORK not in self._attached:
            keep = (packed.cre != 0) | (packed.cim != 0)
            from ._hotpath import _strings_commute
            if not _strings_commute(packed.x[keep], packed.z[keep]):
                raise _lib.VQEError("exact exponential of a non-commuting generator needs three state vectors per rank, "
                                    "which do not fit for shards of 2^%d amplitudes" % self.n_local)
        taylor_exp(self, packed, theta)

    def barrier(self):
        _lib.check(self._lib.vqe_shard_barrier(self.handle))

    # -- reductions: local partial -> all-gather -> fixed-order sum -------------------------------------
    def _total(self, parts):
        return sum_in_rank_order(allgather_f64(parts, self.group))

    def expectation(self, ps: PauliSum, buf=BUF_PSI) -> complex:
        part = super().expectation(ps, buf)
        t = self._total([part.real, part.imag])
        return complex(t[0], t[1])

    def pool_overlaps(self, pool: PackedTerms, bra=BUF_SIGMA, ket=BUF_PSI):
        if BUF_SIGMA in (bra, ket):
            self._need_scratch("the ADAPT pool sweep")
        part = super().pool_overlaps(pool, bra, ket)
        return self._total(part.view(np.float64)).view(np.complex128)

    def norm2(self, buf=BUF_PSI) -> float:
        return float(self._total([super().norm2(buf)])[0])

    def inner(self, a, b) -> complex:
        part = super().inner(a, b)
        t = self._total([part.real, part.imag])
        return complex(t[0], t[1])

    def overlap_host(self, vec, buf=BUF_PSI) -> complex:
        v = np.asarray(vec, dtype=np.complex128).reshape(-1)
        if v.shape[0] == 1 << self.n:
            v = v[self.rank << self.n_local:(self.rank + 1) << self.n_local]
        part = super().overlap_host(v, buf)
        t = self._total([part.real, part.imag])
        return complex(t[0], t[1])


def enable(min_qubits: int = 34, group=None):
    """Make ``get_engine`` (and with it every reference-shaped entry point of ``openvqe_b200.ucc_family`` /
    ``openvqe_b200.adapt``) return a ``ShardedEngine`` for states of ``min_qubits`` or more.  Call once per
    process after ``torch.distributed.init_process_group``."""
    def factory(n_qubits, device):
        if n_qubits >= min_qubits and dist_ready() and _dist().get_world_size(group) > 1:
            world = _dist().get_world_size(group)
            # sigma is attached whenever two shards fit one GPU (2 x 2^32 x 16 B = 137 GB of 180 GB)
            return ShardedEngine(n_qubits, device, group=group, attach_scratch=n_qubits - n_global_for(world) <= 32)
        return None
    engine_mod._ENGINE_FACTORY = factory


def disable():
    engine_mod._ENGINE_FACTORY = None


# ---- all ranks in one process -------------------------------------------------------------------------
class GroupPauliSum:
    def __init__(self, group, packed: PackedTerms):
        self.per_rank = [PauliSum(e, packed) for e in group.ranks]
        self.handles = (C.c_void_p * len(self.per_rank))(*[p.handle for p in self.per_rank])
        self.n_groups = self.per_rank[0].n_groups
        self.n_passes = self.per_rank[0].n_passes


class ShardGroup:
    """All 2^n_global ranks of a sharded state, driven by one host thread.  ``devices`` may repeat a device
    (virtual ranks on one GPU)."""

    def __init__(self, n_qubits: int, n_global: int, devices=None):
        self.n, self.n_global = int(n_qubits), int(n_global)
        self.world = 1 << self.n_global
        self.n_local = self.n - self.n_global
        devices = list(devices) if devices is not None else [0] * self.world
        if len(devices) != self.world:
            raise ValueError("need %d devices, got %d" % (self.world, len(devices)))
        self._lib = _lib.load()
        self.ranks = [Engine(self.n, devices[r], n_global=self.n_global, rank=r) for r in range(self.world)]
        for a in self.ranks:
            for b in self.ranks:
                if a is not b:
                    _lib.check(self._lib.vqe_shard_attach_local(a.handle, b.handle))
        self._handles = (C.c_void_p * self.world)(*[e.handle for e in self.ranks])

    # -- state ---------------------------------------------------------------------------------------------
    def set_basis_state(self, index: int):
        for e in self.ranks:
            e.set_basis_state(index)

    def set_state(self, vec, buf=BUF_PSI):
        v = np.asarray(vec, dtype=np.complex128).reshape(-1)
        if v.shape[0] != 1 << self.n:
            raise ValueError("state has %d amplitudes, expected 2^%d" % (v.shape[0], self.n))
        for r, e in enumerate(self.ranks):
            e.set_state(v[r << self.n_local:(r + 1) << self.n_local], buf)

    def get_state(self, buf=BUF_PSI):
        return np.concatenate([e.get_state(buf) for e in self.ranks])

    def synchronize(self):
        for e in self.ranks:
            e.synchronize()

    def copy_buffer(self, dst, src):
        for e in self.ranks:
            e.copy_buffer(dst, src)

    def axpby(self, dst, x, alpha=1.0, beta=1.0):
        for e in self.ranks:
            e.axpby(dst, x, alpha, beta)

    def apply_exp(self, packed: PackedTerms, theta: float):
        taylor_exp(self, packed, theta)

    # -- operations ----------------------------------------------------------------------------------------
    def apply_rotations(self, x, z, ny, angles):
        x = np.ascontiguousarray(x, dtype=np.uint64)
        z = np.ascontiguousarray(z, dtype=np.uint64)
        ny = np.ascontiguousarray(ny, dtype=np.int32)
        a = np.ascontiguousarray(angles, dtype=np.float64)
        _lib.check(self._lib.vqe_group_apply_pauli_rotations(self._handles, self.world, int(x.shape[0]), _ptr(x),
                                                             _ptr(z), _ptr(ny), _ptr(a)))

    def apply_gates(self, kinds, q0, q1, angles):
        k = np.ascontiguousarray(kinds, dtype=np.int32)
        a0 = np.ascontiguousarray(q0, dtype=np.int32)
        a1 = np.ascontiguousarray(q1, dtype=np.int32)
        an = np.ascontiguousarray(angles, dtype=np.float64)
        _lib.check(self._lib.vqe_group_apply_gates(self._handles, self.world, int(k.shape[0]), _ptr(k), _ptr(a0),
                                                   _ptr(a1), _ptr(an)))

    def paulisum(self, operator) -> GroupPauliSum:
        packed = operator if isinstance(operator, PackedTerms) else pack_operator(operator, with_constant=True)
        return GroupPauliSum(self, packed)

    def expectation(self, ps: GroupPauliSum, buf=BUF_PSI) -> complex:
        out = (C.c_double * 2)()
        _lib.check(self._lib.vqe_group_expectation(self._handles, self.world, buf, ps.handles, out))
        return complex(out[0], out[1])

    def apply_paulisum(self, ps: GroupPauliSum, dst=BUF_SIGMA, src=BUF_PSI):
        _lib.check(self._lib.vqe_group_apply_paulisum(self._handles, self.world, dst, src, ps.handles))

    def pool_overlaps(self, pool: PackedTerms, bra=BUF_SIGMA, ket=BUF_PSI):
        n_ops = int(pool.offsets.shape[0]) - 1
        out = np.zeros(n_ops, dtype=np.complex128)
        _lib.check(self._lib.vqe_group_pool_overlaps(self._handles, self.world, bra, ket, n_ops, _ptr(pool.offsets),
                                                     _ptr(pool.x), _ptr(pool.z), _ptr(pool.ny), _ptr(pool.cre),
                                                     _ptr(pool.cim), _ptr(out)))
        return out

    def norm2(self, buf=BUF_PSI) -> float:
        return float(sum(e.norm2(buf) for e in self.ranks))

    @property
    def launch_count(self) -> int:
        return sum(e.launch_count for e in self.ranks)


def plan_rotations(n_qubits, n_global, x, z, ny, angles, tile_bits=12, low_bits=5):
    """Host-only view of the pass planner: list of (kind, pattern, n_ops, tile_mask) per pass, kind 0 = local,
    1 = peer pass between ranks r and r ^ pattern in exchange form, 2 = the same in gather form.  No GPU needed."""
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.uint64)
    z = np.ascontiguousarray(z, dtype=np.uint64)
    ny = np.ascontiguousarray(ny, dtype=np.int32)
    a = np.ascontiguousarray(angles, dtype=np.float64)
    cap = max(1, int(x.shape[0]))
    n_passes = C.c_int32()
    kind = np.zeros(cap, dtype=np.int32)
    pat = np.zeros(cap, dtype=np.uint64)
    nops = np.zeros(cap, dtype=np.int32)
    tmask = np.zeros(cap, dtype=np.uint64)
    _lib.check(lib.vqe_plan_rotations(int(n_qubits), int(n_global), int(tile_bits), int(low_bits), int(x.shape[0]),
                                      _ptr(x), _ptr(z), _ptr(ny), _ptr(a), cap, C.byref(n_passes), _ptr(kind),
                                      _ptr(pat), _ptr(nops), _ptr(tmask)))
    k = n_passes.value
    return [(int(kind[i]), int(pat[i]), int(nops[i]), int(tmask[i])) for i in range(k)]
