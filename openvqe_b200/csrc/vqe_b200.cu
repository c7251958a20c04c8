// vqe_b200.cu -- B200-native (sm_100a) state-vector engine for the OpenVQE hot path.
//
// Everything here is streaming complex128 work on a state vector (0.2 flop/byte at face value): no tensor cores.
// Design:
//   * The state lives in HBM as interleaved complex128.  Every kernel works on TILES: a tile is the set of 2^T
//     amplitudes obtained by fixing all index bits outside a chosen set of T "tile bits".  The low L tile bits
//     are always index bits 0..L-1, so a tile is a gather of 2^(T-L) contiguous 16*2^L-byte segments (coalesced
//     16-byte accesses, 512 B per warp request, cp.async into shared memory).
//   * A CTA stages one tile in shared memory, applies EVERY consecutive operation whose X-mask lies inside the
//     tile bits (Z-masks may touch any bit: bits outside the tile only contribute a per-tile sign), and writes the
//     tile back: r rotations per HBM pass instead of one.
//   * Rotations with the same X-mask act in the same planes and commute: a run of them is COLLAPSED into one plane
//     rotation whose angle is tabulated per occupation pattern (the 8 strings of a JW double excitation touch 1/8
//     of the pairs, once); what does not collapse is applied on 8-amplitude ORBITS held in registers (up to three
//     generators per shared-memory round trip).  The QUCCSD gate templates are tabulated plane rotations too.
//   * <psi|H|psi>, H|psi> and the ADAPT pool sweep use the same tiles: H is grouped by X-mask, the groups are packed
//     into passes by covering their X-masks with tile-bit sets, a group's Z-variants are tabulated per occupation
//     pattern and only the coupled patterns are visited.  Reductions are fp64 warp-shuffle + block + fixed-order
//     final pass (bit-reproducible run to run).
//   * Sharded states: the top index bits are the rank; operations that flip a global bit run as PEER PASSES -- the
//     same kernels on tiles with one virtual bit that selects between this rank's HBM and the partner's shard
//     (peer memory over NVLink), ordered by a device-side flag barrier.
//   * Persistent grids sized from the SM count; one stream per context.
//
// No CPU fallback: every entry point that computes needs a CUDA device.
#include <cuda.h>  // CUtensorMap (the encoder entry point is fetched at run time, libcuda is not linked)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdarg.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <string>
#include <map>
#include <string>
#include <vector>

#include "vqe_b200.h"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(VQE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),      \
                        __FILE__, __LINE__);                                                       \
    } while (0)

extern "C" const char* vqe_last_error(void) { return g_err.c_str(); }
extern "C" int vqe_version(void) { return 100; }
extern "C" int vqe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

// ------------------------------------------------------------------------------------------
// device structures
// ------------------------------------------------------------------------------------------
enum { OP_ROT = 0, OP_GATE1 = 1, OP_CNOT = 2, OP_ROTF = 3, OP_PLANE = 4 };

struct DevOp {       // 64 bytes
    uint32_t lx;     // ROT: local X mask | GATE1: local bit mask | CNOT: local target mask
    uint32_t lz;     // ROT: local Z mask | CNOT: local control mask (0 if control is outside the tile)
    uint64_t zout;   // ROT: Z mask outside the tile | CNOT: control mask outside the tile
    double c, s;     // ROT: cos, sin
    uint32_t k4;     // ROT: (ny + 3) & 3, the unit phase (-i) * i^ny = i^k4
    uint32_t jmask;  // ROTF: bit j = parity(u_j & lz), u_j = index bits contributed by the j-th pair of a thread
    uint32_t kind;   // OP_*
    uint32_t hb;     // highest set bit of lx
    uint32_t run;    // ROT: number of consecutive ROT ops (starting here) sharing lx
    uint32_t mat;    // GATE1: index into the matrix array (8 doubles each)
    uint32_t nyodd;  // ROT: ny & 1
    uint32_t imag;   // ROTF: 1 when the unit phase is +-i (ny even)
};

// Tensor-map form of a tile (TMA, cp.async.bulk.tensor).  The shard is described to the TMA unit as a <= 5-dimensional
// array of doubles whose dimensions start at the runs of adjacent tile bits: dimension i covers the index bits
// [shift[i], shift[i+1]), its box is the run of tile bits at its start.  One request then fetches the whole tile (or, when
// the tile bits form more than five runs, the part selected by the remaining high tile bits: n_req <= 16 requests)
// instead of one bulk copy per 512-byte segment -- the TMA unit serves a request in ~40 cycles whatever its size, so
// per-segment copies cap a tile pass at about half of the HBM bandwidth.
struct TmaGeom {
    uint32_t shift[5];     // first index bit of dimension i
    uint32_t cmask[5];     // (1 << number of index bits covered by dimension i) - 1
    uint32_t n_req;        // requests per tile
    uint32_t req_amps;     // amplitudes per request
    uint64_t req_bits[16]; // index bits of request r (its pattern on the tile bits beyond the fifth run)
};

struct TileGeom {
    uint64_t comp_mask;  // (shard-)local index bits NOT in the tile
    uint64_t n_tiles;    // number of tiles THIS launch walks
    const uint64_t* scat;  // 2^(T-L) entries: deposit of the high local bits into index space
    uint32_t tbits, lbits;
    // --- sharded state (see "sharding" below); all zero / identity for a single-GPU context ---
    uint64_t sign_base;    // global index bits of the (lower) rank: only ever used in Z parities
    uint64_t tile_first;   // tile numbers walked: tile_first + k * tile_stride, k < n_tiles
    uint32_t tile_stride;
    uint32_t vbit;         // 1: the top tile bit is VIRTUAL -- it selects the shard (0: p0 = lower rank, 1: p1 = r ^ m)
    uint32_t bulk;         // 1: stage the tile with bulk asynchronous copies (cp.async.bulk, one per contiguous segment)
    uint32_t tma;          // 1: stage the tile with tensor-map requests (tg + the CUtensorMap kernel parameter)
    uint32_t swz;          // 0x70 when the tile sits in shared memory in the TMA 128-byte swizzle (the 16-byte chunk index of
                           // an element, byte-offset bits 4-6, is XORed with bits 7-9): whatever tile bits an operation
                           // fixes, a warp's 32 elements then spread over all eight chunk positions (no bank conflicts
                           // beyond the 16-byte stride itself).  0: natural layout.
    uint32_t rl;           // 1: REAL LAYOUT -- the buffer holds the n_amp real parts as contiguous doubles (see vqe_ctx::real_layout);
                           // tile elements are 8 bytes, the tensor map counts one double per amplitude
    uint32_t rev;          // 1: the tiles are walked downwards (tile_first - k * tile_stride): consecutive passes over a state
                           // of about the size of the L2 alternate, so a pass starts with what the previous one touched last
    TmaGeom tg;
};
// byte offset of a tile element in the (possibly swizzled) shared-memory tile; linear over XOR
__host__ __device__ __forceinline__ uint32_t swz_off(uint32_t off, uint32_t swz) { return off ^ ((off >> 3) & swz); }
// the same on element indices
__host__ __device__ __forceinline__ uint32_t swz_idx(uint32_t l, uint32_t swz) { return l ^ ((l >> 3) & (swz >> 4)); }
// ... and on indices of 8-byte elements (real layout): byte-offset bits 4-6 are element-index bits 1-3, bits 7-9 are bits 4-6
__host__ __device__ __forceinline__ uint32_t swz_idx8(uint32_t l, uint32_t swz) { return l ^ ((l >> 3) & (swz >> 3)); }

// Peer pass, "gather" form.  When the operations of a peer pass couple only a fraction of the partner's amplitudes
// (collapsed runs: 1/8 of them for a JW double excitation), moving whole half-tiles over NVLink is wasteful.
// Instead every rank keeps ALL of its own tiles, first gathers from the partner's shard just the amplitudes its own
// results depend on (host-computed dependency closure, in groups of 4 = 64 bytes) into a local staging buffer
// (k_gather_need, remote READS only), and after a barrier runs the pass on super-tiles whose partner half is filled
// from the staging buffer and whose own half alone is written back (local HBM only).  Both ranks of a pair compute
// the coupled pairs redundantly; nothing is written remotely.
struct GatherGeom {
    const double2* stage;   // staging buffer: [tile of the chunk][need group][4]
    const uint16_t* need;   // needed groups of the partner half (tile-local index / 4), ascending
    uint32_t n_need;        // 0: not a gather launch
    uint32_t own_half;      // 0: this rank's amplitudes are the lower half of the super-tile, 1: the upper half
};

// The two shards a launch touches.  Local passes: p0 = own shard, p1 unused.  Peer passes (vbit): p0 = shard of
// min(r, r^m), p1 = shard of max(r, r^m); one of them is this rank's HBM, the other a peer mapping over NVLink.
struct Shards {
    double2* p0;
    double2* p1;
};

struct DevGroup {     // 40 bytes: one X-mask group of a Pauli sum inside a pass
    uint32_t lx, hb;
    uint32_t t_begin;  // first term (index into the pass-local term list)
    uint32_t n_even;   // terms with even ny come first, then n_odd terms with odd ny
    uint32_t n_odd;
    uint32_t pad;
    // k_tile_expect: inside each parity block the terms are sorted by sign class (see the kernel);
    // lx != 0: cnt[0..3] = even-ny classes, cnt[4..7] = odd-ny classes; lx == 0: cnt[0..7] = the 8 classes
    uint16_t cnt[8];
};
struct DevTerm {      // 32 bytes
    uint64_t zout;
    uint32_t lz;
    uint32_t jmask;   // expectation: bit j = parity(u_j & lz) for the j-th pair (amplitude) of a thread
    double ar, ai;    // expectation: pair weight; apply: c_k * i^ny
};

__device__ __forceinline__ uint64_t pdep64(uint64_t v, uint64_t mask) {
    uint64_t out = 0;
    while (mask) {
        uint64_t low = mask & (0 - mask);
        if (v & 1) out |= low;
        v >>= 1;
        mask ^= low;
    }
    return out;
}

__device__ __forceinline__ uint32_t insert0(uint32_t p, uint32_t bit) {
    return ((p >> bit) << (bit + 1)) | (p & ((1u << bit) - 1u));
}

__device__ __forceinline__ double2 ld_amp(const double2* p) { return *p; }

__device__ __forceinline__ double flipsign(double v, uint32_t par) {
    return __longlong_as_double(__double_as_longlong(v) ^ ((long long)(par & 1u) << 63));
}

// ------------------------------------------------------------------------------------------
// simple kernels
// ------------------------------------------------------------------------------------------
__global__ void k_zero_set(double2* psi, uint64_t n_amp, uint64_t index) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_amp; i += stride) psi[i] = make_double2(i == index ? 1.0 : 0.0, 0.0);
}

__global__ void k_zero_set_real(double* psi, uint64_t n_amp, uint64_t index) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_amp; i += stride) psi[i] = i == index ? 1.0 : 0.0;
}
// real layout -> interleaved complex, in place, one level: the reals [lo, hi) (hi <= 2 lo, or lo = 0 and hi = 1) become the
// complex amplitudes [lo, hi), i.e. the doubles [2 lo, 2 hi), which do not overlap the reals still to be expanded
__global__ void k_expand_level(double* buf, uint64_t lo, uint64_t hi) {
    uint64_t i = lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < hi; i += stride) {
        const double v = buf[i];
        reinterpret_cast<double2*>(buf)[i] = make_double2(v, 0.0);
    }
}

// dst = alpha * x + beta * dst   (complex alpha, beta given as re/im)
__global__ void k_axpby(double2* dst, const double2* x, uint64_t n_amp, double are, double aim, double bre,
                        double bim) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_amp; i += stride) {
        double2 d = dst[i], v = x[i];
        double2 o;
        o.x = are * v.x - aim * v.y + bre * d.x - bim * d.y;
        o.y = are * v.y + aim * v.x + bre * d.y + bim * d.x;
        dst[i] = o;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of (re, im); result valid in thread 0.  red must hold 2*32 doubles.
__device__ __forceinline__ double2 block_sum2(double re, double im, double* red) {
    re = warp_sum(re);
    im = warp_sum(im);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) {
        red[2 * w] = re;
        red[2 * w + 1] = im;
    }
    __syncthreads();
    double2 out = make_double2(0.0, 0.0);
    if (threadIdx.x == 0) {
        int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) {  // fixed order
            out.x += red[2 * i];
            out.y += red[2 * i + 1];
        }
    }
    return out;
}

// partial[b] = sum_i conj(a[i]) * b[i] over this block's grid-stride share
__global__ void k_inner(const double2* a, const double2* b, uint64_t n_amp, double2* partial) {
    __shared__ double red[64];
    double re = 0.0, im = 0.0;
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_amp; i += stride) {
        double2 u = a[i], v = b[i];
        re += u.x * v.x + u.y * v.y;
        im += u.x * v.y - u.y * v.x;
    }
    double2 s = block_sum2(re, im, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[j] = sum_b partial[b * stride + j].  One CTA per output j; thread t adds blocks t, t+T, t+2T, ... in
// that fixed order, then a fixed-shape shuffle/shared-memory tree combines the T partial sums, so the result
// is bit-reproducible run to run.
__global__ void k_reduce_partials(const double2* partial, int n_blocks, int stride, int n_out, double2* out) {
    __shared__ double red[64];
    const int j = blockIdx.x;
    if (j >= n_out) return;
    double re = 0.0, im = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += blockDim.x) {
        double2 p = partial[(size_t)b * stride + j];
        re += p.x;
        im += p.y;
    }
    double2 s = block_sum2(re, im, red);
    if (threadIdx.x == 0) out[j] = s;
}

// ------------------------------------------------------------------------------------------
// tile load / store
// ------------------------------------------------------------------------------------------
#define LOAD_BATCH 8
// address of tile element k (tile-local index) of the tile whose fixed bits are `base`
__device__ __forceinline__ double2* amp_addr(const TileGeom& g, const Shards& sh, uint64_t base, uint32_t k) {
    const uint32_t lmask = (1u << g.lbits) - 1u;
    if (g.vbit) {
        const uint32_t top = g.tbits - 1u;
        const uint32_t kl = k & ((1u << top) - 1u);
        double2* p = (k >> top) ? sh.p1 : sh.p0;  // warp-uniform (top >= lbits >= 5 whenever a tile has >= 64 elements)
        return p + (base | __ldg(g.scat + (kl >> g.lbits)) | (uint64_t)(kl & lmask));
    }
    return sh.p0 + (base | __ldg(g.scat + (k >> g.lbits)) | (uint64_t)(k & lmask));
}
__device__ __forceinline__ uint64_t tile_base(const TileGeom& g, uint64_t t) {
    return pdep64(g.rev ? g.tile_first - t * g.tile_stride : g.tile_first + t * g.tile_stride, g.comp_mask);
}

__device__ __forceinline__ void tile_load(double2* tile, const Shards& src, const TileGeom& g, uint64_t base) {
    const uint32_t ts = 1u << g.tbits;
    // LOAD_BATCH independent 16-byte loads in flight per thread before the first shared-memory store
    for (uint32_t k0 = threadIdx.x; k0 < ts; k0 += blockDim.x * LOAD_BATCH) {
        double2 v[LOAD_BATCH];
#pragma unroll
        for (int j = 0; j < LOAD_BATCH; ++j) {
            const uint32_t k = k0 + j * blockDim.x;
            if (k < ts) v[j] = ld_amp(amp_addr(g, src, base, k));
        }
#pragma unroll
        for (int j = 0; j < LOAD_BATCH; ++j) {
            const uint32_t k = k0 + j * blockDim.x;
            if (k < ts) tile[k] = v[j];
        }
    }
}
__device__ __forceinline__ void tile_store(const double2* tile, const Shards& dst, const TileGeom& g, uint64_t base) {
    const uint32_t ts = 1u << g.tbits;
#pragma unroll 4
    for (uint32_t k = threadIdx.x; k < ts; k += blockDim.x) *amp_addr(g, dst, base, k) = tile[k];
}

// ------------------------------------------------------------------------------------------
// state preparation: fused rotations / gates on a shared-memory tile
// ------------------------------------------------------------------------------------------
#define ROT_PAIRS 4   // fast path: pairs per thread held in registers (512 threads x 4 = half a 2^12 tile)
#define SLOW_PAIRS 4  // general path
#define OPTAB_CAP 512

struct FastOp {  // 16 bytes, shared-memory copy of a fast rotation, refreshed per tile
    double t;       // tan(angle) * (unit-phase sign) * (-1)^popc(base & zout)
    uint32_t lz;
    uint32_t meta;  // jmask | imag << 31
};

__device__ __forceinline__ void rot_update(double2& a, double2& b, const DevOp& op, uint32_t pa) {
    // P psi at l  = i^ny (-1)^pb b ;  at l2 = i^ny (-1)^pa a ;  new = c*old - i s (P psi)
    //             = c*old + s (-1)^p i^k4 other,  k4 = (ny + 3) & 3
    const uint32_t neg = op.k4 >> 1;
    const double sb = flipsign(op.s, pa ^ op.nyodd ^ neg), sa = flipsign(op.s, pa ^ neg);
    double tbr, tbi, tar, tai;
    if (op.k4 & 1u) {  // multiply by i
        tbr = -b.y; tbi = b.x; tar = -a.y; tai = a.x;
    } else {
        tbr = b.x; tbi = b.y; tar = a.x; tai = a.y;
    }
    double2 na, nb;
    na.x = op.c * a.x + sb * tbr;
    na.y = op.c * a.y + sb * tbi;
    nb.x = op.c * b.x + sa * tar;
    nb.y = op.c * b.y + sa * tai;
    a = na;
    b = nb;
}


// Slow, fully general paths (large-angle / diagonal rotations, one-qubit gates, CNOT).  Kept out of line so
// that their register needs do not inflate the fast path of k_tile_ops.
__device__ __noinline__ void slow_rot(double2* tile, const DevOp* __restrict__ ops, int i, uint64_t base, uint32_t ts) {
    const uint32_t half = ts >> 1;
    const DevOp op = ops[i];
                const int run = (int)op.run;
                if (op.lx == 0) {
                    // diagonal run: psi[l] *= prod_r (c_r - i s_r (-1)^par_r)
                    for (uint32_t l0 = threadIdx.x; l0 < ts; l0 += blockDim.x * SLOW_PAIRS) {
                        double2 a[SLOW_PAIRS];
#pragma unroll
                        for (int j = 0; j < SLOW_PAIRS; ++j) {
                            const uint32_t l = l0 + j * blockDim.x;
                            if (l < ts) a[j] = tile[l];
                        }
                        for (int r = 0; r < run; ++r) {
                            const DevOp o2 = ops[i + r];
                            const uint32_t opar = __popcll(base & o2.zout);
#pragma unroll
                            for (int j = 0; j < SLOW_PAIRS; ++j) {
                                const uint32_t l = l0 + j * blockDim.x;
                                const double ss = flipsign(o2.s, __popc(l & o2.lz) + opar);
                                double2 na;
                                na.x = o2.c * a[j].x + ss * a[j].y;
                                na.y = o2.c * a[j].y - ss * a[j].x;
                                a[j] = na;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < SLOW_PAIRS; ++j) {
                            const uint32_t l = l0 + j * blockDim.x;
                            if (l < ts) tile[l] = a[j];
                        }
                    }
                } else {
                    // the pairs (l, l^lx) are invariant under every rotation of the run: keep them
                    // in registers and apply the whole run without touching shared memory again
                    for (uint32_t p0 = threadIdx.x; p0 < half; p0 += blockDim.x * SLOW_PAIRS) {
                        double2 a[SLOW_PAIRS], b[SLOW_PAIRS];
                        uint32_t li[SLOW_PAIRS];
#pragma unroll
                        for (int j = 0; j < SLOW_PAIRS; ++j) {
                            const uint32_t p = p0 + j * blockDim.x;
                            li[j] = insert0(p, op.hb);
                            if (p < half) {
                                a[j] = tile[li[j]];
                                b[j] = tile[li[j] ^ op.lx];
                            }
                        }
                        for (int r = 0; r < run; ++r) {
                            const DevOp o2 = ops[i + r];
                            const uint32_t opar = __popcll(base & o2.zout);
#pragma unroll
                            for (int j = 0; j < SLOW_PAIRS; ++j)
                                rot_update(a[j], b[j], o2, (__popc(li[j] & o2.lz) + opar) & 1u);
                        }
#pragma unroll
                        for (int j = 0; j < SLOW_PAIRS; ++j) {
                            const uint32_t p = p0 + j * blockDim.x;
                            if (p < half) {
                                tile[li[j]] = a[j];
                                tile[li[j] ^ op.lx] = b[j];
                            }
                        }
                    }
                }
}

__device__ __noinline__ void slow_gate1(double2* tile, const DevOp* __restrict__ ops, int i, const double* __restrict__ mats,
                                        uint32_t ts) {
    const uint32_t half = ts >> 1;
    const DevOp op = ops[i];
                const double* m = mats + (size_t)op.mat * 8;
                const double m00r = m[0], m00i = m[1], m01r = m[2], m01i = m[3];
                const double m10r = m[4], m10i = m[5], m11r = m[6], m11i = m[7];
                for (uint32_t p = threadIdx.x; p < half; p += blockDim.x) {
                    const uint32_t l = insert0(p, op.hb), l2 = l | op.lx;
                    double2 a = tile[l], b = tile[l2], na, nb;
                    na.x = m00r * a.x - m00i * a.y + m01r * b.x - m01i * b.y;
                    na.y = m00r * a.y + m00i * a.x + m01r * b.y + m01i * b.x;
                    nb.x = m10r * a.x - m10i * a.y + m11r * b.x - m11i * b.y;
                    nb.y = m10r * a.y + m10i * a.x + m11r * b.y + m11i * b.x;
                    tile[l] = na;
                    tile[l2] = nb;
                }
}

__device__ __noinline__ void slow_cnot(double2* tile, const DevOp* __restrict__ ops, int i, uint64_t base, uint32_t ts) {
    const uint32_t half = ts >> 1;
    const DevOp op = ops[i];
                const bool on = (op.zout == 0) || ((base & op.zout) != 0);
                if (on) {
                    for (uint32_t p = threadIdx.x; p < half; p += blockDim.x) {
                        const uint32_t l = insert0(p, op.hb), l2 = l | op.lx;
                        if (op.lz == 0 || (l & op.lz)) {
                            double2 a = tile[l];
                            tile[l] = tile[l2];
                            tile[l2] = a;
                        }
                    }
                }
}


// Fast rotation run in tangent form.  Every rotation of the run acts on the same pairs (l, l^lx):
//   R = c [[1, -+t], [+-t, 1]]  ->  apply the unnormalised updates with one FMA per component and multiply
// by prod(c) once at the end of the run.  parity(l & lz) splits into a per-thread part (one popcount per
// rotation) and a per-pair part that is the same for all threads (host-computed jmask), so a pair costs
// one LOP3 + four DFMA per rotation.  IMAG: unit phase +-i (ny even) instead of +-1 (ny odd).
template <bool IMAG>
__device__ __forceinline__ void rot_fast_run(double2* tile, const FastOp* __restrict__ tab, uint32_t lx, uint32_t hb,
                                             int run, double cscale, uint32_t half) {
    const uint32_t l0 = insert0(threadIdx.x, hb);
    for (uint32_t it = 0; it * blockDim.x * ROT_PAIRS < half; ++it) {
        double ax[ROT_PAIRS], ay[ROT_PAIRS], bx[ROT_PAIRS], by[ROT_PAIRS];
#pragma unroll
        for (int j = 0; j < ROT_PAIRS; ++j) {
            const uint32_t pidx = threadIdx.x + (it * ROT_PAIRS + j) * blockDim.x;
            const uint32_t lj = insert0(pidx, hb);
            ax[j] = ay[j] = bx[j] = by[j] = 0.0;
            if (pidx < half) {
                const double2 va = tile[lj], vb = tile[lj ^ lx];
                ax[j] = va.x; ay[j] = va.y; bx[j] = vb.x; by[j] = vb.y;
            }
        }
#pragma unroll 1
        for (int r = 0; r < run; ++r) {
            const FastOp f = tab[r];
            const uint32_t thi = (uint32_t)__double2hiint(f.t) ^ ((uint32_t)__popc(l0 & f.lz) << 31);
            const int tlo = __double2loint(f.t);
            const uint32_t jm = f.meta >> (it * ROT_PAIRS);
#pragma unroll
            for (int j = 0; j < ROT_PAIRS; ++j) {
                const double tj = __hiloint2double((int)(thi ^ ((jm << (31 - j)) & 0x80000000u)), tlo);
                if (IMAG) {
                    const double nax = fma(-tj, by[j], ax[j]), nay = fma(tj, bx[j], ay[j]);
                    const double nbx = fma(-tj, ay[j], bx[j]), nby = fma(tj, ax[j], by[j]);
                    ax[j] = nax; ay[j] = nay; bx[j] = nbx; by[j] = nby;
                } else {
                    const double nax = fma(-tj, bx[j], ax[j]), nay = fma(-tj, by[j], ay[j]);
                    const double nbx = fma(tj, ax[j], bx[j]), nby = fma(tj, ay[j], by[j]);
                    ax[j] = nax; ay[j] = nay; bx[j] = nbx; by[j] = nby;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < ROT_PAIRS; ++j) {
            const uint32_t pidx = threadIdx.x + (it * ROT_PAIRS + j) * blockDim.x;
            const uint32_t lj = insert0(pidx, hb);
            if (pidx < half) {
                tile[lj] = make_double2(cscale * ax[j], cscale * ay[j]);
                tile[lj ^ lx] = make_double2(cscale * bx[j], cscale * by[j]);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------
// Dedicated kernel for passes that consist only of fast (tangent-form) rotation runs -- the UCC case.
// Kept separate from the general kernel so that it fits in 64 registers (2 x 512 threads per SM).
// ------------------------------------------------------------------------------------------
// k_tile_rot works on ORBITS: a thread holds the 8 amplitudes x_beta = psi[l0 ^ beta0 v0 ^ beta1 v1 ^ beta2 v2]
// (v_i = tile-index vectors chosen by the host, l0 = the coset representative with zeros at the 3 pivot bits).
// Any rotation whose X-mask lies in span(v0, v1, v2) acts inside the orbit, so up to three different generators
// (24 Pauli strings) are applied between one shared-memory load and store of the amplitudes.
struct RotOp {         // 32 bytes, shared-memory copy of a fast rotation, refreshed per tile
    double t;          // tan(angle) * (unit-phase sign) * (-1)^popc(outside-tile index bits & z)
    uint32_t lz;       // Z letters inside the tile
    uint32_t mq[3];    // sign words (0 or 0x80000000): parity(v_i & lz), the sign step along orbit direction i
    uint32_t pad[2];
};
struct DevSub {        // 16 bytes: consecutive fast rotations with the same orbit pattern and phase type
    uint32_t begin, len;   // into the pass-local RotOp table
    uint32_t c;            // pairs are (beta, beta ^ c), c in 1..7 (orbit coordinates of the X-mask)
    uint32_t imag;         // unit phase +-i (ny even) instead of +-1
};
// COLLAPSED runs.  Rotations with the same X-mask and phase type act in the same 2-d planes (a, b = a ^ lx) and
// commute, so their angles ADD: a run of R strings is ONE plane rotation by Phi(l) = sum_r (+-)phi_r, the signs
// being the strings' Z parities at l.  All parities agree with the first string's except on the few tile bits D
// where the strings' Z letters differ, so Phi(l) = (-1)^parity(l & lz_1) * F[l restricted to D]: the host tabulates
// cos F, sin F per pattern.  For the JW image of a fermionic excitation F vanishes for all but one occupation
// pattern -- the 8 strings of a double excitation touch 1/8 of the amplitude pairs, once.
struct DevColEntry {   // 32 bytes
    double c, s;       // cos F, sin F of this pattern
    uint32_t pat;      // the pattern as tile-index bits (a-side: the highest X bit is clear)
    uint32_t pad[3];
};
struct DevCol {        // 64 bytes
    uint64_t zout;         // Z letters outside the tile (same for every string of the run)
    uint32_t lx, lz;       // X-mask; Z letters of the first string inside the tile
    uint32_t n_active;     // patterns with a non-zero angle
    uint32_t nd;           // fixed positions (D and the highest X bit), ascending
    uint32_t dpos[6];      // as masks ~((1 << pos) - 1): insert0(l, pos) = l + (l & mask)
    uint32_t ent_begin;    // into the pass-local entry table
    uint32_t free_log;     // log2(number of a-side indices per pattern)
    uint32_t imag;
    uint32_t pad;
};
struct DevSuper {      // 64 bytes: one shared-memory round trip (orbit), or one collapsed run
    uint32_t e0, e1, e2;   // ascending pivot positions: zeros are inserted there into the thread index
    uint32_t sub_begin, sub_count;   // collapsed run: sub_count = 0xffffffff, sub_begin = index of its DevCol
    uint32_t hb_log;       // tiles with one pair per thread: log2(pairs in the tile)
    double cscale;         // 1.0, or the pending product of cosines when it must be applied now (overflow guard)
    uint32_t off[8];       // off[beta] = XOR of the v_i selected by beta
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// asynchronous tile load: no register staging, every thread's 16-byte copies are all in flight at once
__device__ __forceinline__ void tile_load_async(double2* tile, const Shards& src, const TileGeom& g, uint64_t base) {
    const uint32_t ts = 1u << g.tbits;
    for (uint32_t k = threadIdx.x; k < ts; k += blockDim.x) cp_async16(tile + k, amp_addr(g, src, base, k));
}
// Per-thread address pieces of the tile gather.  Element k = tid + j * blockDim of a tile sits at amplitude index
// base | scat[k >> lbits] | (k & lmask); the deposit `scat` is bitwise, so with a power-of-two block size it splits
// into a per-thread constant (a_fix, one global load per kernel) and a per-j constant (s_boff[j], shared memory):
// the per-element address is one 64-bit OR and one add instead of a table load and a handful of shifts.
struct TileAddr {
    uint64_t a_fix;
    bool fast;
};
__device__ __forceinline__ TileAddr tile_addr_init(const TileGeom& g, uint64_t* s_boff /* >= 16 entries */) {
    TileAddr ta;
    const uint32_t ts = 1u << g.tbits;
    const uint32_t top = g.tbits - 1u;
    ta.fast = (ts % blockDim.x == 0) && (ts / blockDim.x <= 16u) && (blockDim.x >> g.lbits) >= 1u &&
              (!g.vbit || blockDim.x <= (1u << top));
    ta.a_fix = 0;
    if (ta.fast) {
        const uint32_t lmask = (1u << g.lbits) - 1u;
        if (threadIdx.x < ts / blockDim.x) {
            uint32_t kk = threadIdx.x * blockDim.x;
            if (g.vbit) kk &= (1u << top) - 1u;
            s_boff[threadIdx.x] = __ldg(g.scat + (kk >> g.lbits));
        }
        ta.a_fix = __ldg(g.scat + (threadIdx.x >> g.lbits)) | (uint64_t)(threadIdx.x & lmask);
    }
    __syncthreads();
    return ta;
}
__device__ __forceinline__ double2* amp_addr_fast(const TileGeom& g, const Shards& sh, const TileAddr& ta,
                                                  const uint64_t* s_boff, uint64_t base, uint32_t j) {
    double2* p = (g.vbit && ((j * blockDim.x) >> (g.tbits - 1u))) ? sh.p1 : sh.p0;
    return p + (base | ta.a_fix | s_boff[j]);
}
// ---- bulk asynchronous tile load (TMA engine, 1-d form): a tile is 2^(T-L) contiguous segments of 16 * 2^L bytes,
// each fetched by ONE cp.async.bulk that signals an mbarrier with its byte count; the threads only wait on the
// barrier's phase.  No per-thread address arithmetic, no registers, copies in flight while the per-tile tables are
// refreshed.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t phase) {
    // bounded spin: a protocol error must never hang the GPU.  Returns false on a timeout; every caller then raises the
    // context's device error flag (vqe_ctx::d_err = 2), which vqe_synchronize / the reductions / vqe_shard_status report.
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t ok;
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
        if (ok) return true;
    }
    return false;
}
__device__ __forceinline__ void tile_load_bulk(double2* tile, const Shards& src, const TileGeom& g, uint64_t base, uint64_t* bar) {
    const uint32_t ts = 1u << g.tbits;
    const uint32_t nseg = ts >> g.lbits;
    const uint32_t seg_bytes = 16u << g.lbits;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses to the tile come first
    if (threadIdx.x == 0) mbar_arrive_expect_tx(bar, ts * 16u);
    for (uint32_t sgm = threadIdx.x; sgm < nseg; sgm += blockDim.x) {
        const uint32_t k = sgm << g.lbits;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(tile + k)),
                     "l"(amp_addr(g, src, base, k)), "r"(seg_bytes), "r"(smem_u32(bar))
                     : "memory");
    }
}

// bulk asynchronous tile store: one cp.async.bulk per segment, shared -> global; returns once the copies have READ
// the shared-memory tile (it may then be refilled), the global writes complete in the background
__device__ __forceinline__ void tile_store_bulk(const double2* tile, const Shards& dst, const TileGeom& g, uint64_t base) {
    const uint32_t nseg = (1u << g.tbits) >> g.lbits;
    const uint32_t seg_bytes = 16u << g.lbits;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the tile was written through the generic proxy
    for (uint32_t sgm = threadIdx.x; sgm < nseg; sgm += blockDim.x) {
        const uint32_t k = sgm << g.lbits;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(amp_addr(g, dst, base, k)),
                     "r"(smem_u32(tile + k)), "r"(seg_bytes)
                     : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- tensor-map tile load / store: issued by ONE thread -------------------------------------------------------------
__device__ __forceinline__ void tma_coords(const TmaGeom& tg, uint64_t idx, int (&c)[5]) {
#pragma unroll
    for (int i = 0; i < 5; ++i) c[i] = (int)((uint32_t)(idx >> tg.shift[i]) & tg.cmask[i]);
}
__device__ __forceinline__ void tma_load_tile(double2* tile, const CUtensorMap* tm, const TileGeom& g, uint64_t base, uint64_t* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses to the slot come first
    const uint32_t esz = g.rl ? 8u : 16u;
    mbar_arrive_expect_tx(bar, esz << g.tbits);
    for (uint32_t r = 0; r < g.tg.n_req; ++r) {
        int c[5];
        tma_coords(g.tg, base | g.tg.req_bits[r], c);
        if (!g.rl) c[0] <<= 1;  // dimension 0 counts doubles: (re, im) per amplitude, one per amplitude in the real layout
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                smem_u32(reinterpret_cast<char*>(tile) + (size_t)r * g.tg.req_amps * esz)),
            "l"(tm), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(smem_u32(bar))
            : "memory");
    }
}
// the caller has made the tile's generic writes visible to the async proxy (fence.proxy.async by the writers + barrier)
__device__ __forceinline__ void tma_store_tile(const double2* tile, const CUtensorMap* tm, const TileGeom& g, uint64_t base) {
    const uint32_t esz = g.rl ? 8u : 16u;
    for (uint32_t r = 0; r < g.tg.n_req; ++r) {
        int c[5];
        tma_coords(g.tg, base | g.tg.req_bits[r], c);
        if (!g.rl) c[0] <<= 1;
        asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(tm), "r"(c[0]),
                     "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]),
                     "r"(smem_u32(reinterpret_cast<const char*>(tile) + (size_t)r * g.tg.req_amps * esz))
                     : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__device__ __forceinline__ void tile_load_async_fast(double2* tile, const Shards& src, const TileGeom& g, const TileAddr& ta,
                                                     const uint64_t* s_boff, uint64_t base) {
    if (!ta.fast) {
        tile_load_async(tile, src, g, base);
        return;
    }
    const uint32_t nj = (1u << g.tbits) / blockDim.x;
    for (uint32_t j = 0; j < nj; ++j)
        cp_async16(tile + threadIdx.x + j * blockDim.x, amp_addr_fast(g, src, ta, s_boff, base, j));
}
__device__ __forceinline__ void tile_store_scaled_fast(const double2* tile, const Shards& dst, const TileGeom& g,
                                                       const TileAddr& ta, const uint64_t* s_boff, uint64_t base,
                                                       double scale) {
    const uint32_t ts = 1u << g.tbits;
    if (!ta.fast) {
        for (uint32_t k = threadIdx.x; k < ts; k += blockDim.x) {
            const double2 v = tile[k];
            *amp_addr(g, dst, base, k) = make_double2(scale * v.x, scale * v.y);
        }
        return;
    }
    const uint32_t nj = ts / blockDim.x;
    if (scale == 1.0) {  // every run collapsed: nothing to rescale
        for (uint32_t j = 0; j < nj; ++j) *amp_addr_fast(g, dst, ta, s_boff, base, j) = tile[threadIdx.x + j * blockDim.x];
    } else {
        for (uint32_t j = 0; j < nj; ++j) {
            const double2 v = tile[threadIdx.x + j * blockDim.x];
            *amp_addr_fast(g, dst, ta, s_boff, base, j) = make_double2(scale * v.x, scale * v.y);
        }
    }
}
// One sub-run on the 8-amplitude orbit of a thread: the rotations pair (beta, beta ^ C).
//   Tangent form R = c [[1, -+t], [+-t, 1]]: the unnormalised update costs one DFMA per real component; the
//   cosines are collected by the host into ONE scale per pass, applied when the tile is stored (every rotation
//   touches every amplitude, so the product is common to the whole tile).
//   Sign of pair beta (the element whose bit ctz(C) is clear): (-1)^parity(l & lz) = s(thread) * prod_i mq_i^beta_i:
//   one POPC per rotation for s and three XORs for the four pair variants of t -- no per-pair work besides the
//   DFMAs: LDS.128 x2 + 8 integer + 16 DFMA per rotation (8 DFMA when REAL), FP64-pipe bound.
//   REAL: the tile is purely real and the phase is +-1: the imaginary halves are skipped.
//   Two rotations per trip with the partner registers ping-ponging (b -> tmp -> b): no register copies.
template <int C, bool IMAG, bool REAL>
__device__ __forceinline__ void orbit_sub(double2 (&x)[8], const RotOp* __restrict__ tab, int len, uint32_t l0) {
    constexpr int K = (C & 1) ? 0 : ((C & 2) ? 1 : 2);              // pivot coordinate: a-side has beta_K = 0
    constexpr int I = (K == 0) ? 1 : 0, J = (K == 2) ? 1 : 2;       // the two free coordinates, I < J
    constexpr int A0 = 0, A1 = 1 << I, A2 = 1 << J, A3 = (1 << I) | (1 << J);  // a-side elements, pair index = beta_I + 2 beta_J
#define ORB_STEP(A, B, D, T)                                                                      \
    {                                                                                            \
        if (IMAG) {                                                                              \
            D.x = fma(-T, A.y, B.x);                                                             \
            D.y = fma(T, A.x, B.y);                                                              \
            A.x = fma(-T, B.y, A.x);                                                             \
            A.y = fma(T, B.x, A.y);                                                              \
        } else {                                                                                 \
            D.x = fma(T, A.x, B.x);                                                              \
            A.x = fma(-T, B.x, A.x);                                                             \
            if (!REAL) {                                                                         \
                D.y = fma(T, A.y, B.y);                                                          \
                A.y = fma(-T, B.y, A.y);                                                         \
            }                                                                                    \
        }                                                                                        \
    }
#define ORB_SIGNS(F, T0, T1, T2, T3)                                                              \
    double T0, T1, T2, T3;                                                                        \
    {                                                                                            \
        const uint32_t hi = (uint32_t)__double2hiint((F).t) ^ ((uint32_t)__popc(l0 & (F).lz) << 31); \
        const int lo = __double2loint((F).t);                                                    \
        T0 = __hiloint2double((int)hi, lo);                                                      \
        T1 = __hiloint2double((int)(hi ^ (F).mq[I]), lo);                                        \
        T2 = __hiloint2double((int)(hi ^ (F).mq[J]), lo);                                        \
        T3 = __hiloint2double((int)(hi ^ (F).mq[I] ^ (F).mq[J]), lo);                            \
    }
    double2 d0, d1, d2, d3;
    int r = 0;
#pragma unroll 1
    for (; r + 1 < len; r += 2) {
        const RotOp f = tab[r], h = tab[r + 1];
        ORB_SIGNS(f, t0, t1, t2, t3)
        ORB_SIGNS(h, u0, u1, u2, u3)
        ORB_STEP(x[A0], x[A0 ^ C], d0, t0)
        ORB_STEP(x[A1], x[A1 ^ C], d1, t1)
        ORB_STEP(x[A2], x[A2 ^ C], d2, t2)
        ORB_STEP(x[A3], x[A3 ^ C], d3, t3)
        ORB_STEP(x[A0], d0, x[A0 ^ C], u0)
        ORB_STEP(x[A1], d1, x[A1 ^ C], u1)
        ORB_STEP(x[A2], d2, x[A2 ^ C], u2)
        ORB_STEP(x[A3], d3, x[A3 ^ C], u3)
    }
    if (r < len) {
        const RotOp f = tab[r];
        ORB_SIGNS(f, t0, t1, t2, t3)
        ORB_STEP(x[A0], x[A0 ^ C], d0, t0)
        ORB_STEP(x[A1], x[A1 ^ C], d1, t1)
        ORB_STEP(x[A2], x[A2 ^ C], d2, t2)
        ORB_STEP(x[A3], x[A3 ^ C], d3, t3)
        x[A0 ^ C].x = d0.x; x[A1 ^ C].x = d1.x; x[A2 ^ C].x = d2.x; x[A3 ^ C].x = d3.x;
        if (!REAL) { x[A0 ^ C].y = d0.y; x[A1 ^ C].y = d1.y; x[A2 ^ C].y = d2.y; x[A3 ^ C].y = d3.y; }
    }
#undef ORB_SIGNS
#undef ORB_STEP
}

template <bool IMAG, bool REAL>
__device__ __forceinline__ void orbit_dispatch(double2 (&x)[8], const RotOp* __restrict__ tab, int len, uint32_t l0, uint32_t c) {
    switch (c) {  // warp-uniform
        case 1: orbit_sub<1, IMAG, REAL>(x, tab, len, l0); break;
        case 2: orbit_sub<2, IMAG, REAL>(x, tab, len, l0); break;
        case 3: orbit_sub<3, IMAG, REAL>(x, tab, len, l0); break;
        case 4: orbit_sub<4, IMAG, REAL>(x, tab, len, l0); break;
        case 5: orbit_sub<5, IMAG, REAL>(x, tab, len, l0); break;
        case 6: orbit_sub<6, IMAG, REAL>(x, tab, len, l0); break;
        default: orbit_sub<7, IMAG, REAL>(x, tab, len, l0); break;
    }
}

// tiles with a single pair per thread (registers of fewer than 12 local qubits: tests and tiny molecules)
template <bool REAL>
__device__ __forceinline__ void rot_sub1(double2* tile, const RotOp* __restrict__ tab, const DevSub& sb, uint32_t lx, uint32_t hb,
                                         uint32_t n_pairs, double cs) {
    if (threadIdx.x >= n_pairs) return;
    const uint32_t l0 = insert0(threadIdx.x, hb);
    double2 a = tile[l0], b = tile[l0 ^ lx];
    for (uint32_t r = 0; r < sb.len; ++r) {
        const RotOp f = tab[r];
        const double t = flipsign(f.t, __popc(l0 & f.lz));
        double2 na, nb;
        if (sb.imag && !REAL) {
            nb.x = fma(-t, a.y, b.x); nb.y = fma(t, a.x, b.y);
            na.x = fma(-t, b.y, a.x); na.y = fma(t, b.x, a.y);
        } else {
            nb.x = fma(t, a.x, b.x); nb.y = fma(t, a.y, b.y);
            na.x = fma(-t, b.x, a.x); na.y = fma(-t, b.y, a.y);
        }
        a = na;
        b = nb;
    }
    if (cs != 1.0) { a.x *= cs; a.y *= cs; b.x *= cs; b.y *= cs; }
    tile[l0] = a;
    tile[l0 ^ lx] = b;
}

// ------------------------------------------------------------------------------------------
// Warp-parallel tile base.  base = pdep(tile number, comp_mask): lane i owns the i-th set bit of comp_mask (found once
// per kernel), a tile costs one shift/and per lane and two warp-wide OR reductions (REDUX) instead of a serial
// bit loop in every thread.
// ------------------------------------------------------------------------------------------
struct BaseLane {
    uint32_t pos;   // index-bit position of this lane's bit of comp_mask
    uint32_t act;   // 0: comp_mask has fewer set bits than this lane's number
};
__device__ __forceinline__ BaseLane base_lane_init(const TileGeom& g) {
    BaseLane bl;
    const uint32_t lane = threadIdx.x & 31u;
    uint64_t m = g.comp_mask;
    for (uint32_t i = 0; i < lane && m; ++i) m &= m - 1;  // drop the lane lowest set bits
    bl.act = m ? 1u : 0u;
    bl.pos = m ? (uint32_t)(__ffsll((long long)m) - 1) : 0u;
    return bl;
}
__device__ __forceinline__ uint64_t tile_base_warp(const TileGeom& g, const BaseLane& bl, uint64_t t) {
    const uint64_t tnum = g.rev ? g.tile_first - t * g.tile_stride : g.tile_first + t * g.tile_stride;
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t bit = (bl.act && ((tnum >> lane) & 1ull)) ? (1ull << bl.pos) : 0ull;
    const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)bit);
    const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(bit >> 32));
    return ((uint64_t)hi << 32) | lo;
}

// REAL: the state is known to be purely real on entry and every rotation of the pass has a +-1 phase (ny odd: the
// UCC case -- JW images of T - T^dagger), so the imaginary parts stay exactly zero and are never touched.
// phase A of a gather-form peer pass: copy the needed partner amplitudes of tiles [g.tile_first, +g.n_tiles) into
// the local staging buffer
// One WARP per tile: the lanes walk the tile's need groups, one group = 64 contiguous bytes = four independent 16-byte
// loads in flight per lane (remote reads over NVLink have microseconds of latency: the copy is latency-bound, not
// issue-bound); the tile base is formed warp-parallel.  Called by the stand-alone kernel for the first chunk of a pass
// and by the gather CTAs that ride along in the pass kernel of chunk k to fetch chunk k + 1 under its arithmetic.
__device__ __forceinline__ void gather_need_warps(const Shards& psi, const TileGeom& g, const GatherGeom& gg, uint64_t tile_first,
                                                  uint64_t n_tiles, double2* __restrict__ stage_out, uint32_t warp_id, uint32_t n_warps) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lmask = (1u << g.lbits) - 1u;
    const double2* src = gg.own_half ? psi.p0 : psi.p1;  // the partner's shard
    const BaseLane bl = base_lane_init(g);
    for (uint64_t t = warp_id; t < n_tiles; t += n_warps) {
        const uint64_t tnum = tile_first + t * g.tile_stride;
        const uint64_t bit = (bl.act && ((tnum >> lane) & 1ull)) ? (1ull << bl.pos) : 0ull;
        const uint64_t base = ((uint64_t)__reduce_or_sync(0xffffffffu, (uint32_t)(bit >> 32)) << 32) | __reduce_or_sync(0xffffffffu, (uint32_t)bit);
        double2* out = stage_out + t * (uint64_t)gg.n_need * 4u;
        for (uint32_t q = lane; q < gg.n_need; q += 32u) {
            const uint32_t kl = (uint32_t)__ldg(gg.need + q) * 4u;  // tile-local index inside the partner's half
            const double2* p = src + (base | __ldg(g.scat + (kl >> g.lbits)) | (uint64_t)(kl & lmask));
            const double2 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
            double2* o = out + (uint64_t)q * 4u;
            o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
        }
    }
}
__global__ void __launch_bounds__(256) k_gather_need(Shards psi, TileGeom g, GatherGeom gg, double2* __restrict__ stage_out) {
    gather_need_warps(psi, g, gg, g.tile_first, g.n_tiles, stage_out, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5),
                      gridDim.x * (blockDim.x >> 5));
}

template <bool REAL>
__global__ void __launch_bounds__(512, REAL ? 3 : 2) k_tile_rot(Shards psi, TileGeom g, GatherGeom gg,
                                                     const DevOp* __restrict__ ops, int n_ops,
                                                     const DevSuper* __restrict__ supers, int n_supers,
                                                     const DevSub* __restrict__ subs, int n_subs,
                                                     const DevCol* __restrict__ cols, int n_cols,
                                                     const DevColEntry* __restrict__ ents, int n_ents, double pass_scale,
                                                     int* __restrict__ err, GatherGeom gnext, uint64_t next_first, uint64_t next_tiles,
                                                     double2* __restrict__ next_stage, uint32_t n_gctas) {
    extern __shared__ __align__(1024) double2 tile[];
    // gather-form peer pass, chunk k: the first n_gctas CTAs fetch the partner amplitudes chunk k + 1 needs into the other
    // staging buffer while the remaining CTAs run the pass on chunk k (NVLink reads under HBM-bound arithmetic)
    if (blockIdx.x < n_gctas) {
        gather_need_warps(psi, g, gnext, next_first, next_tiles, next_stage, blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5),
                          n_gctas * (blockDim.x >> 5));
        return;
    }
    const uint32_t cta = blockIdx.x - n_gctas, n_cta = gridDim.x - n_gctas;
    const uint32_t ts = 1u << g.tbits;
    const bool four = (ts >> 1) == 4u * blockDim.x;  // host guarantees: 4 pairs (one orbit) per thread, or at most 1 pair
    RotOp* optab = (RotOp*)(tile + ts);
    DevSuper* ssup = (DevSuper*)(optab + n_ops);
    DevCol* scol = (DevCol*)(ssup + n_supers);
    DevColEntry* sent = (DevColEntry*)(scol + n_cols);
    DevSub* ssub = (DevSub*)(sent + n_ents);
    uint32_t* scsign = (uint32_t*)(ssub + n_subs);  // per tile: outside-tile Z parity of every collapsed run
    __shared__ uint64_t s_boff[16];
    __shared__ __align__(8) uint64_t s_mbar;
    const bool bulk = g.bulk && !gg.n_need;
    uint32_t mphase = 0;
    if (bulk && threadIdx.x == 0) mbar_init(&s_mbar, 1);
    const TileAddr ta = tile_addr_init(g, s_boff);
    for (int q = threadIdx.x; q < n_supers; q += blockDim.x) ssup[q] = supers[q];
    for (int q = threadIdx.x; q < n_subs; q += blockDim.x) ssub[q] = subs[q];
    for (int q = threadIdx.x; q < n_cols; q += blockDim.x) scol[q] = cols[q];
    for (int q = threadIdx.x; q < n_ents; q += blockDim.x) sent[q] = ents[q];
    for (int r = threadIdx.x; r < n_ops; r += blockDim.x) {  // tile-independent part of the table
        RotOp f;
        f.t = 0.0;
        f.lz = ops[r].lz;
        f.mq[0] = (ops[r].jmask & 1u) << 31;
        f.mq[1] = (ops[r].jmask & 2u) << 30;
        f.mq[2] = (ops[r].jmask & 4u) << 29;
        f.pad[0] = f.pad[1] = 0;
        optab[r] = f;
    }
    for (uint64_t t = cta; t < g.n_tiles; t += n_cta) {
        const uint64_t base = tile_base(g, t);
        const uint64_t sbase = base | g.sign_base;
        __syncthreads();
        if (gg.n_need) {
            // gather form: own half from the own shard, partner half = zeros + the staged amplitudes
            const uint32_t half_t = ts >> 1;
            const uint32_t own_off = gg.own_half ? half_t : 0u, par_off = gg.own_half ? 0u : half_t;
            for (uint32_t k = threadIdx.x; k < half_t; k += blockDim.x) {
                cp_async16(tile + own_off + k, amp_addr(g, psi, base, own_off + k));
                tile[par_off + k] = make_double2(0.0, 0.0);
            }
            __syncthreads();
            const double2* st = gg.stage + t * (uint64_t)gg.n_need * 4u;
            for (uint32_t q = threadIdx.x; q < gg.n_need * 4u; q += blockDim.x)
                cp_async16(tile + par_off + (uint32_t)gg.need[q >> 2] * 4u + (q & 3u), st + q);
        } else if (bulk) {
            tile_load_bulk(tile, psi, g, base, &s_mbar);
        } else {
            tile_load_async_fast(tile, psi, g, ta, s_boff, base);
        }
        for (int r = threadIdx.x; r < n_ops; r += blockDim.x)
            optab[r].t = flipsign(ops[r].s, __popcll(sbase & ops[r].zout));
        for (int r = threadIdx.x; r < n_cols; r += blockDim.x) scsign[r] = (uint32_t)__popcll(sbase & scol[r].zout) & 1u;
        if (bulk) {
            if (!mbar_wait(&s_mbar, mphase) && err) *err = 2;
            mphase ^= 1u;
        } else {
            cp_async_wait_all();
        }
        for (int q = 0; q < n_supers; ++q) {
            __syncthreads();
            const DevSuper& su = ssup[q];
            if (su.sub_count == 0xffffffffu) {
                // collapsed run: one plane rotation per ACTIVE (pattern, free index) pair
                const DevCol& co = scol[su.sub_begin];
                const uint32_t items = co.n_active << co.free_log;
                const uint32_t fmask = (1u << co.free_log) - 1u;
                const uint32_t tsig = scsign[su.sub_begin];
                for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
                    const DevColEntry en = sent[co.ent_begin + (it >> co.free_log)];
                    uint32_t l = it & fmask;
                    // deposit the free index around the fixed positions (unused slots hold a zero mask)
                    l += l & co.dpos[0];
                    l += l & co.dpos[1];
                    l += l & co.dpos[2];
                    l += l & co.dpos[3];
                    if (co.nd > 4) {
                        l += l & co.dpos[4];
                        l += l & co.dpos[5];
                    }
                    l |= en.pat;
                    const double sn = flipsign(en.s, tsig + (uint32_t)__popc(l & co.lz));
                    if (REAL) {
                        const double a = tile[l].x, b = tile[l ^ co.lx].x;
                        tile[l].x = fma(en.c, a, -sn * b);
                        tile[l ^ co.lx].x = fma(en.c, b, sn * a);
                    } else {
                        const double2 a = tile[l], b = tile[l ^ co.lx];
                        double2 na, nb;
                        if (co.imag) {  // a' = c a + i s b, b' = c b + i s a
                            na.x = fma(en.c, a.x, -sn * b.y); na.y = fma(en.c, a.y, sn * b.x);
                            nb.x = fma(en.c, b.x, -sn * a.y); nb.y = fma(en.c, b.y, sn * a.x);
                        } else {        // a' = c a - s b, b' = c b + s a
                            na.x = fma(en.c, a.x, -sn * b.x); na.y = fma(en.c, a.y, -sn * b.y);
                            nb.x = fma(en.c, b.x, sn * a.x); nb.y = fma(en.c, b.y, sn * a.y);
                        }
                        tile[l] = na;
                        tile[l ^ co.lx] = nb;
                    }
                }
            } else if (four) {
                const uint32_t l0 = insert0(insert0(insert0(threadIdx.x, su.e0), su.e1), su.e2);
                double2 x[8];
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if (REAL) { x[b].x = tile[l0 ^ su.off[b]].x; x[b].y = 0.0; }
                    else x[b] = tile[l0 ^ su.off[b]];
                }
                for (uint32_t sidx = 0; sidx < su.sub_count; ++sidx) {
                    const DevSub sb = ssub[su.sub_begin + sidx];
                    if (REAL) orbit_dispatch<false, true>(x, optab + sb.begin, (int)sb.len, l0, sb.c);
                    else if (sb.imag) orbit_dispatch<true, false>(x, optab + sb.begin, (int)sb.len, l0, sb.c);
                    else orbit_dispatch<false, false>(x, optab + sb.begin, (int)sb.len, l0, sb.c);
                }
                const double cs = su.cscale;
                if (cs != 1.0) {
#pragma unroll
                    for (int b = 0; b < 8; ++b) { x[b].x *= cs; x[b].y *= cs; }
                }
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if (REAL) tile[l0 ^ su.off[b]].x = x[b].x;
                    else tile[l0 ^ su.off[b]] = x[b];
                }
            } else {
                // one sub-run per round trip: su.off[1] = the X-mask, su.e0 = its highest bit
                const DevSub sb = ssub[su.sub_begin];
                rot_sub1<REAL>(tile, optab + sb.begin, sb, su.off[1], su.e0, 1u << su.hb_log, su.cscale);
            }
        }
        __syncthreads();
        if (gg.n_need) {
            const uint32_t half_t = ts >> 1;
            const uint32_t own_off = gg.own_half ? half_t : 0u;
            for (uint32_t k = threadIdx.x; k < half_t; k += blockDim.x) {
                const double2 v = tile[own_off + k];
                *amp_addr(g, psi, base, own_off + k) = make_double2(pass_scale * v.x, pass_scale * v.y);
            }
        } else if (bulk && pass_scale == 1.0) {
            tile_store_bulk(tile, psi, g, base);
        } else {
            tile_store_scaled_fast(tile, psi, g, ta, s_boff, base, pass_scale);
        }
    }
    if (bulk) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores have landed
}

// ------------------------------------------------------------------------------------------
// k_tile_col: passes that consist ONLY of collapsed runs / tabulated plane rotations (every pass of a UCCSD or
// QUCCSD program).  A collapsed JW double excitation is 256 plane rotations per 4096-amplitude tile, so the kernel is
// built around the per-run cost: 256-thread CTAs (one item per thread, three CTAs per SM, 8-warp barriers), the
// descriptor read and the index deposit of run q+1 are issued BEFORE the barrier that ends run q, so the chain
// between two barriers is  LDS pair -> 3 FP64 ops -> STS pair.
// ------------------------------------------------------------------------------------------
// Per-run descriptor in shared memory (64 bytes, built once per CTA from DevCol + the run's first table entry).  Runs with
// a single active pattern -- every JW single / double excitation -- never touch the entry table again.
struct ColLite {
    double c, s;             // cos, sin of entry 0
    uint32_t lxs, lz;        // X-mask as an element index in the tile's shared-memory layout (swizzled when the tile is) | Z letters
    uint32_t pat, items;     // pattern of entry 0 | n_active << free_log
    uint32_t dm[6];          // deposit masks of the fixed positions, ascending, 0 when unused
    uint32_t free_log;
    uint32_t meta;           // bit 0: unit phase +-i | bit 1: single active pattern | bits 16..: first entry of the run
};
static_assert(sizeof(ColLite) == 64, "ColLite layout");
struct ColItem {
    uint32_t ls, lxs;    // a-side element (shared-memory layout) | X-mask (shared-memory layout)
    double c, sn;        // cos, signed sin
    uint32_t imag, valid;
    uint32_t items;      // items of the run (for the threads that own more than one)
};
__device__ __forceinline__ void col_lite_build(ColLite* lite, const DevCol* __restrict__ cols, const DevColEntry* __restrict__ ents,
                                               int n_cols, uint32_t swz, bool rl = false) {
    for (int q = threadIdx.x; q < n_cols; q += blockDim.x) {
        const DevCol co = cols[q];
        const DevColEntry e0 = ents[co.ent_begin];
        ColLite L;
        L.c = e0.c;
        L.s = e0.s;
        L.lxs = rl ? swz_idx8(co.lx, swz) : swz_idx(co.lx, swz);
        L.lz = co.lz;
        L.pat = e0.pat;
        L.items = co.n_active << co.free_log;
#pragma unroll
        for (int d = 0; d < 6; ++d) L.dm[d] = co.dpos[d];
        L.free_log = co.free_log;
        L.meta = (co.imag & 1u) | (co.n_active == 1u ? 2u : 0u) | (co.ent_begin << 16);
        lite[q] = L;
    }
}
template <bool RL = false>
__device__ __forceinline__ ColItem col_prep(const ColLite* lite, const DevColEntry* sent, const uint32_t* csign, int q, int n_cols,
                                            uint32_t it, uint32_t swz) {
    ColItem ci;
    ci.valid = 0;
    ci.ls = ci.lxs = ci.imag = ci.items = 0;
    ci.c = 1.0;
    ci.sn = 0.0;
    if (q >= n_cols) return ci;
    const ColLite& L = lite[q];
    ci.items = L.items;
    ci.valid = it < L.items ? 1u : 0u;
    uint32_t l = it & ((1u << L.free_log) - 1u);
    l += l & L.dm[0];
    l += l & L.dm[1];
    l += l & L.dm[2];
    l += l & L.dm[3];
    l += l & L.dm[4];
    l += l & L.dm[5];
    double c = L.c, sv = L.s;
    uint32_t pat = L.pat;
    if (!(L.meta & 2u) && ci.valid) {  // several active patterns (tabulated plane rotations): this item's entry
        const DevColEntry& en = sent[(L.meta >> 16) + (it >> L.free_log)];
        c = en.c;
        sv = en.s;
        pat = en.pat;
    }
    l |= pat;
    ci.ls = RL ? swz_idx8(l, swz) : swz_idx(l, swz);
    ci.lxs = L.lxs;
    ci.c = c;
    ci.sn = flipsign(sv, csign[q] + (uint32_t)__popc(l & L.lz));
    ci.imag = L.meta & 1u;
    return ci;
}
// one plane rotation; LOADS first, so that the caller's index work for the next run overlaps their latency
template <bool REAL>
struct ColPair {
    double2 a, b;
};
template <bool REAL, bool RL = false>
__device__ __forceinline__ void col_load(const double2* tile, const ColItem& ci, ColPair<REAL>& pr) {
    if (RL) {  // real layout: the tile holds doubles
        const double* tr = reinterpret_cast<const double*>(tile);
        pr.a.x = tr[ci.ls];
        pr.b.x = tr[ci.ls ^ ci.lxs];
        pr.a.y = pr.b.y = 0.0;
    } else if (REAL) {
        pr.a.x = tile[ci.ls].x;
        pr.b.x = tile[ci.ls ^ ci.lxs].x;
        pr.a.y = pr.b.y = 0.0;
    } else {
        pr.a = tile[ci.ls];
        pr.b = tile[ci.ls ^ ci.lxs];
    }
}
template <bool REAL, bool RL = false>
__device__ __forceinline__ void col_store(double2* tile, const ColItem& ci, const ColPair<REAL>& pr) {
    const double2 a = pr.a, b = pr.b;
    if (RL) {
        double* tr = reinterpret_cast<double*>(tile);
        tr[ci.ls] = fma(ci.c, a.x, -ci.sn * b.x);
        tr[ci.ls ^ ci.lxs] = fma(ci.c, b.x, ci.sn * a.x);
    } else if (REAL) {
        tile[ci.ls].x = fma(ci.c, a.x, -ci.sn * b.x);
        tile[ci.ls ^ ci.lxs].x = fma(ci.c, b.x, ci.sn * a.x);
    } else {
        double2 na, nb;
        if (ci.imag) {  // a' = c a + i s b, b' = c b + i s a
            na.x = fma(ci.c, a.x, -ci.sn * b.y); na.y = fma(ci.c, a.y, ci.sn * b.x);
            nb.x = fma(ci.c, b.x, -ci.sn * a.y); nb.y = fma(ci.c, b.y, ci.sn * a.x);
        } else {        // a' = c a - s b, b' = c b + s a
            na.x = fma(ci.c, a.x, -ci.sn * b.x); na.y = fma(ci.c, a.y, -ci.sn * b.y);
            nb.x = fma(ci.c, b.x, ci.sn * a.x); nb.y = fma(ci.c, b.y, ci.sn * a.y);
        }
        tile[ci.ls] = na;
        tile[ci.ls ^ ci.lxs] = nb;
    }
}
// all runs of a pass on the tile in shared memory; ends with a CTA barrier (every thread has passed `fence` before it)
template <bool REAL, bool RL = false>
__device__ __forceinline__ void col_runs(double2* tile, const ColLite* lite, const DevColEntry* sent, const uint32_t* csign, int n_cols,
                                         uint32_t swz, bool fence, ColItem cur) {
    const uint32_t bd = blockDim.x;
    for (int q = 0; q < n_cols; ++q) {
        const uint32_t items = cur.items;
        ColPair<REAL> pr;
        if (cur.valid) col_load<REAL, RL>(tile, cur, pr);
        const ColItem nxt = col_prep<RL>(lite, sent, csign, q + 1, n_cols, threadIdx.x, swz);  // overlaps the loads above
        if (cur.valid) col_store<REAL, RL>(tile, cur, pr);
        for (uint32_t it = threadIdx.x + bd; it < items; it += bd) {
            const ColItem ci = col_prep<RL>(lite, sent, csign, q, n_cols, it, swz);
            ColPair<REAL> p2;
            col_load<REAL, RL>(tile, ci, p2);
            col_store<REAL, RL>(tile, ci, p2);
        }
        cur = nxt;
        if (fence && q + 1 == n_cols) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my writes -> the async-proxy store
        __syncthreads();
    }
}
// RL: the buffer is in the REAL LAYOUT (vqe_ctx::real_layout): tile elements are doubles (half the HBM traffic and half the
// shared-memory wavefronts of the interleaved form), tensor-map staging only.
template <bool REAL, bool RL>
__global__ void __launch_bounds__(256, RL ? 4 : 3) k_tile_col(const __grid_constant__ CUtensorMap tmap, Shards psi, TileGeom g,
                                                     const DevCol* __restrict__ cols, int n_cols,
                                                     const DevColEntry* __restrict__ ents, int n_ents, int* __restrict__ err) {
    extern __shared__ __align__(1024) double2 tile[];
    const uint32_t ts = 1u << g.tbits;
    ColLite* lite = RL ? (ColLite*)(reinterpret_cast<double*>(tile) + ts) : (ColLite*)(tile + ts);
    DevColEntry* sent = (DevColEntry*)(lite + n_cols);
    uint64_t* szout = (uint64_t*)(sent + n_ents);
    uint32_t* scsign = (uint32_t*)(szout + n_cols);  // per tile: outside-tile Z parity of every run
    __shared__ uint64_t s_boff[16];
    __shared__ __align__(8) uint64_t s_mbar;
    const bool tma = g.tma != 0;
    const bool bulk = g.bulk != 0 || tma;
    uint32_t mphase = 0;
    if (bulk && threadIdx.x == 0) mbar_init(&s_mbar, 1);
    const TileAddr ta = tile_addr_init(g, s_boff);
    col_lite_build(lite, cols, ents, n_cols, g.swz, RL);
    for (int q = threadIdx.x; q < n_cols; q += blockDim.x) szout[q] = cols[q].zout;
    for (int q = threadIdx.x; q < n_ents; q += blockDim.x) sent[q] = ents[q];
    const BaseLane bl = base_lane_init(g);
    const uint32_t bd = blockDim.x;
    for (uint64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        const uint64_t base = tile_base_warp(g, bl, t);
        const uint64_t sbase = base | g.sign_base;
        __syncthreads();  // the previous tile has left shared memory (bulk store has read it); tables are in place
        if (tma) {
            if (threadIdx.x == 0) tma_load_tile(tile, &tmap, g, base, &s_mbar);
        } else if (bulk) tile_load_bulk(tile, psi, g, base, &s_mbar);
        else tile_load_async_fast(tile, psi, g, ta, s_boff, base);
        for (int r = threadIdx.x; r < n_cols; r += bd) scsign[r] = (uint32_t)__popcll(sbase & szout[r]) & 1u;
        __syncthreads();  // scsign visible
        const ColItem first = col_prep<RL>(lite, sent, scsign, 0, n_cols, threadIdx.x, g.swz);
        if (bulk) {
            if (!mbar_wait(&s_mbar, mphase) && err) *err = 2;
            mphase ^= 1u;
        } else {
            cp_async_wait_all();
            __syncthreads();
        }
        col_runs<REAL, RL>(tile, lite, sent, scsign, n_cols, g.swz, tma, first);
        if (tma) {
            if (threadIdx.x == 0) {
                tma_store_tile(tile, &tmap, g, base);
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else if (bulk) tile_store_bulk(tile, psi, g, base);
        else tile_store_scaled_fast(tile, psi, g, ta, s_boff, base, 1.0);
    }
    if (bulk) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores have landed
}

// ------------------------------------------------------------------------------------------
// k_col_tab: the real-layout collapsed-run pass with a host-built ITEM TABLE.  Which shared-memory element an item of a
// run touches (index deposit, pattern, swizzle) and the Z parity of that element depend only on the masks of the program,
// not on the tile or the angles, so the host tabulates them once per plan: one 32-bit word per item,
//   bits 0-15  byte offset of the a-side element in the (swizzled) tile,  bit 16  parity(l & lz),  bits 17-  pattern number.
// k_tile_col spends ~45 of its ~95 warp instructions per (warp, run) on that index work; here an item costs one coalesced
// 4-byte load (through L1: every tile of the pass reads the same words), a sign flip, the pair load, 4 FP64 operations and
// the pair store.  The words of run q+1 are fetched before the barrier that ends run q.
// Programmatic dependent launch: the prologue (descriptor build) runs under the tail of the previous pass; the state is
// first touched after griddepcontrol.wait.
// ------------------------------------------------------------------------------------------
struct ColTabRun {         // 32 bytes, shared memory, built once per CTA
    double c, s;           // cos, sin of entry 0 (runs with a single active pattern never read the entry table)
    uint32_t lxb;          // X-mask as a BYTE offset in the tile's shared-memory layout
    uint32_t items;
    uint32_t tab_off;      // first word of the run in the pass's item table
    uint32_t ent0;         // first entry of the run | bit 31: several active patterns
};
static_assert(sizeof(ColTabRun) == 32, "ColTabRun layout");
#define COLTAB_INVALID 0xffffffffu
__device__ __forceinline__ void coltab_rotate(char* tb, uint32_t w, const ColTabRun& R, const double2* scs, uint32_t cs, double a, double b) {
    double c = R.c, s = R.s;
    if (R.ent0 >> 31) {
        const double2 e = scs[(R.ent0 & 0x7fffffffu) + (w >> 17)];
        c = e.x;
        s = e.y;
    }
    const double sn = flipsign(s, (w >> 16) ^ cs);
    const uint32_t off = w & 0xffffu;
    *reinterpret_cast<double*>(tb + off) = fma(c, a, -sn * b);
    *reinterpret_cast<double*>(tb + (off ^ R.lxb)) = fma(c, b, sn * a);
}
__global__ void __launch_bounds__(256, 4) k_col_tab(const __grid_constant__ CUtensorMap tmap, TileGeom g, const DevCol* __restrict__ cols,
                                                    int n_cols, const DevColEntry* __restrict__ ents, int n_ents,
                                                    const uint32_t* __restrict__ tab, int skeleton, int* __restrict__ err) {
    extern __shared__ __align__(1024) double2 tile[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t ts = 1u << g.tbits;
    char* tb = reinterpret_cast<char*>(tile);
    ColTabRun* run = reinterpret_cast<ColTabRun*>(tb + ((size_t)ts << 3));
    double2* scs = reinterpret_cast<double2*>(run + n_cols);
    uint64_t* szout = reinterpret_cast<uint64_t*>(scs + n_ents);
    uint32_t* scsign = reinterpret_cast<uint32_t*>(szout + n_cols);
    __shared__ __align__(8) uint64_t s_mbar;
    uint32_t mphase = 0;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) mbar_init(&s_mbar, 1);
    for (int q = tid; q < n_cols; q += 256) {
        const DevCol co = cols[q];
        const DevColEntry e0 = ents[co.ent_begin];
        ColTabRun R;
        R.c = e0.c;
        R.s = e0.s;
        R.lxb = swz_idx8(co.lx, g.swz) << 3;
        R.items = co.n_active << co.free_log;
        R.tab_off = co.pad;
        R.ent0 = co.ent_begin | (co.n_active > 1u ? 0x80000000u : 0u);
        run[q] = R;
        szout[q] = co.zout;
    }
    for (int q = tid; q < n_ents; q += 256) scs[q] = make_double2(ents[q].c, ents[q].s);
    const BaseLane bl = base_lane_init(g);
    __syncthreads();
    // the item words of run 0 are the same for every tile
    uint32_t f0 = COLTAB_INVALID, f1 = COLTAB_INVALID;
    if (n_cols > 0) {
        const uint32_t it0 = run[0].items, o0 = run[0].tab_off;
        if (tid < it0) f0 = __ldg(tab + o0 + tid);
        if (tid + 256u < it0) f1 = __ldg(tab + o0 + tid + 256u);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous pass has written the whole state
    for (uint64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        const uint64_t base = tile_base_warp(g, bl, t);
        const uint64_t sbase = base | g.sign_base;
        __syncthreads();  // the store of the previous tile has read shared memory (thread 0 waited for it)
        if (tid == 0) tma_load_tile(tile, &tmap, g, base, &s_mbar);
        for (int r = tid; r < n_cols; r += 256) scsign[r] = (uint32_t)__popcll(sbase & szout[r]) & 1u;
        __syncthreads();  // scsign visible
        if (!mbar_wait(&s_mbar, mphase) && err) *err = 2;
        mphase ^= 1u;
        uint32_t w0 = f0, w1 = f1;
        const int nq = skeleton ? 0 : n_cols;
        for (int q = 0; q < nq; ++q) {
            const ColTabRun R = run[q];
            const uint32_t cs = scsign[q];
            double a0 = 0.0, b0 = 0.0, a1 = 0.0, b1 = 0.0;
            if (w0 != COLTAB_INVALID) {
                a0 = *reinterpret_cast<const double*>(tb + (w0 & 0xffffu));
                b0 = *reinterpret_cast<const double*>(tb + ((w0 & 0xffffu) ^ R.lxb));
            }
            if (w1 != COLTAB_INVALID) {
                a1 = *reinterpret_cast<const double*>(tb + (w1 & 0xffffu));
                b1 = *reinterpret_cast<const double*>(tb + ((w1 & 0xffffu) ^ R.lxb));
            }
            uint32_t n0 = COLTAB_INVALID, n1 = COLTAB_INVALID;
            if (q + 1 < n_cols) {  // next run's words: in flight across the arithmetic and the barrier
                const uint32_t itn = run[q + 1].items, on = run[q + 1].tab_off;
                if (tid < itn) n0 = __ldg(tab + on + tid);
                if (tid + 256u < itn) n1 = __ldg(tab + on + tid + 256u);
            }
            if (w0 != COLTAB_INVALID) coltab_rotate(tb, w0, R, scs, cs, a0, b0);
            if (w1 != COLTAB_INVALID) coltab_rotate(tb, w1, R, scs, cs, a1, b1);
            // items beyond two per thread (JW singles: 8 per thread in a 13-bit tile; tabulated plane rotations): four at a time
            for (uint32_t it = tid + 512u; it < R.items; it += 1024u) {
                uint32_t w[4];
                double a[4], b[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) w[k] = (it + 256u * k < R.items) ? __ldg(tab + R.tab_off + it + 256u * k) : COLTAB_INVALID;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (w[k] != COLTAB_INVALID) {
                        a[k] = *reinterpret_cast<const double*>(tb + (w[k] & 0xffffu));
                        b[k] = *reinterpret_cast<const double*>(tb + ((w[k] & 0xffffu) ^ R.lxb));
                    }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (w[k] != COLTAB_INVALID) coltab_rotate(tb, w[k], R, scs, cs, a[k], b[k]);
            }
            w0 = n0;
            w1 = n1;
            if (q + 1 == n_cols) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my writes -> the async-proxy store
            __syncthreads();
        }
        if (nq == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
        }
        if (tid == 0) {
            tma_store_tile(tile, &tmap, g, base);
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores have landed
}

// ------------------------------------------------------------------------------------------
// k_col_stab: k_col_tab with the pass's item table in SHARED memory, 16 bits per item: (element index in the swizzled tile)
// << 3 | parity(l & lz), i.e. the byte offset with the sign in bit 0; 0xffff = no item.  Global-memory words cost an L2 round
// trip per run while the L2 is saturated by the streaming tiles; a copy next to the tile costs two LDS.U16 per item.
// FACTORED: the index deposit, the swizzle and the parity are all linear over XOR, so the word of slot s of a segment is
// LANE[s & 31] ^ GROUP[s >> 5] (the pattern and the segment's own index bits folded into GROUP): 48 words = 96 bytes per
// segment instead of 512 words, so every pass keeps three CTAs per SM.
// The table is cut into SEGMENTS of 512 slots: every active pattern of a run is ceil(items / 512) segments (a JW double in
// a 13-bit tile: exactly one, a JW single: four), each with its own 32-byte descriptor (cos, sin, X offset, flags); only the
// last segment of a run ends with the CTA barrier -- the items of a run touch disjoint pairs.  A thread's slot of segment
// e + 1 is fetched before the barrier of segment e, so the chain of a run is  pair load -> 4 FP64 operations -> pair store
// -> barrier.  Shared memory is addressed through 32-bit shared-window addresses (ld.shared / st.shared).
// Passes whose table does not fit next to the tile keep k_tile_col.
// ------------------------------------------------------------------------------------------
#define COLSEG 512u
#define COLFACT 48u      // 16-bit words of a segment's factored table: 32 lane parts + 16 group parts
struct ColSub {            // 32 bytes, shared memory, one per segment
    double c, s;
    uint32_t lxb;          // X-mask as a BYTE offset in the tile's shared-memory layout
    uint32_t meta;         // bit 0: outside-tile Z parity of the tile at hand | bit 16: last segment of its run (barrier) | bit 17: all 512 slots hold items
    uint32_t zsel;         // which of the run's outside-tile Z masks (index into the per-segment array)
    uint32_t pad;
};
static_assert(sizeof(ColSub) == 32, "ColSub layout");
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
// slot word = lane part ^ group part (0xffff in either: the slot is empty)
__device__ __forceinline__ uint32_t col_slot(uint32_t lane_addr, uint32_t grp_addr) {
    const uint32_t wl = lds_u16(lane_addr), wg = lds_u16(grp_addr);
    return (wl == 0xffffu || wg == 0xffffu) ? 0xffffu : (wl ^ wg);
}
// THREADS = 512: one slot per thread and segment; 256: two.
template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_col_stab(const __grid_constant__ CUtensorMap tmap, TileGeom g, const DevCol* __restrict__ cols,
                                                     int n_cols, const DevColEntry* __restrict__ ents, int n_seg,
                                                     const uint16_t* __restrict__ tab, int skeleton, int* __restrict__ err) {
    constexpr int PF = (int)COLSEG / THREADS;   // slots per thread and segment
    extern __shared__ __align__(1024) double2 tile[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t ts = 1u << g.tbits;
    char* tb = reinterpret_cast<char*>(tile);
    ColSub* sub = reinterpret_cast<ColSub*>(tb + ((size_t)ts << 3));
    uint64_t* szout = reinterpret_cast<uint64_t*>(sub + n_seg);
    uint16_t* stab = reinterpret_cast<uint16_t*>(szout + n_seg);   // n_seg * 40 bytes after the tile: 8-byte aligned; the copy below needs 16
    stab = reinterpret_cast<uint16_t*>((reinterpret_cast<uintptr_t>(stab) + 15u) & ~(uintptr_t)15u);   // [n_seg][COLFACT]
    __shared__ __align__(8) uint64_t s_mbar;
    uint32_t mphase = 0;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) mbar_init(&s_mbar, 1);
    for (int q = tid; q < n_cols; q += THREADS) {
        const DevCol co = cols[q];
        const uint32_t lxb = swz_idx8(co.lx, g.swz) << 3;
        const uint32_t per_pat = max(1u, (1u << co.free_log) / COLSEG);
        const bool full = (1u << co.free_log) >= COLSEG;
        uint32_t sg = __ldg(reinterpret_cast<const uint32_t*>(tab) + q);   // first segment of the run in the pass (table header)
        for (uint32_t pi = 0; pi < co.n_active; ++pi) {
            const DevColEntry en = ents[co.ent_begin + pi];
            for (uint32_t k = 0; k < per_pat; ++k, ++sg) {
                ColSub S;
                S.c = en.c;
                S.s = en.s;
                S.lxb = lxb;
                S.meta = ((pi + 1u == co.n_active && k + 1u == per_pat) ? 0x10000u : 0u) | (full ? 0x20000u : 0u);
                S.zsel = 0;
                S.pad = 0;
                sub[sg] = S;
                szout[sg] = co.zout;
            }
        }
    }
    {   // the table: 16-byte copies (a segment is 96 bytes)
        const uint4* src = reinterpret_cast<const uint4*>(tab + ((2 * n_cols + 7) & ~7));
        uint4* dst = reinterpret_cast<uint4*>(stab);
        const int n16 = n_seg * (int)(COLFACT / 8u);
        for (int k = tid; k < n16; k += THREADS) dst[k] = __ldg(src + k);
    }
    const BaseLane bl = base_lane_init(g);
    // a thread's slot k of a segment is  s = tid + THREADS * k:  lane part s & 31, group part s >> 5
    const uint32_t tile32 = smem_u32(tile), sub32 = smem_u32(sub), lane32 = smem_u32(stab) + 2u * (tid & 31u),
                   grp32 = smem_u32(stab) + 64u + 2u * (tid >> 5);
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous pass has written the whole state
    const int ns = skeleton ? 0 : n_seg;
    for (uint64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        const uint64_t base = tile_base_warp(g, bl, t);
        const uint64_t sbase = base | g.sign_base;
        __syncthreads();  // the store of the previous tile has read shared memory (thread 0 waited for it)
        if (tid == 0) tma_load_tile(tile, &tmap, g, base, &s_mbar);
        for (int r = tid; r < n_seg; r += THREADS) sub[r].meta = (sub[r].meta & ~1u) | ((uint32_t)__popcll(sbase & szout[r]) & 1u);
        uint32_t w[PF];
#pragma unroll
        for (int k = 0; k < PF; ++k) w[k] = ns > 0 ? col_slot(lane32, grp32 + (THREADS / 16) * k) : 0xffffu;
        __syncthreads();  // signs visible
        if (!mbar_wait(&s_mbar, mphase) && err) *err = 2;
        mphase ^= 1u;
        for (int e = 0; e < ns; ++e) {
            const uint4 d0 = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(sub) + 32 * e);
            const uint2 d1 = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(sub) + 32 * e + 16);
            const double c = __hiloint2double((int)d0.y, (int)d0.x);
            const uint32_t lxb = d1.x, meta = d1.y;
            double a[PF], b[PF];
            uint32_t off[PF];
            if (meta & 0x20000u) {  // every slot of the segment holds an item (uniform): no per-thread checks
#pragma unroll
                for (int k = 0; k < PF; ++k) {
                    off[k] = w[k] & 0xfff8u;   // relative to the tile: the partner is an XOR of the RELATIVE offset
                    a[k] = lds_f64(tile32 + off[k]);
                    b[k] = lds_f64(tile32 + (off[k] ^ lxb));
                }
                uint32_t nw[PF];
#pragma unroll
                for (int k = 0; k < PF; ++k)
                    nw[k] = e + 1 < ns ? col_slot(lane32 + 2u * COLFACT * (uint32_t)(e + 1), grp32 + 2u * COLFACT * (uint32_t)(e + 1) + (THREADS / 16) * k) : 0xffffu;
#pragma unroll
                for (int k = 0; k < PF; ++k) {
                    const double sn = __hiloint2double((int)(d0.w ^ ((w[k] ^ meta) << 31)), (int)d0.z);
                    sts_f64(tile32 + off[k], fma(c, a[k], -sn * b[k]));
                    sts_f64(tile32 + (off[k] ^ lxb), fma(c, b[k], sn * a[k]));
                }
#pragma unroll
                for (int k = 0; k < PF; ++k) w[k] = nw[k];
            } else {
#pragma unroll
                for (int k = 0; k < PF; ++k) {
                    off[k] = w[k] & 0xfff8u;
                    a[k] = b[k] = 0.0;
                    if (w[k] != 0xffffu) {
                        a[k] = lds_f64(tile32 + off[k]);
                        b[k] = lds_f64(tile32 + (off[k] ^ lxb));
                    }
                }
                uint32_t nw[PF];
#pragma unroll
                for (int k = 0; k < PF; ++k)
                    nw[k] = e + 1 < ns ? col_slot(lane32 + 2u * COLFACT * (uint32_t)(e + 1), grp32 + 2u * COLFACT * (uint32_t)(e + 1) + (THREADS / 16) * k) : 0xffffu;
#pragma unroll
                for (int k = 0; k < PF; ++k)
                    if (w[k] != 0xffffu) {
                        const double sn = __hiloint2double((int)(d0.w ^ ((w[k] ^ meta) << 31)), (int)d0.z);
                        sts_f64(tile32 + off[k], fma(c, a[k], -sn * b[k]));
                        sts_f64(tile32 + (off[k] ^ lxb), fma(c, b[k], sn * a[k]));
                    }
#pragma unroll
                for (int k = 0; k < PF; ++k) w[k] = nw[k];
            }
            if (e + 1 == ns) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my writes -> the async-proxy store
            if (meta & 0x10000u) __syncthreads();
        }
        if (ns == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
        }
        if (tid == 0) {
            tma_store_tile(tile, &tmap, g, base);
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores have landed
    (void)sub32;
}

// ------------------------------------------------------------------------------------------
// Tile ring.  One persistent CTA per SM owns NSLOT shared-memory tile slots (3 x 64 KiB).  Warp 0 is the copy issuer:
// bulk loads (cp.async.bulk + mbarrier complete_tx) run two tiles ahead of the arithmetic and the bulk store of tile k
// drains while tile k+1 is being worked on, so HBM traffic and arithmetic overlap by construction instead of relying on
// several resident CTAs happening to be in different phases.
//   iteration k (slot k % NSLOT):  wait full[slot] -> arithmetic on the slot -> (store: fence, barrier, warp 0 issues
//   the bulk store and commits) -> warp 0 waits until the store of tile k-1 has READ its slot and refills that slot
//   with tile k+2.
// ------------------------------------------------------------------------------------------
#define NSLOT 3
__device__ __forceinline__ void ring_issue_load(double2* slot, const Shards& src, const TileGeom& g, uint64_t base, uint64_t* bar,
                                                const CUtensorMap* tm) {
    // called by all 32 lanes of warp 0
    if (g.tma) {
        if ((threadIdx.x & 31u) == 0) tma_load_tile(slot, tm, g, base, bar);
        return;
    }
    const uint32_t ts = 1u << g.tbits;
    const uint32_t nseg = ts >> g.lbits;
    const uint32_t seg_bytes = 16u << g.lbits;
    const uint32_t lane = threadIdx.x & 31u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic accesses to the slot come first
    if (lane == 0) mbar_arrive_expect_tx(bar, ts * 16u);
    for (uint32_t sgm = lane; sgm < nseg; sgm += 32u) {
        const uint32_t k = sgm << g.lbits;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(slot + k)),
                     "l"(amp_addr(g, src, base, k)), "r"(seg_bytes), "r"(smem_u32(bar))
                     : "memory");
    }
}
__device__ __forceinline__ void ring_issue_store(const double2* slot, const Shards& dst, const TileGeom& g, uint64_t base,
                                                 const CUtensorMap* tm) {
    // called by all 32 lanes of warp 0, after every writer has executed fence.proxy.async and the CTA barrier
    if (g.tma) {
        if ((threadIdx.x & 31u) == 0) tma_store_tile(slot, tm, g, base);
        return;
    }
    const uint32_t nseg = (1u << g.tbits) >> g.lbits;
    const uint32_t seg_bytes = 16u << g.lbits;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t sgm = lane; sgm < nseg; sgm += 32u) {
        const uint32_t k = sgm << g.lbits;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(amp_addr(g, dst, base, k)),
                     "r"(smem_u32(slot + k)), "r"(seg_bytes)
                     : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Collapsed-run passes on the tile ring (see k_tile_col for the per-run arithmetic).  blockDim = 256 when every run has
// at most 256 items per tile (JW doubles), 512 otherwise.
template <bool REAL>
__global__ void __launch_bounds__(512, 1) k_col_pipe(const __grid_constant__ CUtensorMap tmap, Shards psi, TileGeom g,
                                                     const DevCol* __restrict__ cols, int n_cols,
                                                     const DevColEntry* __restrict__ ents, int n_ents, int* __restrict__ err) {
    extern __shared__ __align__(1024) double2 smem_tiles[];
    const uint32_t ts = 1u << g.tbits;
    ColLite* lite = (ColLite*)(smem_tiles + (size_t)NSLOT * ts);
    DevColEntry* sent = (DevColEntry*)(lite + n_cols);
    uint64_t* szout = (uint64_t*)(sent + n_ents);
    uint32_t* scsign = (uint32_t*)(szout + n_cols);  // [2][n_cols]: outside-tile Z parity of every run, double-buffered per tile
    __shared__ __align__(8) uint64_t s_full[NSLOT];
    const uint32_t warp = threadIdx.x >> 5, bd = blockDim.x;
    if (threadIdx.x == 0)
        for (int i = 0; i < NSLOT; ++i) mbar_init(&s_full[i], 1);
    col_lite_build(lite, cols, ents, n_cols, g.swz);
    for (int q = threadIdx.x; q < n_cols; q += bd) szout[q] = cols[q].zout;
    for (int q = threadIdx.x; q < n_ents; q += bd) sent[q] = ents[q];
    const BaseLane bl = base_lane_init(g);
    const uint64_t n_my = g.n_tiles > blockIdx.x ? (g.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    __syncthreads();  // barriers initialised, tables in place
    if (warp == 0)
        for (uint64_t k = 0; k < 2 && k < n_my; ++k)
            ring_issue_load(smem_tiles + (size_t)(k % NSLOT) * ts, psi, g, tile_base_warp(g, bl, blockIdx.x + k * gridDim.x), &s_full[k % NSLOT], &tmap);
    for (uint64_t k = 0; k < n_my; ++k) {
        const uint32_t sl = (uint32_t)(k % NSLOT);
        double2* tile = smem_tiles + (size_t)sl * ts;
        const uint64_t base = tile_base_warp(g, bl, blockIdx.x + k * gridDim.x);
        const uint64_t sbase = base | g.sign_base;
        uint32_t* csign = scsign + (k & 1u) * n_cols;
        for (int r = threadIdx.x; r < n_cols; r += bd) csign[r] = (uint32_t)__popcll(sbase & szout[r]) & 1u;
        __syncthreads();  // signs of this tile visible (the other half of the double buffer may still be read by stragglers)
        const ColItem first = col_prep(lite, sent, csign, 0, n_cols, threadIdx.x, g.swz);
        if (!mbar_wait(&s_full[sl], (uint32_t)((k / NSLOT) & 1u)) && err) *err = 2;
        col_runs<REAL>(tile, lite, sent, csign, n_cols, g.swz, true, first);
        if (n_cols == 0) __syncthreads();  // (VQE_DEBUG_SKELETON: copy skeleton only)
        if (warp == 0) {
            ring_issue_store(tile, psi, g, base, &tmap);
            if (k + 2 < n_my) {
                // the slot of tile k-1 is refilled with tile k+2 once its store has read shared memory
                const uint64_t base2 = tile_base_warp(g, bl, blockIdx.x + (k + 2) * gridDim.x);
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                ring_issue_load(smem_tiles + (size_t)((k + 2) % NSLOT) * ts, psi, g, base2, &s_full[(k + 2) % NSLOT], &tmap);
            }
        }
    }
    if (warp == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all bulk stores have landed
}

__global__ void __launch_bounds__(512, 2) k_tile_ops(Shards psi, TileGeom g,
                                                  const DevOp* __restrict__ ops, int n_ops,
                                                  const double* __restrict__ mats) {
    extern __shared__ __align__(1024) double2 tile[];
    const uint32_t ts = 1u << g.tbits;
    const uint32_t half = ts >> 1;
    FastOp* optab = (FastOp*)(tile + ts);  // n_ops entries (host caps n_ops per pass at OPTAB_CAP)
    for (uint64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        const uint64_t base = tile_base(g, t);
        const uint64_t sbase = base | g.sign_base;
        // per-tile table of the fast rotations: tangent with the sign of the outside-tile Z parity folded in
        for (int r = threadIdx.x; r < n_ops; r += blockDim.x) {
            const DevOp o = ops[r];
            FastOp f;
            f.t = flipsign(o.s, __popcll(sbase & o.zout));
            f.lz = o.lz;
            f.meta = o.jmask | (o.imag << 31);
            optab[r] = f;
        }
        tile_load(tile, psi, g, base);
        int i = 0;
        while (i < n_ops) {
            __syncthreads();
            const uint32_t kind = ops[i].kind;
            if (kind == OP_ROTF) {
                // Tangent form (see rot_fast_run): one run = consecutive rotations sharing lx and phase type
                if (ops[i].imag) rot_fast_run<true>(tile, optab + i, ops[i].lx, ops[i].hb, (int)ops[i].run, ops[i].c, half);
                else rot_fast_run<false>(tile, optab + i, ops[i].lx, ops[i].hb, (int)ops[i].run, ops[i].c, half);
                i += (int)ops[i].run;
            } else if (kind == OP_ROT) {
                slow_rot(tile, ops, i, sbase, ts);
                i += (int)ops[i].run;
            } else if (kind == OP_GATE1) {
                slow_gate1(tile, ops, i, mats, ts);
                i += 1;
            } else {  // OP_CNOT
                slow_cnot(tile, ops, i, sbase, ts);
                i += 1;
            }
        }
        __syncthreads();
        tile_store(tile, psi, g, base);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// <psi| O |psi> for the X-mask groups of one pass.  grid = (tile workers, group chunks)
// ------------------------------------------------------------------------------------------
#define TERM_CAP 384    // terms per pass (host splits passes accordingly)
#define GROUP_CAP 80    // X-mask groups per pass
#define GCOL_ENT_CAP 128  // (pattern, weight) entries of collapsed groups per pass

// COLLAPSED group of an expectation pass (the analogue of DevCol): the Z-variants of an X-mask group differ only on
// a few tile bits D, so sum_t c_t (-1)^parity(l & z_t) = (-1)^parity(l & z_1) * F[l restricted to D] with F tabulated
// by the host.  For the two-body terms of a molecular Hamiltonian F vanishes on most occupation patterns, so only
// the ACTIVE (pattern, free index) pairs are visited: one pair product and one DFMA each.
struct DevGColEntry {  // 32 bytes
    double fr, fi;     // weights of 2 Re(conj(b) a) and 2 Im(conj(b) a)
    uint32_t pat;      // pattern as tile-index bits (a-side)
    uint32_t pad[3];
};
// Flat work list of a pass: every collapsed group with at least 256 free indices per pattern is cut into entries of
// exactly 256 (pattern, free index) pairs, and the threads of a CTA walk the concatenated list -- no per-group
// loop, no idle threads when a group has fewer pairs than the CTA has threads.
struct DevFlat {       // 32 bytes
    uint32_t lx, lz;   // X-mask, Z letters of the group's first term inside the tile
    uint32_t pat;      // pattern bits | free-index bits above the low 8 (already deposited)
    uint32_t zsel;     // bits 0-15: index of the entry's outside-tile Z mask in the pass's zout table;
                       // bits 16-23 / 24-31: tile-bit position of free-index bit 6 / 7 after the deposit
    uint16_t himask[4];  // ~((1 << pos) - 1) for the (up to 4) fixed positions, ascending, 0 when unused:
                         // insert0(l, pos) = l + (l & himask)
    double fr;         // weight of Re(conj(b) a), factor 2 included
};
#define FLAT_CAP 320   // entries per pass (15 KiB)
struct DevGCol {       // 64 bytes
    uint64_t zout;     // Z letters outside the tile (same for every term of the group)
    uint32_t lz;       // Z letters of the first term inside the tile
    uint32_t n_active, nd;
    uint32_t dpos[6];
    uint32_t ent_begin, free_log;
    uint32_t pad[3];
};
#define EXP_PAIRS 4

// <psi|O|psi> for one pass.  Group headers and term tables are staged in shared memory ONCE per CTA; per tile
// only the outside-tile Z parity of each term is refreshed (signed coefficient table), the tile itself arrives
// through cp.async.
//   A thread holds 4 pair products w_j = 2 conj(b_j) a_j of a group (8 densities |a_j|^2 for the diagonal group).
//   The sign of term t on pair j is (-1)^parity(l_j & lz_t) = s_t(thread) * (-1)^(j . q_t), where q_t are the
//   strings' Z letters on the two (three) tile bits that enumerate the thread's pairs: only 4 (8) sign CLASSES
//   exist.  So the thread Walsh-Hadamard-transforms its pair products once per group, W_q = sum_j (-1)^(j.q) w_j,
//   the host sorts the terms of a group by class, and a term costs ONE DFMA per thread:
//       E += (c_t s_t) * W_{q_t}            (LDS coefficient + LDS lz + LOP + POPC + 2 integer ops + DFMA).
template <bool CPLX>
__global__ void __launch_bounds__(512, 2) k_tile_expect(Shards psi, TileGeom g,
                                                        const DevGroup* __restrict__ groups, int n_groups,
                                                        const DevTerm* __restrict__ terms,
                                                        const DevGCol* __restrict__ gcols, int n_gcols,
                                                        const DevGColEntry* __restrict__ gents, int n_gents,
                                                        const DevFlat* __restrict__ flats, int n_flats,
                                                        const uint64_t* __restrict__ fzout,
                                                        double2* __restrict__ partial, int* __restrict__ err) {
    extern __shared__ __align__(1024) double2 tile[];
    __shared__ double red[64];
    const uint32_t ts = 1u << g.tbits;
    const uint32_t half = ts >> 1;
    DevTerm* s_term = (DevTerm*)(tile + ts);               // TERM_CAP
    double2* s_sc = (double2*)(s_term + TERM_CAP);         // TERM_CAP signed coefficients (per tile)
    DevGroup* s_grp = (DevGroup*)(s_sc + TERM_CAP);        // GROUP_CAP
    DevGCol* s_gcol = (DevGCol*)(s_grp + GROUP_CAP);       // GROUP_CAP
    DevGColEntry* s_gent = (DevGColEntry*)(s_gcol + GROUP_CAP);  // GCOL_ENT_CAP
    DevFlat* s_flat = (DevFlat*)(s_gent + GCOL_ENT_CAP);         // FLAT_CAP
    double* s_ffr = (double*)(s_flat + FLAT_CAP);                // FLAT_CAP: per tile, weight with the outside-tile parity folded in
    for (int q = threadIdx.x; q < n_gcols; q += blockDim.x) s_gcol[q] = gcols[q];
    for (int q = threadIdx.x; q < n_gents; q += blockDim.x) s_gent[q] = gents[q];
    // this CTA's share of the flat list (blockIdx.y splits it like the groups)
    __shared__ uint64_t s_boff[16];
    __shared__ __align__(8) uint64_t s_mbar;
    uint32_t mphase = 0;
    if (g.bulk && threadIdx.x == 0) mbar_init(&s_mbar, 1);
    const TileAddr ta = tile_addr_init(g, s_boff);
    const int fper = (n_flats + gridDim.y - 1) / gridDim.y;
    const int f0 = min(n_flats, (int)blockIdx.y * fper), f1 = min(n_flats, f0 + fper);
    const int nfl = f1 - f0;
    for (int q = threadIdx.x; q < nfl; q += blockDim.x) s_flat[q] = flats[f0 + q];
    // group chunk of this CTA (blockIdx.y splits the groups of a pass when there are few tiles)
    const int per = (n_groups + gridDim.y - 1) / gridDim.y;
    const int g0 = blockIdx.y * per, g1 = min(n_groups, g0 + per);
    const int ng = max(0, g1 - g0);
    uint32_t tb0 = 0, nt = 0;
    if (ng > 0) {
        tb0 = groups[g0].t_begin;
        const DevGroup last = groups[g1 - 1];
        nt = last.t_begin + last.n_even + last.n_odd - tb0;
    }
    for (int q = threadIdx.x; q < ng; q += blockDim.x) s_grp[q] = groups[g0 + q];
    for (uint32_t k = threadIdx.x; k < nt; k += blockDim.x) s_term[k] = terms[tb0 + k];
    double er = 0.0, ei = 0.0;
    const uint32_t bd = blockDim.x;
#define EXP_TERM(W)                                                                                        \
    {                                                                                                      \
        const uint32_t sg = (uint32_t)__popc(l0 & s_term[k].lz) << 31;                                      \
        if (CPLX) {                                                                                        \
            const double2 c = s_sc[k];                                                                     \
            er = fma(__hiloint2double((int)((uint32_t)__double2hiint(c.x) ^ sg), __double2loint(c.x)), W, er); \
            ei = fma(__hiloint2double((int)((uint32_t)__double2hiint(c.y) ^ sg), __double2loint(c.y)), W, ei); \
        } else {                                                                                           \
            const double c = s_sc[k].x;                                                                    \
            er = fma(__hiloint2double((int)((uint32_t)__double2hiint(c) ^ sg), __double2loint(c)), W, er);   \
        }                                                                                                  \
        ++k;                                                                                               \
    }
    for (uint64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        const uint64_t base = tile_base(g, t);
        const uint64_t sbase = base | g.sign_base;
        __syncthreads();  // previous tile fully consumed (and, first time, the tables are in place)
        if (g.bulk) tile_load_bulk(tile, psi, g, base, &s_mbar);
        else tile_load_async_fast(tile, psi, g, ta, s_boff, base);
        for (uint32_t k = threadIdx.x; k < nt; k += blockDim.x) {
            const uint32_t par = __popcll(sbase & s_term[k].zout);
            s_sc[k] = make_double2(flipsign(s_term[k].ar, par), flipsign(s_term[k].ai, par));
        }
        for (int q = threadIdx.x; q < nfl; q += blockDim.x)
            s_ffr[q] = flipsign(s_flat[q].fr, __popcll(sbase & fzout[s_flat[q].zsel & 0xffffu]));
        if (g.bulk) {
            if (!mbar_wait(&s_mbar, mphase) && err) *err = 2;
            mphase ^= 1u;
        } else {
            cp_async_wait_all();
        }
        __syncthreads();
        // flat list of the collapsed groups: 256 (pattern, free index) pairs per entry
        // a work unit = 4 pairs of one entry (free-index bits 6 and 7 enumerate them): the entry is read and the low
        // six free bits are deposited once per unit
        for (uint32_t u = threadIdx.x; u < ((uint32_t)nfl << 6); u += bd) {
            const DevFlat& fe = s_flat[u >> 6];
            const uint2 hm = *reinterpret_cast<const uint2*>(fe.himask);
            uint32_t l = u & 63u;
            l += l & (hm.x & 0xffffu);
            l += l & (hm.x >> 16);
            l += l & (hm.y & 0xffffu);
            l += l & (hm.y >> 16);
            l |= fe.pat;
            const uint32_t lx = fe.lx, lz = fe.lz;
            const uint32_t b6 = 1u << ((fe.zsel >> 16) & 0xffu), b7 = 1u << (fe.zsel >> 24);
            const double fr = s_ffr[u >> 6];
            double part = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t lj = l | ((j & 1) ? b6 : 0u) | ((j & 2) ? b7 : 0u);
                const double2 a = tile[lj], b = tile[lj ^ lx];
                const double w = fma(b.x, a.x, b.y * a.y);  // Re(conj(b) a)
                part += flipsign(w, __popc(lj & lz));
            }
            er = fma(fr, part, er);
        }
        for (int q = 0; q < ng; ++q) {
            const DevGroup& G = s_grp[q];
            if (G.pad == 0xffffffffu) continue;  // in the flat list
            const uint32_t lx = G.lx, hb = G.hb;
            uint32_t k = G.t_begin - tb0;
            if (G.pad != 0) {
                // collapsed group: only the active (pattern, free index) pairs
                const DevGCol& co = s_gcol[G.pad - 1];
                const uint32_t items = co.n_active << co.free_log;
                const uint32_t fmask = (1u << co.free_log) - 1u;
                const uint32_t tsig = (uint32_t)__popcll(sbase & co.zout);
                for (uint32_t it = threadIdx.x; it < items; it += bd) {
                    const DevGColEntry en = s_gent[co.ent_begin + (it >> co.free_log)];
                    uint32_t l = it & fmask;
#pragma unroll
                    for (int d = 0; d < 6; ++d)
                        if (d < (int)co.nd) l = insert0(l, co.dpos[d]);
                    l |= en.pat;
                    const double2 a = tile[l], b = tile[l ^ lx];
                    const double wr = 2.0 * fma(b.x, a.x, b.y * a.y);
                    const double wi = 2.0 * fma(b.x, a.y, -b.y * a.x);
                    er += flipsign(fma(en.fr, wr, en.fi * wi), tsig + (uint32_t)__popc(l & co.lz));
                }
            } else if (lx == 0) {
                // diagonal group: 8 densities per thread, 8 sign classes
                const uint32_t kstart = k;
                for (uint32_t p0 = threadIdx.x; p0 < ts; p0 += 8 * bd) {
                    double v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t l = p0 + j * bd;
                        v[j] = 0.0;
                        if (l < ts) {
                            const double2 a = tile[l];
                            v[j] = fma(a.x, a.x, a.y * a.y);
                        }
                    }
#pragma unroll
                    for (int st = 1; st < 8; st <<= 1)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (!(j & st)) {
                                const double u = v[j], w = v[j | st];
                                v[j] = u + w;
                                v[j | st] = u - w;
                            }
                    const uint32_t l0 = p0;
                    k = kstart;
#pragma unroll
                    for (int cls = 0; cls < 8; ++cls) {
                        const uint32_t n = G.cnt[cls];
                        for (uint32_t e = 0; e < n; ++e) EXP_TERM(v[cls])
                    }
                }
            } else {
                const uint32_t kstart = k;
                const bool any_odd = G.n_odd != 0;
                for (uint32_t p0 = threadIdx.x; p0 < half; p0 += 4 * bd) {
                    double wr[4], wi[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t pidx = p0 + j * bd;
                        wr[j] = wi[j] = 0.0;
                        if (pidx < half) {
                            const uint32_t l = insert0(pidx, hb);
                            const double2 a = tile[l], b = tile[l ^ lx];
                            wr[j] = 2.0 * fma(b.x, a.x, b.y * a.y);   // 2 Re(conj(b) a)
                            wi[j] = 2.0 * fma(b.x, a.y, -b.y * a.x);  // 2 Im(conj(b) a)
                        }
                    }
#pragma unroll
                    for (int st = 1; st < 4; st <<= 1)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (!(j & st)) {
                                const double u = wr[j], w = wr[j | st];
                                wr[j] = u + w;
                                wr[j | st] = u - w;
                                const double ui = wi[j], wq = wi[j | st];
                                wi[j] = ui + wq;
                                wi[j | st] = ui - wq;
                            }
                    const uint32_t l0 = insert0(p0, hb);
                    k = kstart;
#pragma unroll
                    for (int cls = 0; cls < 4; ++cls) {
                        const uint32_t n = G.cnt[cls];
                        for (uint32_t e = 0; e < n; ++e) EXP_TERM(wr[cls])
                    }
                    if (any_odd) {
#pragma unroll
                        for (int cls = 0; cls < 4; ++cls) {
                            const uint32_t n = G.cnt[4 + cls];
                            for (uint32_t e = 0; e < n; ++e) EXP_TERM(wi[cls])
                        }
                    }
                }
            }
        }
    }
#undef EXP_TERM
    double2 sres = block_sum2(er, ei, red);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = sres;
}

// ------------------------------------------------------------------------------------------
// dst (+)= O src for the groups of one pass.  One CTA per tile (persistent over tiles).
// ------------------------------------------------------------------------------------------
// sigma (+)= O psi for the groups of one pass.  One CTA of up to 1024 threads per SM holds the source tile AND an
// accumulator tile in shared memory and walks the groups one after the other (a barrier between two groups: inside a
// group every destination element is written by exactly one thread, so the accumulation is conflict free and its
// order is fixed -> bit-reproducible).
//   Collapsed groups (Z-variants differ on <= 4 tile bits): the host tabulates the complex coupling F per source
//   occupation pattern and lists only the patterns with F != 0, cut into entries of 256 (pattern, free index) pairs:
//       acc[m ^ lx] += (-1)^parity(m & lz_1) * F * psi[m]           one visit per COUPLED source amplitude.
//   Other groups (diagonal group, groups dressed with many number operators): one thread per destination element,
//   loop over the group's terms with popcount signs.
struct DevAFlat {      // 48 bytes
    double fr, fi;     // coupling of this source pattern (c_k i^ny summed over the group's strings, relative signs in)
    uint32_t lx, lz;   // X-mask, Z letters of the group's first term inside the tile
    uint32_t pat;      // source pattern bits | free-index bits above the low 8 (already deposited)
    uint32_t zsel;     // index into the pass's table of outside-tile Z masks
    uint16_t himask[4];
    uint32_t pad[2];
};
#define AFLAT_CAP 1024
__global__ void __launch_bounds__(1024, 1) k_tile_apply(Shards src, Shards dst, TileGeom g, const DevGroup* __restrict__ groups,
                                                        int n_groups, const DevTerm* __restrict__ terms,
                                                        const DevAFlat* __restrict__ aflat, const uint32_t* __restrict__ aoff,
                                                        const uint64_t* __restrict__ azout, int accumulate) {
    extern __shared__ __align__(1024) double2 tile[];
    const uint32_t ts = 1u << g.tbits;
    double2* acc = tile + ts;
    double2* s_coef = acc + ts;                                  // TERM_CAP
    uint32_t* s_lz = (uint32_t*)(s_coef + TERM_CAP);             // TERM_CAP
    DevAFlat* s_fl = (DevAFlat*)(s_lz + TERM_CAP);               // AFLAT_CAP
    double2* s_ff = (double2*)(s_fl + AFLAT_CAP);                // AFLAT_CAP: per tile, F with the outside parity folded in
    uint32_t* s_off = (uint32_t*)(s_ff + AFLAT_CAP);             // GROUP_CAP + 1
    const uint32_t bd = blockDim.x;
    const uint32_t n_fl = aoff[n_groups];
    for (uint32_t q = threadIdx.x; q < n_fl; q += bd) s_fl[q] = aflat[q];
    for (uint32_t q = threadIdx.x; q <= (uint32_t)n_groups; q += bd) s_off[q] = aoff[q];
    for (uint64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        const uint64_t base = tile_base(g, t);
        const uint64_t sbase = base | g.sign_base;
        __syncthreads();  // previous tile stored (and, first time, the tables are in place)
        tile_load_async(tile, src, g, base);
        for (uint32_t k = threadIdx.x; k < ts; k += bd) acc[k] = make_double2(0.0, 0.0);
        for (uint32_t q = threadIdx.x; q < n_fl; q += bd) {
            const uint32_t par = __popcll(sbase & azout[s_fl[q].zsel]);
            s_ff[q] = make_double2(flipsign(s_fl[q].fr, par), flipsign(s_fl[q].fi, par));
        }
        cp_async_wait_all();
        __syncthreads();
        for (int q = 0; q < n_groups; ++q) {
            const uint32_t e0 = s_off[q], e1 = s_off[q + 1];
            if (e1 > e0) {
                const uint32_t items = (e1 - e0) << 8;
                for (uint32_t it = threadIdx.x; it < items; it += bd) {
                    const uint32_t ei = e0 + (it >> 8);
                    const DevAFlat& fe = s_fl[ei];
                    const uint2 hm = *reinterpret_cast<const uint2*>(fe.himask);
                    uint32_t m = it & 255u;
                    m += m & (hm.x & 0xffffu);
                    m += m & (hm.x >> 16);
                    m += m & (hm.y & 0xffffu);
                    m += m & (hm.y >> 16);
                    m |= fe.pat;
                    const double2 f = s_ff[ei];
                    const double2 v = tile[m];
                    const uint32_t sg = (uint32_t)__popc(m & fe.lz);
                    double2* a = acc + (m ^ fe.lx);
                    double2 o = *a;
                    o.x += flipsign(f.x * v.x - f.y * v.y, sg);
                    o.y += flipsign(f.x * v.y + f.y * v.x, sg);
                    *a = o;
                }
            } else {
                const DevGroup gr = groups[q];
                const uint32_t nk = gr.n_even + gr.n_odd;
                for (uint32_t k0 = 0; k0 < nk; k0 += TERM_CAP) {  // (groups are split at TERM_CAP terms by the host)
                    const uint32_t nn = min(nk - k0, (uint32_t)TERM_CAP);
                    __syncthreads();
                    for (uint32_t k = threadIdx.x; k < nn; k += bd) {
                        const DevTerm tm = terms[gr.t_begin + k0 + k];
                        const uint32_t par = __popcll(sbase & tm.zout) & 1u;
                        s_coef[k] = make_double2(flipsign(tm.ar, par), flipsign(tm.ai, par));
                        s_lz[k] = tm.lz;
                    }
                    __syncthreads();
                    for (uint32_t l = threadIdx.x; l < ts; l += bd) {
                        const uint32_t sidx = l ^ gr.lx;
                        double sr = 0.0, si = 0.0;
                        for (uint32_t k = 0; k < nn; ++k) {
                            const uint32_t par = __popc(sidx & s_lz[k]);
                            const double2 c = s_coef[k];
                            sr += flipsign(c.x, par);
                            si += flipsign(c.y, par);
                        }
                        const double2 v = tile[sidx];
                        double2 o = acc[l];
                        o.x += sr * v.x - si * v.y;
                        o.y += sr * v.y + si * v.x;
                        acc[l] = o;
                    }
                }
            }
            __syncthreads();
        }
        for (uint32_t k = threadIdx.x; k < ts; k += bd) {
            double2* dp = amp_addr(g, dst, base, k);
            double2 o = acc[k];
            if (accumulate) {
                const double2 d = *dp;
                o.x += d.x;
                o.y += d.y;
            }
            *dp = o;
        }
    }
}

// ------------------------------------------------------------------------------------------
// ADAPT pool sweep: out_k = <bra| A_k |ket>, one warp per pool operator, whole pool per launch.
// grid = (tile workers, operator chunks).  Both tiles are staged in shared memory.
// ------------------------------------------------------------------------------------------
struct DevPoolOp {
    uint32_t t_begin, n_terms;  // terms of this operator (pass-local list)
    uint32_t out_index;         // pool index
    uint32_t col;               // 0, or 1 + index of the operator's collapsed form (DevGCol with lx in pad[0]):
                                // all strings share one X-mask, their Z letters differ on <= 5 tile bits, so
                                // A_k psi at source index m is psi[m] * (-1)^parity(m & lz_1) * F[m restricted to D]
                                // and only the patterns with F != 0 are visited (2 of 16 for a JW double excitation)
};
struct DevPoolTerm {  // 32 bytes
    uint64_t zout;
    uint32_t lx, lz;
    double ar, ai;    // c_k * i^ny
};

__global__ void __launch_bounds__(512) k_tile_pool(Shards bra, Shards ket, TileGeom g, const DevPoolOp* __restrict__ pops, int n_pops,
                                                   const DevPoolTerm* __restrict__ terms,
                                                   const DevGCol* __restrict__ pcols, const DevGColEntry* __restrict__ pents,
                                                   double2* __restrict__ partial /* [gridDim.x][n_pops] */) {
    extern __shared__ __align__(1024) double2 tile[];
    const uint32_t ts = 1u << g.tbits;
    double2* tbra = tile;
    double2* tket = tile + ts;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int per = (n_pops + gridDim.y - 1) / gridDim.y;
    const int o0 = blockIdx.y * per, o1 = min(n_pops, o0 + per);
    const bool same = (bra.p0 == ket.p0);
    // zero this CTA's partial slots (accumulated over its tiles below)
    for (int o = o0 + threadIdx.x; o < o1; o += blockDim.x)
        partial[(size_t)blockIdx.x * n_pops + o] = make_double2(0.0, 0.0);
    for (uint64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        const uint64_t base = tile_base(g, t);
        const uint64_t sbase = base | g.sign_base;
        __syncthreads();
        tile_load(tket, ket, g, base);
        if (!same) tile_load(tbra, bra, g, base);
        __syncthreads();
        const double2* tb = same ? tket : tbra;
        for (int o = o0 + warp; o < o1; o += nwarps) {
            const DevPoolOp po = pops[o];
            double re = 0.0, im = 0.0;
            if (po.col) {
                const DevGCol co = pcols[po.col - 1];
                const uint32_t lx = co.pad[0];
                const uint32_t items = co.n_active << co.free_log;
                const uint32_t fmask = (1u << co.free_log) - 1u;
                const uint32_t tsig = (uint32_t)__popcll(sbase & co.zout);
                for (uint32_t it = lane; it < items; it += 32) {
                    const DevGColEntry en = pents[co.ent_begin + (it >> co.free_log)];
                    uint32_t m = it & fmask;
                    m += m & co.dpos[0];
                    m += m & co.dpos[1];
                    m += m & co.dpos[2];
                    m += m & co.dpos[3];
                    if (co.nd > 4) {
                        m += m & co.dpos[4];
                        m += m & co.dpos[5];
                    }
                    m |= en.pat;
                    const double2 v = tket[m];
                    const double2 b = tb[m ^ lx];
                    const double pr = b.x * v.x + b.y * v.y, pi = b.x * v.y - b.y * v.x;  // conj(b) * v
                    const uint32_t sg = tsig + (uint32_t)__popc(m & co.lz);
                    re += flipsign(en.fr * pr - en.fi * pi, sg);
                    im += flipsign(en.fr * pi + en.fi * pr, sg);
                }
            }
            for (uint32_t k = 0; k < po.n_terms; ++k) {
                const DevPoolTerm tm = terms[po.t_begin + k];
                const uint32_t opar = __popcll(sbase & tm.zout) & 1u;
                const double cr = flipsign(tm.ar, opar), ci = flipsign(tm.ai, opar);
                double tr = 0.0, ti = 0.0;
                for (uint32_t l = lane; l < ts; l += 32) {
                    const uint32_t sidx = l ^ tm.lx;
                    const uint32_t par = __popc(sidx & tm.lz);
                    const double2 v = tket[sidx];
                    const double2 b = tb[l];
                    // conj(b) * v, signed
                    tr += flipsign(b.x * v.x + b.y * v.y, par);
                    ti += flipsign(b.x * v.y - b.y * v.x, par);
                }
                re += cr * tr - ci * ti;
                im += cr * ti + ci * tr;
            }
            re = warp_sum(re);
            im = warp_sum(im);
            if (lane == 0) {
                double2* slot = partial + (size_t)blockIdx.x * n_pops + o;
                double2 cur = *slot;
                cur.x += re;
                cur.y += im;
                *slot = cur;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// LEAN Pauli-sum passes.  An X-mask group of a real Hamiltonian whose Z-variants differ from the first string only
//   (i) on the group's own X positions (XX/YY/XY families of one excitation), and
//   (ii) by at most ONE further Z letter (number-operator dressing n_r of a one-body term)
// is a function  F(l) = (-1)^parity(l & z_1) * [ beta_0(pi) + sum_r beta_r(pi) (-1)^(l_r) ]  of the occupation pattern pi
// on the X positions: the host tabulates it per a-side pattern, drops the patterns where it vanishes (most of them:
// particle-number and spin symmetry) and cuts every active pattern into ENTRIES of 256 (pattern, free index) pairs.
// Entries live in global memory (read through L1 with warp-uniform 16-byte loads), so a lean pass needs no
// shared-memory tables besides the tile: 3 CTAs of 256 threads per SM.
//   expectation: one WARP per entry, 8 pairs per lane: index deposit and sign work once per 8 pairs;
//   sigma = O psi: one THREAD per pair, groups separated by barriers (conflict-free, fixed accumulation order).
// Everything is kept in the byte-offset domain (tile index * 16 < 65536) so that an index is used as an address.
// ------------------------------------------------------------------------------------------
struct DevFlat2 {           // 48 bytes = 3 x uint4
    double fr;              // plain entry: 2 * beta_0(pattern); additive entry: 2.0 (the weight comes from the tables)
    uint16_t lx16, lz16;    // X-mask | Z letters of the group's first string inside the tile        (byte offsets)
    uint16_t pat16, zsel;   // a-side pattern (byte offset) | index into the pass's outside-tile Z masks
    uint16_t hm16[4];       // deposit masks of the X positions for the lane's five free-index bits (byte-offset domain)
    uint16_t jsign, tab;    // bit j = parity(o_j & z_1) | 0xffff: plain, else offset (doubles) of T_lo[32] in the table array
    uint16_t hi0, bidx;     // additive: offset of this chunk's T_hi[8] | index of the pattern's per-tile constant
    uint16_t o16[8];        // byte offset of free-index bits 5.. for j = 0..7 (chunk bits included)
};
static_assert(sizeof(DevFlat2) == 48, "DevFlat2 layout");
struct DevAddPat {          // 16 bytes: per-tile constant of an additive pattern
    double beta0;           // beta_0 + sum over outside-tile r of beta_r (-1)^(bit r of the tile base)
    uint32_t out_begin, out_count;
};
struct DevAddOut {          // 16 bytes
    double c;
    uint64_t zrel;          // outside-tile Z mask relative to the group's first string
};
struct LeanUnit {           // decoded entry, per lane
    uint32_t v;             // byte offset of the lane's a-side element for j = 0 (pattern included), NATURAL layout
    uint32_t lx16, s0, jsign, tab, hi0, bidx;
    uint32_t o[8];
    double fr;
};
// lane_shift: log2 of the tile element size (4: interleaved complex, 3: real layout -- the entry must then be the real-layout
// form produced by lean_entry_to_rl)
__host__ __device__ __forceinline__ void lean_decode(const uint4& q0, const uint4& q1, const uint4& q2, uint32_t lane, LeanUnit& u,
                                                     uint32_t lane_shift = 4) {
#ifdef __CUDA_ARCH__
    u.fr = __hiloint2double((int)q0.y, (int)q0.x);
#else
    uint64_t bits = ((uint64_t)q0.y << 32) | q0.x;
    memcpy(&u.fr, &bits, 8);
#endif
    u.lx16 = q0.z & 0xffffu;
    const uint32_t lz16 = q0.z >> 16, pat16 = q0.w & 0xffffu;
    uint32_t v = lane << lane_shift;
    v += v & (q1.x & 0xffffu);
    v += v & (q1.x >> 16);
    v += v & (q1.y & 0xffffu);
    v += v & (q1.y >> 16);
    v |= pat16;
    u.v = v;
#ifdef __CUDA_ARCH__
    u.s0 = (uint32_t)__popc(v & lz16);
#else
    u.s0 = (uint32_t)__builtin_popcount(v & lz16);
#endif
    u.jsign = q1.z & 0xffffu;
    u.tab = q1.z >> 16;
    u.hi0 = q1.w & 0xffffu;
    u.bidx = q1.w >> 16;
    u.o[0] = q2.x & 0xffffu; u.o[1] = q2.x >> 16;
    u.o[2] = q2.y & 0xffffu; u.o[3] = q2.y >> 16;
    u.o[4] = q2.z & 0xffffu; u.o[5] = q2.z >> 16;
    u.o[6] = q2.w & 0xffffu; u.o[7] = q2.w >> 16;
}

// the fields of an entry that do not depend on the lane (the lane part then comes from a lane table, see PSPass::rlp)
__host__ __device__ __forceinline__ void lean_decode_uniform(const uint4& q0, const uint4& q1, const uint4& q2, LeanUnit& u) {
#ifdef __CUDA_ARCH__
    u.fr = __hiloint2double((int)q0.y, (int)q0.x);
#else
    uint64_t bits = ((uint64_t)q0.y << 32) | q0.x;
    memcpy(&u.fr, &bits, 8);
#endif
    u.lx16 = q0.z & 0xffffu;
    u.v = 0;
    u.s0 = 0;
    u.jsign = q1.z & 0xffffu;
    u.tab = q1.z >> 16;
    u.hi0 = q1.w & 0xffffu;
    u.bidx = q1.w >> 16;
    u.o[0] = q2.x & 0xffffu; u.o[1] = q2.x >> 16;
    u.o[2] = q2.y & 0xffffu; u.o[3] = q2.y >> 16;
    u.o[4] = q2.z & 0xffffu; u.o[5] = q2.z >> 16;
    u.o[6] = q2.w & 0xffffu; u.o[7] = q2.w >> 16;
}

// The real-layout form of an entry: every byte offset is halved (8-byte tile elements) and re-swizzled for the layout the
// real-layout tile has in shared memory (swz_c / swz_r: the complex / real tile is 128-byte swizzled).
static inline DevFlat2 lean_entry_to_rl(const DevFlat2& f, bool swz_c, bool swz_r) {
    auto conv = [&](uint16_t off16) -> uint16_t {
        const uint32_t nat = swz_c ? swz_off(off16, 0x70u) : off16;  // the swizzle is an involution
        const uint32_t half = nat >> 1;
        return (uint16_t)(swz_r ? swz_off(half, 0x70u) : half);
    };
    DevFlat2 r = f;
    r.lx16 = conv(f.lx16);
    r.lz16 = (uint16_t)(f.lz16 >> 1);    // natural-layout masks
    r.pat16 = (uint16_t)(f.pat16 >> 1);
    for (int k = 0; k < 4; ++k) r.hm16[k] = f.hm16[k] ? (uint16_t)((f.hm16[k] >> 1) | 0x8000u) : 0;
    for (int j = 0; j < 8; ++j) r.o16[j] = conv(f.o16[j]);
    return r;
}

// per-tile constants of the additive patterns (one thread per pattern, while the tile load is in flight)
__device__ __forceinline__ void lean_betas(double* s_beta, const DevAddPat* __restrict__ addpat, int n_addpat,
                                           const DevAddOut* __restrict__ addout, uint64_t sbase) {
    for (int k = threadIdx.x; k < n_addpat; k += blockDim.x) {
        const DevAddPat ap = addpat[k];
        double b = ap.beta0;
        for (uint32_t q = 0; q < ap.out_count; ++q) {
            const DevAddOut ao = addout[ap.out_begin + q];
            b += flipsign(ao.c, (uint32_t)__popcll(sbase & ao.zrel));
        }
        s_beta[k] = b;
    }
}

// Entries e0, e0 + stride, ... < e1 of one warp on one tile.  The descriptor (3 x 16 bytes through L1) and the
// outside-tile Z mask of entry k+1 are fetched while entry k is evaluated.  Offsets are used in the tile's shared-memory
// layout: with the 128-byte swizzle (swz = 0x70) the lane part is swizzled here, once per entry; the per-j offsets and the
// X-mask were swizzled by the host (the map is linear over XOR).
template <bool REAL, bool RL = false>
__device__ __forceinline__ double lean_entries(const char* tb, const DevFlat2* __restrict__ flats, int e0, int e1, int stride,
                                               uint32_t lane, uint64_t sbase, const uint64_t* __restrict__ fzout,
                                               const double* __restrict__ addtab, const double* s_beta, uint32_t swz) {
    double er = 0.0;
    if (e0 >= e1) return er;
    const uint4* ep = reinterpret_cast<const uint4*>(flats + e0);
    uint4 q0 = __ldg(ep), q1 = __ldg(ep + 1), q2 = __ldg(ep + 2);
    uint64_t zo = __ldg(fzout + (q0.w >> 16));
    for (int e = e0; e < e1; e += stride) {
        uint4 n0 = q0, n1 = q1, n2 = q2;
        uint64_t nzo = zo;
        if (e + stride < e1) {
            const uint4* np = reinterpret_cast<const uint4*>(flats + e + stride);
            n0 = __ldg(np);
            n1 = __ldg(np + 1);
            n2 = __ldg(np + 2);
        }
        LeanUnit u;
        lean_decode(q0, q1, q2, lane, u, RL ? 3u : 4u);
        const uint32_t sg = u.s0 + (uint32_t)__popcll(sbase & zo);
        const uint32_t vs = swz_off(u.v, swz);
        double part = 0.0;
        if (u.tab == 0xffffu) {
            // The sign of pair j is LINEAR in the bits of j (o_j is the XOR of three basis offsets, plus the chunk part that
            // all eight share): sum_j (-1)^sigma(j) w_j is a three-level butterfly with one +-1.0 factor per level -- seven
            // DFMAs and no per-pair sign arithmetic.  The common sign (bit 0 of jsign) joins the entry's sign below.
            double wj[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t off = vs ^ u.o[j];
                if (REAL) {
                    wj[j] = *reinterpret_cast<const double*>(tb + off) * *reinterpret_cast<const double*>(tb + (off ^ u.lx16));
                } else {
                    const double2 a = *reinterpret_cast<const double2*>(tb + off);
                    const double2 b = *reinterpret_cast<const double2*>(tb + (off ^ u.lx16));
                    wj[j] = fma(b.x, a.x, b.y * a.y);  // Re(conj(b) a)
                }
            }
            const double g0 = flipsign(1.0, u.jsign ^ (u.jsign >> 1)), g1 = flipsign(1.0, u.jsign ^ (u.jsign >> 2)),
                         g2 = flipsign(1.0, u.jsign ^ (u.jsign >> 4));
            const double t0 = fma(g0, wj[1], wj[0]), t1 = fma(g0, wj[3], wj[2]), t2 = fma(g0, wj[5], wj[4]), t3 = fma(g0, wj[7], wj[6]);
            part = flipsign(fma(g2, fma(g1, t3, t2), fma(g1, t1, t0)), u.jsign);
        } else {
            const double gl = s_beta[u.bidx] + __ldg(addtab + u.tab + lane);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t off = vs ^ u.o[j];
                double w;
                if (REAL) {
                    w = *reinterpret_cast<const double*>(tb + off) * *reinterpret_cast<const double*>(tb + (off ^ u.lx16));
                } else {
                    const double2 a = *reinterpret_cast<const double2*>(tb + off);
                    const double2 b = *reinterpret_cast<const double2*>(tb + (off ^ u.lx16));
                    w = fma(b.x, a.x, b.y * a.y);
                }
                part = fma(flipsign(w, u.jsign >> j), gl + __ldg(addtab + u.hi0 + j), part);
            }
        }
        er = fma(u.fr, flipsign(part, sg), er);
        if (e + stride < e1) nzo = __ldg(fzout + (n0.w >> 16));
        q0 = n0; q1 = n1; q2 = n2;
        zo = nzo;
    }
    return er;
}
// RL: real layout of the state buffer (8-byte tile elements; `flats` are then the real-layout entries, REAL is implied)
template <bool REAL, int THREADS, bool RL = false>
__global__ void __launch_bounds__(THREADS, RL ? 4 : 3) k_expect_lean(const __grid_constant__ CUtensorMap tmap, Shards psi, TileGeom g,
                                                        const DevFlat2* __restrict__ flats, int n_flats,
                                                        const uint64_t* __restrict__ fzout, const double* __restrict__ addtab,
                                                        const DevAddPat* __restrict__ addpat, int n_addpat,
                                                        const DevAddOut* __restrict__ addout, double2* __restrict__ partial,
                                                        int* __restrict__ err) {
    extern __shared__ __align__(1024) double2 tile[];
    __shared__ double red[64];
    __shared__ uint64_t s_boff[16];
    __shared__ __align__(8) uint64_t s_mbar;
    const uint32_t ts = 1u << g.tbits;
    double* s_beta = RL ? reinterpret_cast<double*>(tile) + ts : (double*)(tile + ts);
    const bool tma = g.tma != 0;
    const bool bulk = g.bulk != 0 || tma;
    uint32_t mphase = 0;
    if (bulk && threadIdx.x == 0) mbar_init(&s_mbar, 1);
    const TileAddr ta = tile_addr_init(g, s_boff);
    const BaseLane bl = base_lane_init(g);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // blockIdx.y splits the entry list (states with few tiles)
    const int fper = (n_flats + gridDim.y - 1) / gridDim.y;
    const int f0 = min(n_flats, (int)blockIdx.y * fper), f1 = min(n_flats, f0 + fper);
    const char* tb = (const char*)tile;
    double er = 0.0;
    for (uint64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        const uint64_t base = tile_base_warp(g, bl, t);
        const uint64_t sbase = base | g.sign_base;
        __syncthreads();  // previous tile fully consumed
        if (tma) {
            if (threadIdx.x == 0) tma_load_tile(tile, &tmap, g, base, &s_mbar);
        } else if (bulk) tile_load_bulk(tile, psi, g, base, &s_mbar);
        else tile_load_async_fast(tile, psi, g, ta, s_boff, base);
        lean_betas(s_beta, addpat, n_addpat, addout, sbase);
        if (bulk) {
            if (!mbar_wait(&s_mbar, mphase) && err) *err = 2;
            mphase ^= 1u;
        } else {
            cp_async_wait_all();
        }
        __syncthreads();
        er += lean_entries<REAL, RL>(tb, flats, f0 + (int)warp, f1, (int)nw, lane, sbase, fzout, addtab, s_beta, g.swz);
    }
    double2 sres = block_sum2(er, 0.0, red);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = sres;
}

// ------------------------------------------------------------------------------------------
// k_expect_diag2_rl: <psi|D|psi> on the REAL LAYOUT for a diagonal operator that is at most quadratic in the Z letters,
//   D(l) = c0 + sum_p a_p s_p + sum_{p<q} b_pq s_p s_q,   s_p = (-1)^(bit p of l)
// -- the X-mask-0 group of a molecular Hamiltonian (number operators and Coulomb / exchange pairs: Z_p, Z_p Z_q).  The
// general expectation kernel needs the interleaved complex form (one in-place expansion of the state per evaluation) and
// a Walsh-Hadamard pass over hundreds of strings; here a chunk of 2^tb contiguous amplitudes (index bits t below tb, o above)
// splits D into  K(o) + sum_{p<tb} s_p v_p(o) + T(t):  K and the tb coefficients v_p = a_p + sum_{q>=tb} b_pq s_q are
// computed once per chunk, the linear part is tabulated as A[t & 31] + B[t >> 5], and T (2^tb doubles, the same for every
// chunk) comes from the host.  Per amplitude: psi^2 times three table values; the state streams through once, coalesced.
// Fixed grid, per-CTA partials, fixed-order block reduction: bit-reproducible like the other expectation kernels.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_expect_diag2_rl(const double* __restrict__ psi, uint64_t n_amp, int n, int tb, uint64_t sign_base,
                                                         double c0, const double* __restrict__ a, const double* __restrict__ b,
                                                         const double* __restrict__ t_tab, int rev, double2* __restrict__ partial) {
    __shared__ double sv[16], sA[32], sB[256], sK, red[64];
    const uint32_t tid = threadIdx.x, chunk = 1u << tb;
    const uint64_t n_chunks = n_amp >> tb;
    const int lb = tb < 5 ? tb : 5;
    double acc = 0.0;
    for (uint64_t k = blockIdx.x; k < n_chunks; k += gridDim.x) {
        const uint64_t ch = rev ? n_chunks - 1 - k : k;   // alternating walk (see launch_plan): start where the previous pass ended
        const uint64_t o = (ch << tb) | sign_base;   // the chunk's amplitudes are o | t, t < 2^tb
        __syncthreads();                             // the tables of the previous chunk are consumed
        if ((int)tid < tb) {
            double v = a[tid];
            for (int q = tb; q < n; ++q) {
                const double bq = b[(size_t)tid * n + q];
                v += ((o >> q) & 1ull) ? -bq : bq;
            }
            sv[tid] = v;
        }
        if (tid >= 32u && tid < 64u) {   // K(o): lane L sums the rows p = tb + L, tb + L + 32, ...; fixed-shape warp reduction
            double k = 0.0;
            for (int pp = tb + (int)(tid - 32u); pp < n; pp += 32) {
                double inner = a[pp];
                for (int q = pp + 1; q < n; ++q) {
                    const double bq = b[(size_t)pp * n + q];
                    inner += ((o >> q) & 1ull) ? -bq : bq;
                }
                k += ((o >> pp) & 1ull) ? -inner : inner;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) k += __shfl_xor_sync(0xffffffffu, k, d);
            if (tid == 32u) sK = k + c0;
        }
        __syncthreads();
        if (tid < 32u) {
            double sgm = 0.0;
            for (int pp = 0; pp < lb; ++pp) sgm += ((tid >> pp) & 1u) ? -sv[pp] : sv[pp];
            sA[tid] = sgm;
        }
        for (uint32_t j = tid; j < (chunk >> lb) && j < 256u; j += 256u) {
            double sgm = 0.0;
            for (int pp = lb; pp < tb; ++pp) sgm += ((j >> (pp - lb)) & 1u) ? -sv[pp] : sv[pp];
            sB[j] = sgm;
        }
        __syncthreads();
        const double ka = sK + sA[tid & 31u];   // t & 31 == tid & 31 for every t this thread visits (stride 256)
        const double* __restrict__ src = psi + (ch << tb);
        for (uint32_t t = tid; t < chunk; t += 256u) {
            const double x = src[t];
            acc = fma(x * x, ka + sB[t >> lb] + __ldg(t_tab + t), acc);
        }
    }
    const double2 sres = block_sum2(acc, 0.0, red);
    if (tid == 0) partial[blockIdx.x] = sres;
}

// One entry's pair sum on one real-layout tile (see lean_entries): eight pairs per lane, signs by the three-level butterfly.
__device__ __forceinline__ double lean_part_rl(const char* tb, const LeanUnit& u, uint32_t vs, uint32_t lane, const double* __restrict__ addtab,
                                               const double* s_beta) {
    double part;
    if (u.tab == 0xffffu) {
        double wj[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t off = vs ^ u.o[j];
            wj[j] = *reinterpret_cast<const double*>(tb + off) * *reinterpret_cast<const double*>(tb + (off ^ u.lx16));
        }
        const double g0 = flipsign(1.0, u.jsign ^ (u.jsign >> 1)), g1 = flipsign(1.0, u.jsign ^ (u.jsign >> 2)),
                     g2 = flipsign(1.0, u.jsign ^ (u.jsign >> 4));
        const double t0 = fma(g0, wj[1], wj[0]), t1 = fma(g0, wj[3], wj[2]), t2 = fma(g0, wj[5], wj[4]), t3 = fma(g0, wj[7], wj[6]);
        part = flipsign(fma(g2, fma(g1, t3, t2), fma(g1, t1, t0)), u.jsign);
    } else {
        part = 0.0;
        const double gl = s_beta[u.bidx] + __ldg(addtab + u.tab + lane);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t off = vs ^ u.o[j];
            const double w = *reinterpret_cast<const double*>(tb + off) * *reinterpret_cast<const double*>(tb + (off ^ u.lx16));
            part = fma(flipsign(w, u.jsign >> j), gl + __ldg(addtab + u.hi0 + j), part);
        }
    }
    return part;
}
// lean_entries for TWO real-layout tiles resident at once (tb1 == nullptr: one): an entry is fetched and decoded once
// (about two thirds of the instructions of an entry are decode: descriptor unpack, index deposit, parity) and evaluated on
// both tiles -- only the outside-tile sign and the per-tile constants differ.
// lanetab != nullptr: the entries are those of the real-layout twin (PSPass::rlp) -- the per-lane part of an entry (byte offset
// of the lane's a-side element for j = 0, pattern and swizzle included, with the Z parity in bit 0) is read from the twin's
// lane table, 32 words per entry, instead of being deposited from the lane number.
__device__ __forceinline__ double lean_entries_pair(const char* tb0, const char* tb1, const DevFlat2* __restrict__ flats, int e0, int e1,
                                                    int stride, uint32_t lane, uint64_t sbase0, uint64_t sbase1,
                                                    const uint64_t* __restrict__ fzout, const double* __restrict__ addtab,
                                                    const double* s_beta0, const double* s_beta1, uint32_t swz,
                                                    const uint16_t* __restrict__ lanetab = nullptr) {
    double er = 0.0;
    if (e0 >= e1) return er;
    const uint4* ep = reinterpret_cast<const uint4*>(flats + e0);
    uint4 q0 = __ldg(ep), q1 = __ldg(ep + 1), q2 = __ldg(ep + 2);
    uint64_t zo = __ldg(fzout + (q0.w >> 16));
    uint32_t lw = lanetab ? (uint32_t)__ldg(lanetab + (size_t)e0 * 32u + lane) : 0u;
    for (int e = e0; e < e1; e += stride) {
        uint4 n0 = q0, n1 = q1, n2 = q2;
        uint64_t nzo = zo;
        uint32_t nlw = lw;
        if (e + stride < e1) {
            const uint4* np = reinterpret_cast<const uint4*>(flats + e + stride);
            n0 = __ldg(np);
            n1 = __ldg(np + 1);
            n2 = __ldg(np + 2);
            if (lanetab) nlw = (uint32_t)__ldg(lanetab + (size_t)(e + stride) * 32u + lane);
        }
        LeanUnit u;
        uint32_t vs;
        if (lanetab) {
            lean_decode_uniform(q0, q1, q2, u);
            vs = lw & 0xfff8u;
            u.s0 = lw & 1u;
        } else {
            lean_decode(q0, q1, q2, lane, u, 3u);
            vs = swz_off(u.v, swz);
        }
        double part = flipsign(lean_part_rl(tb0, u, vs, lane, addtab, s_beta0), u.s0 + (uint32_t)__popcll(sbase0 & zo));
        if (tb1) part += flipsign(lean_part_rl(tb1, u, vs, lane, addtab, s_beta1), u.s0 + (uint32_t)__popcll(sbase1 & zo));
        er = fma(u.fr, part, er);
        if (e + stride < e1) nzo = __ldg(fzout + (n0.w >> 16));
        q0 = n0; q1 = n1; q2 = n2;
        zo = nzo;
        lw = nlw;
    }
    return er;
}
// Real-layout expectation pass, pair mode: both tile slots of the CTA are loaded and every entry is evaluated on both
// (lean_entries_pair); other CTAs of the SM cover the load.  Same launch geometry and shared-memory layout as k_expect_rl2.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 3) k_expect_rlp(const __grid_constant__ CUtensorMap tmap, TileGeom g,
                                                        const DevFlat2* __restrict__ flats, int n_flats,
                                                        const uint64_t* __restrict__ fzout, const double* __restrict__ addtab,
                                                        const DevAddPat* __restrict__ addpat, int n_addpat,
                                                        const DevAddOut* __restrict__ addout, const uint16_t* __restrict__ lanetab,
                                                        double2* __restrict__ partial, int* __restrict__ err) {
    extern __shared__ __align__(1024) double2 tile[];
    __shared__ double red[64];
    __shared__ __align__(8) uint64_t s_mbar;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t slot_bytes = 8u << g.tbits;
    char* tb = reinterpret_cast<char*>(tile);
    double* s_beta = reinterpret_cast<double*>(tb + 2u * slot_bytes);   // [2][n_addpat]
    if (threadIdx.x == 0) mbar_init(&s_mbar, 2);
    const BaseLane bl = base_lane_init(g);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int fper = (n_flats + gridDim.y - 1) / gridDim.y;   // blockIdx.y splits the entry list (states with few tiles)
    const int f0 = min(n_flats, (int)blockIdx.y * fper), f1 = min(n_flats, f0 + fper);
    uint32_t mphase = 0;
    double er = 0.0;
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the state is complete
    for (uint64_t t = blockIdx.x; t < g.n_tiles; t += 2ull * gridDim.x) {
        const uint64_t t1 = t + gridDim.x;
        const bool two = t1 < g.n_tiles;
        const uint64_t base0 = tile_base_warp(g, bl, t);
        const uint64_t base1 = tile_base_warp(g, bl, two ? t1 : t);
        __syncthreads();  // the previous pair is fully consumed
        if (threadIdx.x == 0) {
            tma_load_tile(tile, &tmap, g, base0, &s_mbar);
            if (two) tma_load_tile(reinterpret_cast<double2*>(tb + slot_bytes), &tmap, g, base1, &s_mbar);
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&s_mbar)) : "memory");
        }
        lean_betas(s_beta, addpat, n_addpat, addout, base0 | g.sign_base);
        if (two) lean_betas(s_beta + n_addpat, addpat, n_addpat, addout, base1 | g.sign_base);
        if (!mbar_wait(&s_mbar, mphase) && err) *err = 2;
        mphase ^= 1u;
        __syncthreads();
        er += lean_entries_pair(tb, two ? tb + slot_bytes : nullptr, flats, f0 + (int)warp, f1, (int)nw, lane, base0 | g.sign_base,
                                    base1 | g.sign_base, fzout, addtab, s_beta, s_beta + n_addpat, g.swz, lanetab);
    }
    double2 sres = block_sum2(er, 0.0, red);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = sres;
}

// Real-layout expectation pass with TWO tile slots per CTA: the pass only reads the state, so the tensor-map load of the
// CTA's next tile (and its per-tile constants) is issued before the entries of the current tile are evaluated -- no warp
// ever waits for HBM after the first tile.  Three CTAs per SM (2 x 32 KiB each).  Programmatic dependent launch: the
// prologue runs under the tail of the previous pass; the state is first read after griddepcontrol.wait.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 3) k_expect_rl2(const __grid_constant__ CUtensorMap tmap, TileGeom g,
                                                        const DevFlat2* __restrict__ flats, int n_flats,
                                                        const uint64_t* __restrict__ fzout, const double* __restrict__ addtab,
                                                        const DevAddPat* __restrict__ addpat, int n_addpat,
                                                        const DevAddOut* __restrict__ addout, const uint16_t* __restrict__ lanetab,
                                                        double2* __restrict__ partial, int* __restrict__ err) {
    extern __shared__ __align__(1024) double2 tile[];
    __shared__ double red[64];
    __shared__ __align__(8) uint64_t s_mbar[2];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t slot_bytes = 8u << g.tbits;
    char* tb = reinterpret_cast<char*>(tile);
    double* s_beta = reinterpret_cast<double*>(tb + 2u * slot_bytes);   // [2][n_addpat]
    if (threadIdx.x == 0) {
        mbar_init(&s_mbar[0], 1);
        mbar_init(&s_mbar[1], 1);
    }
    const BaseLane bl = base_lane_init(g);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int fper = (n_flats + gridDim.y - 1) / gridDim.y;   // blockIdx.y splits the entry list (states with few tiles)
    const int f0 = min(n_flats, (int)blockIdx.y * fper), f1 = min(n_flats, f0 + fper);
    uint32_t ph0 = 0, ph1 = 0;
    double er = 0.0;
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the state is complete
    uint64_t t = blockIdx.x;
    uint64_t base = 0;
    if (t < g.n_tiles) {
        base = tile_base_warp(g, bl, t);
        if (threadIdx.x == 0) tma_load_tile(tile, &tmap, g, base, &s_mbar[0]);
        lean_betas(s_beta, addpat, n_addpat, addout, base | g.sign_base);
    }
    __syncthreads();
    for (uint32_t k = 0; t < g.n_tiles; t += gridDim.x, ++k) {
        const uint32_t sl = k & 1u;
        const uint64_t sbase = base | g.sign_base;
        const uint64_t tn = t + gridDim.x;
        if (tn < g.n_tiles) {  // the other slot was consumed before the barrier that ended the previous iteration
            base = tile_base_warp(g, bl, tn);
            if (threadIdx.x == 0)
                tma_load_tile(reinterpret_cast<double2*>(tb + (sl ^ 1u) * slot_bytes), &tmap, g, base, &s_mbar[sl ^ 1u]);
            lean_betas(s_beta + (sl ^ 1u) * n_addpat, addpat, n_addpat, addout, base | g.sign_base);
        }
        const bool ok = mbar_wait(&s_mbar[sl], sl ? ph1 : ph0);
        if (!ok && err) *err = 2;
        if (sl) ph1 ^= 1u; else ph0 ^= 1u;
        er += lean_entries_pair(tb + sl * slot_bytes, nullptr, flats, f0 + (int)warp, f1, (int)nw, lane, sbase, sbase, fzout, addtab,
                                s_beta + sl * n_addpat, s_beta + sl * n_addpat, g.swz, lanetab);
        __syncthreads();  // slot sl is free again; the constants of the next tile are in place
    }
    double2 sres = block_sum2(er, 0.0, red);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = sres;
}

// The same evaluation on the tile ring (one persistent CTA per SM, loads three tiles ahead of the arithmetic).
template <bool REAL>
__global__ void __launch_bounds__(512, 1) k_expect_pipe(const __grid_constant__ CUtensorMap tmap, Shards psi, TileGeom g,
                                                        const DevFlat2* __restrict__ flats, int n_flats,
                                                        const uint64_t* __restrict__ fzout, const double* __restrict__ addtab,
                                                        const DevAddPat* __restrict__ addpat, int n_addpat,
                                                        const DevAddOut* __restrict__ addout, double2* __restrict__ partial,
                                                        int* __restrict__ err) {
    extern __shared__ __align__(1024) double2 smem_tiles[];
    __shared__ double red[64];
    __shared__ __align__(8) uint64_t s_full[3];
    const uint32_t ts = 1u << g.tbits;
    double* s_beta = (double*)(smem_tiles + (size_t)3 * ts);  // [2][n_addpat], double-buffered per tile
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0)
        for (int i = 0; i < 3; ++i) mbar_init(&s_full[i], 1);
    const BaseLane bl = base_lane_init(g);
    const int fper = (n_flats + gridDim.y - 1) / gridDim.y;
    const int f0 = min(n_flats, (int)blockIdx.y * fper), f1 = min(n_flats, f0 + fper);
    const uint64_t n_my = g.n_tiles > blockIdx.x ? (g.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    __syncthreads();
    if (warp == 0)
        for (uint64_t k = 0; k < 3 && k < n_my; ++k)
            ring_issue_load(smem_tiles + (size_t)k * ts, psi, g, tile_base_warp(g, bl, blockIdx.x + k * gridDim.x), &s_full[k], &tmap);
    double er = 0.0;
    for (uint64_t k = 0; k < n_my; ++k) {
        const uint32_t sl = (uint32_t)(k % 3);
        const char* tb = (const char*)(smem_tiles + (size_t)sl * ts);
        const uint64_t sbase = tile_base_warp(g, bl, blockIdx.x + k * gridDim.x) | g.sign_base;
        double* beta = s_beta + (k & 1u) * n_addpat;
        lean_betas(beta, addpat, n_addpat, addout, sbase);
        __syncthreads();  // constants visible; every warp has left tile k-1 (its slot may be refilled below)
        if (warp == 0 && k >= 1 && k + 2 < n_my)
            ring_issue_load(smem_tiles + (size_t)((k + 2) % 3) * ts, psi, g, tile_base_warp(g, bl, blockIdx.x + (k + 2) * gridDim.x),
                            &s_full[(k + 2) % 3], &tmap);
        if (!mbar_wait(&s_full[sl], (uint32_t)((k / 3) & 1u)) && err) *err = 2;
        er += lean_entries<REAL>(tb, flats, f0 + (int)warp, f1, (int)nw, lane, sbase, fzout, addtab, beta, g.swz);
    }
    double2 sres = block_sum2(er, 0.0, red);
    if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = sres;
}

// sigma (+)= O psi for a lean pass.  The CTA holds the source tile and an accumulator tile (REAL: real parts only,
// 32 KiB -> two CTAs per SM) and walks the groups in order: one thread per (pattern, free index) pair,
//     acc[l ^ x] += G(l) psi[l],   acc[l] += G(l) psi[l ^ x]        (even-ny strings: the same weight both ways)
// with a barrier between two groups (inside a group every accumulator element is touched by one thread).
template <bool REAL>
__global__ void __launch_bounds__(512, REAL ? 2 : 1) k_apply_lean(const __grid_constant__ CUtensorMap tmap, Shards src, Shards dst, TileGeom g,
                                                                  const DevFlat2* __restrict__ flats,
                                                                  const uint32_t* __restrict__ goff, int n_groups,
                                                                  const uint64_t* __restrict__ fzout, const double* __restrict__ addtab,
                                                                  const DevAddPat* __restrict__ addpat, int n_addpat,
                                                                  const DevAddOut* __restrict__ addout, int accumulate,
                                                                  int* __restrict__ err) {
    extern __shared__ __align__(1024) double2 tile[];
    __shared__ uint64_t s_boff[16];
    __shared__ __align__(8) uint64_t s_mbar;
    const uint32_t ts = 1u << g.tbits;
    double* accr = (double*)(tile + ts);              // REAL: ts doubles
    double2* accc = (double2*)(tile + ts);            // else: ts double2
    double* s_beta = REAL ? (accr + ts) : (double*)(accc + ts);
    const bool tma = g.tma != 0;
    const bool bulk = g.bulk != 0 || tma;
    uint32_t mphase = 0;
    if (bulk && threadIdx.x == 0) mbar_init(&s_mbar, 1);
    const TileAddr ta = tile_addr_init(g, s_boff);
    const BaseLane bl = base_lane_init(g);
    const uint32_t bd = blockDim.x, lane = threadIdx.x & 31u;
    const char* tb = (const char*)tile;
    for (uint64_t t = blockIdx.x; t < g.n_tiles; t += gridDim.x) {
        const uint64_t base = tile_base_warp(g, bl, t);
        const uint64_t sbase = base | g.sign_base;
        __syncthreads();  // previous tile written out
        if (tma) {
            if (threadIdx.x == 0) tma_load_tile(tile, &tmap, g, base, &s_mbar);
        } else if (bulk) tile_load_bulk(tile, src, g, base, &s_mbar);
        else tile_load_async_fast(tile, src, g, ta, s_boff, base);
        for (uint32_t k = threadIdx.x; k < ts; k += bd) {
            if (REAL) accr[k] = 0.0;
            else accc[k] = make_double2(0.0, 0.0);
        }
        lean_betas(s_beta, addpat, n_addpat, addout, sbase);
        if (bulk) {
            if (!mbar_wait(&s_mbar, mphase) && err) *err = 2;
            mphase ^= 1u;
        } else {
            cp_async_wait_all();
        }
        __syncthreads();
        for (int gi = 0; gi < n_groups; ++gi) {
            const uint32_t e0 = __ldg(goff + gi), e1 = __ldg(goff + gi + 1);
            const uint32_t items = (e1 - e0) << 8;
            for (uint32_t it = threadIdx.x; it < items; it += bd) {
                const uint32_t e = e0 + (it >> 8), j = (it >> 5) & 7u;
                const uint4* ep = reinterpret_cast<const uint4*>(flats + e);
                const uint4 q0 = __ldg(ep), q1 = __ldg(ep + 1), q2 = __ldg(ep + 2);
                LeanUnit u;
                lean_decode(q0, q1, q2, lane, u);
                const uint32_t oj = (j & 1u) ? ((j & 2u) ? ((j & 4u) ? u.o[7] : u.o[3]) : ((j & 4u) ? u.o[5] : u.o[1]))
                                             : ((j & 2u) ? ((j & 4u) ? u.o[6] : u.o[2]) : ((j & 4u) ? u.o[4] : u.o[0]));
                const uint32_t off = swz_off(u.v, g.swz) ^ oj, offb = off ^ u.lx16;  // oj, lx16: swizzled by the host
                const uint32_t sg = u.s0 + (u.jsign >> j) + (uint32_t)__popcll(sbase & __ldg(fzout + (q0.w >> 16)));
                double gw = 0.5 * u.fr;  // the expectation weight carries the factor 2 of the pair
                if (u.tab != 0xffffu) gw *= s_beta[u.bidx] + __ldg(addtab + u.tab + lane) + __ldg(addtab + u.hi0 + j);
                gw = flipsign(gw, sg);
                if (REAL) {
                    const double a = *reinterpret_cast<const double*>(tb + off), b = *reinterpret_cast<const double*>(tb + offb);
                    accr[offb >> 4] = fma(gw, a, accr[offb >> 4]);
                    accr[off >> 4] = fma(gw, b, accr[off >> 4]);
                } else {
                    const double2 a = *reinterpret_cast<const double2*>(tb + off), b = *reinterpret_cast<const double2*>(tb + offb);
                    double2 ob = accc[offb >> 4], oa = accc[off >> 4];
                    ob.x = fma(gw, a.x, ob.x); ob.y = fma(gw, a.y, ob.y);
                    oa.x = fma(gw, b.x, oa.x); oa.y = fma(gw, b.y, oa.y);
                    accc[offb >> 4] = ob;
                    accc[off >> 4] = oa;
                }
            }
            __syncthreads();
        }
        for (uint32_t k = threadIdx.x; k < ts; k += bd) {
            double2* dp = amp_addr(g, dst, base, k);
            const uint32_t ks = swz_idx(k, g.swz);  // the accumulator uses the source tile's layout
            double2 o = REAL ? make_double2(accr[ks], 0.0) : accc[ks];
            if (accumulate) {
                const double2 d = *dp;
                o.x += d.x;
                o.y += d.y;
            }
            *dp = o;
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct TilePlan {
    uint64_t tile_mask = 0, comp_mask = 0, n_tiles = 1;  // over the nl shard-local index bits
    int tbits = 0, lbits = 0;  // tbits counts the virtual bit of a peer pass
    int nl = 0;                // shard-local index bits
    bool vbit = false;         // peer pass: tile = same local tile of ranks r and r ^ gpat
    uint64_t gpat = 0;         // X pattern on the global bits (in rank space)
    std::vector<int> bits;     // local tile bits, ascending
    std::vector<uint64_t> scat;
};

static inline int popc64(uint64_t v) { return __builtin_popcountll(v); }

// The three lowerings of a full-width mask into a plan (see "sharding" in DESIGN.md):
//   lx: X bits inside the tile; the virtual bit is set when the global part equals the pass's pattern
//   lz: Z bits inside the tile; the virtual bit carries parity(z_global & gpat)  (hi rank = lo rank ^ gpat)
//   zout: Z bits outside the tile, global bits included (their value comes from TileGeom::sign_base)
static uint32_t pext_local(uint64_t v, const TilePlan& tp) {
    uint32_t out = 0;
    for (size_t j = 0; j < tp.bits.size(); ++j)
        if ((v >> tp.bits[j]) & 1ull) out |= 1u << j;
    return out;
}
static uint32_t plan_lx(uint64_t x, const TilePlan& tp) {
    uint32_t o = pext_local(x, tp);
    if (tp.vbit && (x >> tp.nl) == tp.gpat) o |= 1u << (tp.tbits - 1);
    return o;
}
static uint32_t plan_lz(uint64_t z, const TilePlan& tp) {
    uint32_t o = pext_local(z, tp);
    if (tp.vbit && (popc64((z >> tp.nl) & tp.gpat) & 1)) o |= 1u << (tp.tbits - 1);
    return o;
}
static uint64_t plan_zout(uint64_t z, const TilePlan& tp) {
    uint64_t zg = z >> tp.nl;
    // peer pass: the sign base is the LOWER rank of the pair, whose top pattern bit is always clear, so a Z letter
    // there never contributes (its effect on the higher rank is the virtual bit of lz).  Dropping it keeps the
    // outside-tile masks of the strings of one generator identical, which lets them collapse.
    if (tp.vbit) zg &= ~(1ull << (63 - __builtin_clzll(tp.gpat)));
    return (z & tp.comp_mask) | (zg << tp.nl);
}

// finalize a plan from a set of required LOCAL bits: add low bits first, then fill up to tbits_max
// (one less when the pass needs the virtual shard bit)
static TilePlan make_plan(int nl, uint64_t need_mask, int tbits_max, int low_bits, uint64_t gpat = 0) {
    TilePlan tp;
    tp.nl = nl;
    tp.vbit = gpat != 0;
    tp.gpat = gpat;
    int tb = std::min(tbits_max - (tp.vbit ? 1 : 0), nl);
    uint64_t m = need_mask;
    int lb = std::min(low_bits, tb);
    for (int b = 0; b < lb; ++b) m |= 1ull << b;
    // fill with the lowest unused bits
    for (int b = 0; b < nl && popc64(m) < tb; ++b) m |= 1ull << b;
    tp.tile_mask = m;
    const int tl = popc64(m);
    for (int b = 0; b < nl; ++b)
        if ((m >> b) & 1ull) tp.bits.push_back(b);
    int l = 0;
    while (l < tl && tp.bits[l] == l) ++l;
    tp.lbits = l;
    uint64_t full = (nl >= 64) ? ~0ull : ((1ull << nl) - 1ull);
    tp.comp_mask = full & ~m;
    tp.n_tiles = 1ull << (nl - tl);
    tp.tbits = tl + (tp.vbit ? 1 : 0);
    int hi = tl - tp.lbits;
    tp.scat.resize(1ull << hi);
    for (uint64_t v = 0; v < (1ull << hi); ++v) {
        uint64_t o = 0;
        for (int j = 0; j < hi; ++j)
            if ((v >> j) & 1ull) o |= 1ull << tp.bits[tp.lbits + j];
        tp.scat[v] = o;
    }
    return tp;
}
// capacity test used by the pass builders: do these local bits (plus the fixed low bits) fit?
static bool plan_fits(uint64_t need_local, uint64_t lowmask, int tbits_max, int nl, bool vbit);
// The fixed low tile bits are a coalescing preference (512-byte segments), not a requirement: an operator with many
// X/Y letters on high qubits gets a tile with a smaller floor (shorter segments).  Largest floor <= lb_max for which
// `need_local` fits the tile, or -1 when it cannot fit at all (more X/Y letters than tile bits).
static int fit_low_bits(uint64_t need_local, int lb_max, int tbits_max, int nl, bool vbit) {
    for (int lb = lb_max; lb >= 0; --lb)
        if (plan_fits(need_local, (1ull << lb) - 1ull, tbits_max, nl, vbit)) return lb;
    return -1;
}
static bool plan_fits(uint64_t need_local, uint64_t lowmask, int tbits_max, int nl, bool vbit) {
    return popc64(need_local | lowmask) <= std::min(tbits_max - (vbit ? 1 : 0), nl);
}

struct KernelProf {
    double ms = 0.0;
    uint64_t launches = 0;
};

#define MAX_RANKS 64
struct PlanCache;                      // cached pass plan of the last rotation program (defined with the planner)
static void free_plan_cache(PlanCache* pc);
struct vqe_ctx {
    int n = 0, device = 0, sm_count = 148;
    uint64_t n_amp = 0;  // amplitudes held by THIS context (2^nl)
    // sharding: the top g index bits are the rank.  Single-GPU context: g = 0, nl = n, rank = 0.
    int nl = 0, g = 0, rank = 0, world = 1;
    double2* peer_buf[3][MAX_RANKS] = {};   // shard pointers of the other ranks (IPC mapping or same-process pointer)
    bool peer_ipc[3][MAX_RANKS] = {};       // mapping came from cudaIpcOpenMemHandle (close on destroy)
    vqe_ctx* peer_ctx[MAX_RANKS] = {};      // same-process peers (their buffers are resolved at launch time)
    // device-side flag barrier (one process per GPU): flags[p] = last epoch rank p has signalled to me
    uint64_t* flags = nullptr;              // own flag array, MAX_RANKS entries, cudaMalloc'ed (IPC-exportable)
    uint64_t* peer_flags[MAX_RANKS] = {};   // the other ranks' flag arrays
    bool peer_flags_ipc[MAX_RANKS] = {};
    uint64_t** d_peer_flags = nullptr;      // device copy of peer_flags
    bool flags_dirty = true;
    uint64_t epoch = 0;
    int* h_err = nullptr;                   // mapped pinned: set by a barrier that timed out
    int* d_err = nullptr;
    cudaEvent_t ev_bar = nullptr;           // in-process group barrier
    bool psi_real = false;                  // buffer 0 is known to be purely real (imaginary parts exactly 0.0)
    // REAL LAYOUT of buffer 0: while the state is known to be purely real (|HF> followed by +-1-phase rotations: every UCC
    // energy evaluation) only the real parts are kept, as n_amp contiguous doubles at the start of the buffer.  Every
    // pass then moves half the bytes and the tiles hold 8-byte elements.  Entry points that are not layout-aware expand
    // the buffer in place first (ensure_complex); real_layout implies psi_real.
    bool real_layout = false;
    bool real_layout_ok = false;            // VQE_REAL_LAYOUT (default on) and an unsharded context
    // QUBIT RELABELLING of a sharded state (buffer 0 only): logical index bit b of the caller's masks lives at physical index
    // bit perm[b] of the shards.  Identity after vqe_set_basis_state / vqe_set_state; the rotation entry point swaps a global
    // with a local bit (k_swap_global_local: half a shard over NVLink, once) instead of running every later rotation that
    // flips that qubit as a peer pass.  Entry points that need the caller's labelling undo the swaps first.
    uint8_t perm[64];
    std::vector<std::pair<int, int>> swap_history;  // (global slot, local slot) of every swap since the last reset
    uint64_t n_swaps = 0, swap_bytes = 0;   // statistics: swaps executed, bytes this rank read from its partners in swaps
    PlanCache* plan_cache = nullptr;        // see rotations_impl
    bool walk_desc = false;                 // the last pass over buffer 0 walked its tiles downwards (see launch_plan: alternating walk)
    size_t l2_persist_bytes = 0, l2_window_max = 0;   // persisting-L2 set-aside and largest access-policy window (0: not used)
    char* d_coltab = nullptr;               // item table of the plan whose identity is coltab_gen (k_col_stab: 16-bit items; k_col_tab: 32-bit)
    size_t coltab_cap = 0;                  // in bytes
    uint64_t coltab_gen = 0;
    int coltab_mode = 0;                    // 1: 32-bit words (k_col_tab), 2: 16-bit items (k_col_stab)
    double2* gstage[2] = {nullptr, nullptr};  // staging buffers of gather-form peer passes (two: chunk k + 1 is fetched under chunk k)
    size_t gstage_cap[2] = {0, 0};            // in amplitudes
    cudaStream_t stream = nullptr;
    double2* buf[4] = {nullptr, nullptr, nullptr, nullptr};  // psi, sigma, work + aux (local only: never a peer-pass operand)
    // staging
    char* h_stage = nullptr;
    char* d_stage = nullptr;
    size_t stage_cap = 0;
    double2* d_partial = nullptr;
    size_t partial_cap = 0;  // in double2
    double2* d_result = nullptr;
    double2* h_result = nullptr;  // pinned
    size_t result_cap = 0;        // in double2
    int tile_bits = 12, low_bits = 5, threads = 512, ctas_per_sm = 2;
    uint64_t launches = 0;
    bool profiling = false;
    KernelProf prof[6];  // 0 state prep, 1 expectation, 2 apply, 3 pool, 4 peer state prep, 5 peer expectation
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // step timer
    std::vector<cudaEvent_t> ev_pool;           // recycled events
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_pending;
    uint64_t h2d_bytes = 0, d2h_bytes = 0;
    uint64_t gather_bytes = 0;   // bytes read from partner shards by k_gather_need (gather-form peer passes)
};

// Non-intrusive per-kernel timing: when profiling is on, every kernel of a class is bracketed by two
// events from a pool (no synchronisation); vqe_profile_read resolves them after the fact.
struct ProfScope {
    vqe_ctx* c;
    int which;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ProfScope(vqe_ctx* c_, int w);
    ~ProfScope();
};
static cudaEvent_t take_event(vqe_ctx* c) {
    if (!c->ev_pool.empty()) {
        cudaEvent_t e = c->ev_pool.back();
        c->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
ProfScope::ProfScope(vqe_ctx* c_, int w) : c(c_), which(w) {
    if (c->profiling) {
        e0 = take_event(c);
        e1 = take_event(c);
        cudaEventRecord(e0, c->stream);
    }
}
ProfScope::~ProfScope() {
    c->prof[which].launches++;
    if (c->profiling) {
        cudaEventRecord(e1, c->stream);
        c->ev_pending.push_back({which, {e0, e1}});
    }
}
static void resolve_profile(vqe_ctx* c) {
    if (c->ev_pending.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto& p : c->ev_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.second.first, p.second.second) == cudaSuccess) c->prof[p.first].ms += ms;
        c->ev_pool.push_back(p.second.first);
        c->ev_pool.push_back(p.second.second);
    }
    c->ev_pending.clear();
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

static int ensure_stage(vqe_ctx* c, size_t bytes) {
    if (bytes <= c->stage_cap) return VQE_OK;
    size_t cap = std::max(bytes, c->stage_cap * 2 + 4096);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->d_stage) cudaFree(c->d_stage);
    c->h_stage = nullptr;
    c->d_stage = nullptr;
    c->stage_cap = 0;
    CK(cudaMallocHost((void**)&c->h_stage, cap));
    CK(cudaMalloc((void**)&c->d_stage, cap));
    c->stage_cap = cap;
    return VQE_OK;
}
static int ensure_partial(vqe_ctx* c, size_t n) {
    if (n <= c->partial_cap) return VQE_OK;
    if (c->d_partial) cudaFree(c->d_partial);
    c->d_partial = nullptr;
    c->partial_cap = 0;
    CK(cudaMalloc((void**)&c->d_partial, n * sizeof(double2)));
    c->partial_cap = n;
    return VQE_OK;
}
static int ensure_result(vqe_ctx* c, size_t n) {
    if (n <= c->result_cap) return VQE_OK;
    if (c->d_result) cudaFree(c->d_result);
    if (c->h_result) cudaFreeHost(c->h_result);
    c->d_result = nullptr;
    c->h_result = nullptr;
    c->result_cap = 0;
    CK(cudaMalloc((void**)&c->d_result, n * sizeof(double2)));
    CK(cudaMallocHost((void**)&c->h_result, n * sizeof(double2)));
    c->result_cap = n;
    return VQE_OK;
}
static int ensure_buf(vqe_ctx* c, int b) {
    if (b < 0 || b > 3) return fail(VQE_ERR_INVALID, "bad buffer id %d", b);
    if (c->buf[b]) return VQE_OK;
    cudaError_t e = cudaMalloc((void**)&c->buf[b], c->n_amp * sizeof(double2));
    if (e != cudaSuccess) {
        c->buf[b] = nullptr;
        return fail(VQE_ERR_NOMEM, "cudaMalloc of state buffer %d (%.1f GB) failed: %s", b,
                    c->n_amp * 16.0 / 1e9, cudaGetErrorString(e));
    }
    CK(cudaMemsetAsync(c->buf[b], 0, c->n_amp * sizeof(double2), c->stream));
    return VQE_OK;
}

static int grid_1d(const vqe_ctx* c, uint64_t n_amp, int threads);
// buffer b in interleaved complex form (only the state buffer is ever kept in the real layout)
static int ensure_complex(vqe_ctx* c, int b) {
    if (b != VQE_BUF_PSI || !c->real_layout) return VQE_OK;
    double* buf = reinterpret_cast<double*>(c->buf[VQE_BUF_PSI]);
    for (uint64_t hi = c->n_amp; hi >= 1; hi >>= 1) {
        const uint64_t lo = hi >> 1;  // levels [N/2, N), [N/4, N/2), ..., [1, 2), [0, 1)
        k_expand_level<<<grid_1d(c, hi - lo, 256), 256, 0, c->stream>>>(buf, lo, hi);
        c->launches++;
        if (hi == 1) break;
    }
    CK(cudaGetLastError());
    c->real_layout = false;
    return VQE_OK;
}

static size_t tile_smem(int tbits, int n_tiles_in_smem, bool term_cache) {
    size_t s = (size_t)n_tiles_in_smem * (16ull << tbits);
    if (term_cache)
        s += TERM_CAP * (sizeof(DevTerm) + sizeof(double2)) + GROUP_CAP * (sizeof(DevGroup) + sizeof(DevGCol)) +
             GCOL_ENT_CAP * sizeof(DevGColEntry) + FLAT_CAP * (sizeof(DevFlat) + 8);
    return s;
}

static bool g_attr_done[64] = {};
static int set_kernel_attrs(int device) {
    if (device >= 0 && device < 64 && g_attr_done[device]) return VQE_OK;
    const int maxs = 227 * 1024;
#define SET_SMEM(k)                                                                              \
    do {                                                                                         \
        cudaFuncAttributes fa_;                                                                  \
        CK(cudaFuncGetAttributes(&fa_, k));                                                      \
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,                  \
                                maxs - (int)fa_.sharedSizeBytes));                               \
    } while (0)
    SET_SMEM(k_tile_ops);
    SET_SMEM(k_tile_rot<false>);
    SET_SMEM(k_tile_rot<true>);
    SET_SMEM((k_tile_col<false, false>));
    SET_SMEM((k_tile_col<true, false>));
    SET_SMEM((k_tile_col<true, true>));
    SET_SMEM(k_col_tab);
    SET_SMEM((k_col_stab<256, 3>));
    SET_SMEM((k_col_stab<512, 2>));
    SET_SMEM((k_col_stab<512, 3>));
    SET_SMEM(k_expect_rl2<256>);
    SET_SMEM(k_expect_rlp<256>);
    SET_SMEM(k_expect_rl2<384>);
    SET_SMEM(k_col_pipe<false>);
    SET_SMEM(k_col_pipe<true>);
    SET_SMEM(k_expect_pipe<false>);
    SET_SMEM(k_expect_pipe<true>);
    SET_SMEM(k_tile_expect<false>);
    SET_SMEM(k_tile_expect<true>);
    SET_SMEM(k_tile_apply);
    SET_SMEM((k_expect_lean<false, 256>));
    SET_SMEM((k_expect_lean<true, 256>));
    SET_SMEM((k_expect_lean<false, 384>));
    SET_SMEM((k_expect_lean<true, 384>));
    SET_SMEM((k_expect_lean<false, 512>));
    SET_SMEM((k_expect_lean<true, 512>));
    SET_SMEM(k_apply_lean<false>);
    SET_SMEM(k_apply_lean<true>);
    SET_SMEM(k_tile_pool);
#undef SET_SMEM
    if (device >= 0 && device < 64) g_attr_done[device] = true;
    return VQE_OK;
}

static void free_ctx(vqe_ctx* c);

static int create_ctx(vqe_ctx** out, int n_qubits, int n_global, int rank, int device) {
    if (!out) return fail(VQE_ERR_INVALID, "out is null");
    *out = nullptr;
    if (n_qubits < 1 || n_qubits > 40) return fail(VQE_ERR_INVALID, "n_qubits=%d out of range [1,40]", n_qubits);
    if (n_global < 0 || n_global > 6 || n_global >= n_qubits)
        return fail(VQE_ERR_INVALID, "n_global=%d out of range [0,min(6,n_qubits-1)]", n_global);
    if (rank < 0 || rank >= (1 << n_global)) return fail(VQE_ERR_INVALID, "rank %d not in [0,2^%d)", rank, n_global);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(VQE_ERR_CUDA, "no CUDA device available (%s): this engine has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(VQE_ERR_INVALID, "device %d not in [0,%d)", device, ndev);
    CK(cudaSetDevice(device));
    vqe_ctx* c = new vqe_ctx();
    c->n = n_qubits;
    c->g = n_global;
    c->nl = n_qubits - n_global;
    c->rank = rank;
    c->world = 1 << n_global;
    c->device = device;
    c->n_amp = 1ull << c->nl;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    // L2 residency of a state that is not much larger than the L2 (24 qubits in the real layout: 134 MB against 126 MB): the
    // pass kernels are launched with an access-policy window over the state that marks a fraction of its lines PERSISTING,
    // sized to the persisting set-aside, and the rest streaming -- consecutive passes then find that fraction in L2 (reads
    // and write-backs) instead of every pass streaming the whole state through HBM.  MEASURED SLOWER on B200 (rotation passes
    // 67 -> 114 us whatever the hit ratio: the set-aside shrinks the L2 left to the streaming tiles): off, VQE_L2_PERSIST=1 enables it.
    if (env_int("VQE_L2_PERSIST", 0) != 0 && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
        const size_t want = (size_t)((double)prop.persistingL2CacheMaxSize * env_int("VQE_L2_PERSIST_PCT", 100) / 100.0);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            c->l2_persist_bytes = want;
            c->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
        } else {
            cudaGetLastError();
        }
    }
    for (int b = 0; b < 64; ++b) c->perm[b] = (uint8_t)b;
    c->tile_bits = env_int("VQE_TILE_BITS", 12);
    // low-bit floor of the tiles: tensor-map (TMA) tile loads make short contiguous runs cheap, so an unsharded context only
    // insists on 16 amplitudes (256 bytes; 128 in the real layout) and leaves 8 tile bits to the planner; the peer passes of
    // a sharded state copy segment by segment and keep 512-byte segments
    c->low_bits = env_int("VQE_LOW_BITS", n_global ? 5 : 4);
    // (a sharded state keeps it through the rotation passes only, which qubit relabelling makes local; see launch_plan)
    c->real_layout_ok = env_int("VQE_REAL_LAYOUT", 1) != 0 && (n_global == 0 || env_int("VQE_RELABEL", 1) != 0);
    c->threads = env_int("VQE_THREADS", 512);
    c->ctas_per_sm = env_int("VQE_CTAS_PER_SM", 2);
    if (c->tile_bits < 6 || c->tile_bits > 12) c->tile_bits = 12;
    if (c->threads < 64 || c->threads > 512 || (c->threads & (c->threads - 1))) c->threads = 512;
    if (c->low_bits < 0 || c->low_bits > c->tile_bits) c->low_bits = 5;
    int rc = VQE_OK;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_bar, cudaEventDisableTiming) != cudaSuccess) {
            rc = fail(VQE_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        if ((rc = set_kernel_attrs(device))) break;
        if ((rc = ensure_buf(c, VQE_BUF_PSI))) break;
        if ((rc = ensure_result(c, 64))) break;
        // device error flag (mapped pinned host memory): 1 = a cross-rank barrier timed out, 2 = a bulk tile copy never
        // signalled its mbarrier.  Checked by vqe_synchronize, the reductions and vqe_shard_status.
        if (cudaHostAlloc((void**)&c->h_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess) {
            rc = fail(VQE_ERR_CUDA, "error flag allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        *c->h_err = 0;
        if (cudaHostGetDevicePointer((void**)&c->d_err, c->h_err, 0) != cudaSuccess) {
            rc = fail(VQE_ERR_CUDA, "cudaHostGetDevicePointer failed");
            break;
        }
        if (c->world > 1) {
            if (cudaMalloc((void**)&c->flags, MAX_RANKS * sizeof(uint64_t)) != cudaSuccess ||
                cudaMemset(c->flags, 0, MAX_RANKS * sizeof(uint64_t)) != cudaSuccess ||
                cudaMalloc((void**)&c->d_peer_flags, MAX_RANKS * sizeof(uint64_t*)) != cudaSuccess) {
                rc = fail(VQE_ERR_CUDA, "barrier flag allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
                break;
            }
            c->peer_flags[c->rank] = c->flags;
        }
    } while (0);
    if (rc) {
        free_ctx(c);
        return rc;
    }
    *out = c;
    return VQE_OK;
}

extern "C" int vqe_create(vqe_ctx** out, int n_qubits, int device) { return create_ctx(out, n_qubits, 0, 0, device); }

static void free_ctx(vqe_ctx* c) {
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int b = 0; b < 3; ++b) {
        for (int r = 0; r < MAX_RANKS; ++r)
            if (c->peer_buf[b][r] && c->peer_ipc[b][r]) cudaIpcCloseMemHandle(c->peer_buf[b][r]);
        if (c->buf[b]) cudaFree(c->buf[b]);
    }
    if (c->buf[3]) cudaFree(c->buf[3]);
    for (int r = 0; r < MAX_RANKS; ++r)
        if (c->peer_flags[r] && c->peer_flags_ipc[r]) cudaIpcCloseMemHandle(c->peer_flags[r]);
    if (c->flags) cudaFree(c->flags);
    if (c->d_peer_flags) cudaFree(c->d_peer_flags);
    if (c->h_err) cudaFreeHost(c->h_err);
    if (c->plan_cache) free_plan_cache(c->plan_cache);
    if (c->l2_persist_bytes) cudaCtxResetPersistingL2Cache();   // the persisting lines of this state go back to normal
    if (c->d_coltab) cudaFree(c->d_coltab);
    for (int sb = 0; sb < 2; ++sb)
        if (c->gstage[sb]) cudaFree(c->gstage[sb]);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->d_stage) cudaFree(c->d_stage);
    if (c->d_partial) cudaFree(c->d_partial);
    if (c->d_result) cudaFree(c->d_result);
    if (c->h_result) cudaFreeHost(c->h_result);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev_bar) cudaEventDestroy(c->ev_bar);
    resolve_profile(c);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" void vqe_destroy(vqe_ctx* c) {
    if (!c) return;
    free_ctx(c);
}

// ------------------------------------------------------------------------------------------
// sharding: one context per rank, the top g index bits are the rank.
//   * one process per GPU: the shards (and a small flag array) are exported as CUDA IPC handles, the Python
//     layer exchanges them through torch.distributed and attaches them here; cross-rank ordering is a
//     device-side flag barrier enqueued on the stream (no host round trip, no NCCL on the data path);
//   * one process driving all ranks (tests on one GPU, or a single-process multi-GPU run): peers are attached
//     by pointer and the vqe_group_* entry points order the ranks' streams with CUDA events.
// ------------------------------------------------------------------------------------------
// Exchange of a GLOBAL with a LOCAL index bit between the shards of ranks lo (global bit 0) and hi (global bit 1): lo's
// amplitudes with local bit L = 1 trade places with hi's amplitudes with L = 0 (all other bits equal).  Pair k of the
// n_amp / 2 pairs: i0 = k with a zero inserted at bit L.  The two ranks of a pair each take half of the pairs, so every GPU
// reads and writes a quarter of a shard remotely: half a shard per NVLink direction.  Consecutive lanes take consecutive
// pairs (512 contiguous bytes per warp and access stream -- a lane stride of 64 bytes cut the NVLink rate to a third),
// four pairs a block-width apart per thread: eight independent loads in flight.  T = double2, or double while the state
// is kept in the real layout.
template <typename T>
__global__ void __launch_bounds__(256) k_swap_global_local(T* __restrict__ lo, T* __restrict__ hi, uint32_t L, uint64_t first,
                                                           uint64_t count) {
    const uint64_t lowmask = (1ull << L) - 1ull, lbit = 1ull << L;
    const uint64_t end = first + count;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4ull;
    for (uint64_t k0 = first + (uint64_t)blockIdx.x * blockDim.x * 4ull + threadIdx.x; k0 < end; k0 += stride) {
        T a[4], b[4];
        uint64_t ia[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t k = k0 + (uint64_t)u * blockDim.x;
            ia[u] = ((k & ~lowmask) << 1) | (k & lowmask);
            if (k < end) {
                a[u] = lo[ia[u] | lbit];
                b[u] = hi[ia[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t k = k0 + (uint64_t)u * blockDim.x;
            if (k < end) {
                lo[ia[u] | lbit] = b[u];
                hi[ia[u]] = a[u];
            }
        }
    }
}

__global__ void k_flag_barrier(uint64_t* const* __restrict__ peer_flags, volatile uint64_t* my_flags, int rank,
                               int world, unsigned long long epoch, int* err) {
    const int p = threadIdx.x;
    if (p >= world || p == rank) return;
    // everything this rank's earlier kernels wrote (own shard and peer shards) is ordered before the signal
    __threadfence_system();
    uint64_t* slot = peer_flags[p] + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(slot), "l"(epoch) : "memory");
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(my_flags + p) : "memory");
        if (v >= epoch) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 60ull * 1000000000ull) {  // 60 s: a peer died; never hang the GPU
            *err = 1;
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

extern "C" int vqe_create_shard(vqe_ctx** out, int n_qubits, int n_global, int rank, int device) {
    return create_ctx(out, n_qubits, n_global, rank, device);
}
extern "C" int vqe_shard_info(const vqe_ctx* c, int* n_global, int* rank, int* n_local) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    if (n_global) *n_global = c->g;
    if (rank) *rank = c->rank;
    if (n_local) *n_local = c->nl;
    return VQE_OK;
}
extern "C" int vqe_shard_export(vqe_ctx* c, int what, void* handle_out) {
    if (!c || !handle_out) return fail(VQE_ERR_INVALID, "null argument");
    if (c->world == 1) return fail(VQE_ERR_INVALID, "not a sharded context");
    CK(cudaSetDevice(c->device));
    void* ptr = nullptr;
    if (what == VQE_SHARD_FLAGS) ptr = c->flags;
    else {
        int rc = ensure_buf(c, what);
        if (rc) return rc;
        ptr = c->buf[what];
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == VQE_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_out, &h, sizeof h);
    return VQE_OK;
}
extern "C" int vqe_shard_attach_ipc(vqe_ctx* c, int peer_rank, int what, const void* handle) {
    if (!c || !handle) return fail(VQE_ERR_INVALID, "null argument");
    if (peer_rank < 0 || peer_rank >= c->world || peer_rank == c->rank)
        return fail(VQE_ERR_INVALID, "peer rank %d invalid for rank %d of %d", peer_rank, c->rank, c->world);
    CK(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    if (what == VQE_SHARD_FLAGS) {
        c->peer_flags[peer_rank] = (uint64_t*)ptr;
        c->peer_flags_ipc[peer_rank] = true;
        c->flags_dirty = true;
    } else if (what >= 0 && what < 3) {
        c->peer_buf[what][peer_rank] = (double2*)ptr;
        c->peer_ipc[what][peer_rank] = true;
    } else {
        cudaIpcCloseMemHandle(ptr);
        return fail(VQE_ERR_INVALID, "bad export id %d", what);
    }
    return VQE_OK;
}
extern "C" int vqe_shard_attach_local(vqe_ctx* c, vqe_ctx* peer) {
    if (!c || !peer) return fail(VQE_ERR_INVALID, "null argument");
    if (peer->n != c->n || peer->g != c->g || peer->rank == c->rank)
        return fail(VQE_ERR_INVALID, "peer context does not belong to the same sharded state");
    if (peer->device != c->device) {
        CK(cudaSetDevice(c->device));
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, c->device, peer->device));
        if (!can) return fail(VQE_ERR_CUDA, "device %d cannot access device %d", c->device, peer->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return fail(VQE_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    c->peer_ctx[peer->rank] = peer;
    c->peer_flags[peer->rank] = peer->flags;
    c->peer_flags_ipc[peer->rank] = false;
    c->flags_dirty = true;
    return VQE_OK;
}

// device-side barrier over all ranks, enqueued on the stream (one-process-per-GPU mode)
static int flag_barrier(vqe_ctx* c) {
    if (c->world == 1) return VQE_OK;
    for (int r = 0; r < c->world; ++r)
        if (!c->peer_flags[r]) return fail(VQE_ERR_INVALID, "rank %d: flag array of rank %d is not attached", c->rank, r);
    if (c->flags_dirty) {
        CK(cudaMemcpyAsync(c->d_peer_flags, c->peer_flags, MAX_RANKS * sizeof(uint64_t*), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        c->flags_dirty = false;
    }
    c->epoch++;
    k_flag_barrier<<<1, MAX_RANKS, 0, c->stream>>>(c->d_peer_flags, c->flags, c->rank, c->world,
                                                  (unsigned long long)c->epoch, c->d_err);
    c->launches++;
    CK(cudaGetLastError());
    return VQE_OK;
}
extern "C" int vqe_shard_barrier(vqe_ctx* c) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    CK(cudaSetDevice(c->device));
    return flag_barrier(c);
}
extern "C" int vqe_shard_status(vqe_ctx* c) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    if (c->h_err && *(volatile int*)c->h_err == 1) return fail(VQE_ERR_CUDA, "rank %d: a cross-rank barrier timed out (peer lost)", c->rank);
    if (c->h_err && *(volatile int*)c->h_err)
        return fail(VQE_ERR_CUDA, "rank %d: a bulk tile copy never completed (mbarrier wait timed out); results are invalid", c->rank);
    return VQE_OK;
}

// The set of ranks one host thread drives: {ctx} in one-process-per-GPU mode (barrier = device flags), or all
// ranks of the state in in-process mode (barrier = events across the ranks' streams).
struct RankSet {
    std::vector<vqe_ctx*> r;
};
static int check_rankset(const RankSet& rs) {
    if (rs.r.empty() || !rs.r[0]) return fail(VQE_ERR_INVALID, "ctx is null");
    const vqe_ctx* c0 = rs.r[0];
    if (rs.r.size() == 1) return VQE_OK;
    if ((int)rs.r.size() != c0->world) return fail(VQE_ERR_INVALID, "group has %zu contexts, the state has %d ranks", rs.r.size(), c0->world);
    for (size_t k = 0; k < rs.r.size(); ++k) {
        const vqe_ctx* c = rs.r[k];
        if (!c || c->n != c0->n || c->g != c0->g || c->rank != (int)k || c->tile_bits != c0->tile_bits ||
            c->low_bits != c0->low_bits || c->threads != c0->threads)
            return fail(VQE_ERR_INVALID, "group contexts must be ranks 0..%d of one sharded state, in order", c0->world - 1);
    }
    return VQE_OK;
}
static int rank_barrier(RankSet& rs) {
    if (rs.r.size() == 1) {
        CK(cudaSetDevice(rs.r[0]->device));
        return flag_barrier(rs.r[0]);
    }
    for (vqe_ctx* c : rs.r) {
        CK(cudaSetDevice(c->device));
        CK(cudaEventRecord(c->ev_bar, c->stream));
    }
    for (vqe_ctx* c : rs.r) {
        CK(cudaSetDevice(c->device));
        for (vqe_ctx* p : rs.r)
            if (p != c) CK(cudaStreamWaitEvent(c->stream, p->ev_bar, 0));
    }
    return VQE_OK;
}

// per-rank launch geometry of a plan
static double2* shard_ptr(const vqe_ctx* c, int buf, int rank) {
    if (rank == c->rank) return c->buf[buf];
    if (buf > 2) return nullptr;  // the aux buffer is not shared between ranks
    if (c->peer_ctx[rank]) return c->peer_ctx[rank]->buf[buf];
    return c->peer_buf[buf][rank];
}
static int make_geom(const vqe_ctx* c, const TilePlan& tp, const uint64_t* d_scat, int buf, TileGeom& g, Shards& sh) {
    g.comp_mask = tp.comp_mask;
    g.scat = d_scat;
    g.tbits = tp.tbits;
    g.lbits = tp.lbits;
    g.vbit = tp.vbit ? 1u : 0u;
    g.bulk = (tp.lbits >= 3 && env_int("VQE_BULK", 1)) ? 1u : 0u;  // segments of at least 128 bytes
    g.tma = 0;
    g.swz = 0;
    g.rl = 0;
    g.rev = 0;
    memset(&g.tg, 0, sizeof g.tg);
    if (!tp.vbit) {
        g.n_tiles = tp.n_tiles;
        g.tile_first = 0;
        g.tile_stride = 1;
        g.sign_base = (uint64_t)c->rank << c->nl;
        sh.p0 = c->buf[buf];
        sh.p1 = nullptr;
        return VQE_OK;
    }
    const int partner = c->rank ^ (int)tp.gpat;
    const int lo = std::min(c->rank, partner), hi = std::max(c->rank, partner);
    g.sign_base = (uint64_t)lo << c->nl;
    if (tp.n_tiles >= 2) {  // the two ranks of a pair split the super-tiles: even ones to lo, odd ones to hi
        g.n_tiles = tp.n_tiles / 2;
        g.tile_stride = 2;
        g.tile_first = (c->rank == lo) ? 0 : 1;
    } else {
        g.n_tiles = (c->rank == lo) ? 1 : 0;
        g.tile_stride = 1;
        g.tile_first = 0;
    }
    sh.p0 = shard_ptr(c, buf, lo);
    sh.p1 = shard_ptr(c, buf, hi);
    if (!sh.p0 || !sh.p1)
        return fail(VQE_ERR_INVALID, "rank %d: buffer %d of rank %d is not attached (vqe_shard_attach_*)", c->rank, buf, partner);
    return VQE_OK;
}

// ---- qubit relabelling of a sharded state (see vqe_ctx::perm) ------------------------------------------------------
static inline bool perm_is_identity(const vqe_ctx* c) {
    for (int b = 0; b < c->n; ++b)
        if (c->perm[b] != b) return false;
    return true;
}
static inline void perm_reset(vqe_ctx* c) {
    for (int b = 0; b < 64; ++b) c->perm[b] = (uint8_t)b;
    c->swap_history.clear();
}
static inline uint64_t perm_mask(const uint8_t* perm, uint64_t m) {
    uint64_t o = 0;
    while (m) {
        const int b = __builtin_ctzll(m);
        o |= 1ull << perm[b];
        m &= m - 1;
    }
    return o;
}
// every rank of the state this context belongs to, when they live in this process (vqe_shard_attach_local); else the
// context alone (one process per GPU: every process issues the same calls)
static RankSet group_of(vqe_ctx* c) {
    RankSet rs;
    bool local_group = c->world > 1;
    for (int r = 0; r < c->world && local_group; ++r)
        if (r != c->rank && !c->peer_ctx[r]) local_group = false;
    if (!local_group) {
        rs.r.push_back(c);
        return rs;
    }
    for (int r = 0; r < c->world; ++r) rs.r.push_back(r == c->rank ? c : c->peer_ctx[r]);
    return rs;
}
// swap the contents of physical index bits gslot (a rank bit, >= nl) and lslot (a local bit) of buffer 0 on all ranks
static int swap_global_local(RankSet& rs, int gslot, int lslot) {
    vqe_ctx* c0 = rs.r[0];
    if (gslot < c0->nl || gslot >= c0->n || lslot < 0 || lslot >= c0->nl) return fail(VQE_ERR_INVALID, "bad qubit swap %d <-> %d", gslot, lslot);
    for (vqe_ctx* c : rs.r)
        if (c->real_layout != c0->real_layout) return fail(VQE_ERR_INVALID, "the ranks disagree on the layout of the state");
    int rc = rank_barrier(rs);  // everything issued so far has finished on both shards
    if (rc) return rc;
    const int gbit = 1 << (gslot - c0->nl);
    for (vqe_ctx* c : rs.r) {
        CK(cudaSetDevice(c->device));
        rc = ensure_buf(c, VQE_BUF_PSI);
        if (rc) return rc;
        const int partner = c->rank ^ gbit;
        const int lo = std::min(c->rank, partner), hi = std::max(c->rank, partner);
        double2* plo = shard_ptr(c, VQE_BUF_PSI, lo);
        double2* phi = shard_ptr(c, VQE_BUF_PSI, hi);
        if (!plo || !phi) return fail(VQE_ERR_INVALID, "rank %d: the state of rank %d is not attached (vqe_shard_attach_*)", c->rank, partner);
        const uint64_t n_pairs = c->n_amp >> 1, half = n_pairs >> 1;
        const uint64_t first = c->rank == lo ? 0 : half, count = c->rank == lo ? half : n_pairs - half;
        const int blocks = (int)std::min<uint64_t>((count + 1023) / 1024, (uint64_t)c->sm_count * 8);
        ProfScope prof(c, 4);
        if (c->real_layout)  // the state is kept as n_amp doubles: the same exchange on 8-byte elements
            k_swap_global_local<double><<<std::max(1, blocks), 256, 0, c->stream>>>(reinterpret_cast<double*>(plo), reinterpret_cast<double*>(phi),
                                                                                     (uint32_t)lslot, first, count);
        else
            k_swap_global_local<double2><<<std::max(1, blocks), 256, 0, c->stream>>>(plo, phi, (uint32_t)lslot, first, count);
        c->launches++;
        c->n_swaps++;
        c->swap_bytes += count * (c->real_layout ? sizeof(double) : sizeof(double2));
        CK(cudaGetLastError());
    }
    rc = rank_barrier(rs);  // the exchange is complete before anyone touches its shard again
    if (rc) return rc;
    for (vqe_ctx* c : rs.r) {
        int a = -1, b = -1;
        for (int q = 0; q < c->n; ++q) {
            if (c->perm[q] == gslot) a = q;
            if (c->perm[q] == lslot) b = q;
        }
        c->perm[a] = (uint8_t)lslot;
        c->perm[b] = (uint8_t)gslot;
        c->swap_history.push_back({gslot, lslot});
    }
    return VQE_OK;
}
// Undo all swaps (in reverse order: a swap is its own inverse): afterwards buffer 0 is labelled as the caller labels it.
static int restore_labelling(RankSet& rs) {
    vqe_ctx* c = rs.r[0];
    if (c->world == 1) return VQE_OK;
    while (!c->swap_history.empty()) {
        const std::pair<int, int> sw = c->swap_history.back();
        int rc = swap_global_local(rs, sw.first, sw.second);  // pushes the swap again ...
        if (rc) return rc;
        for (vqe_ctx* r : rs.r) {                                // ... so two entries go
            r->swap_history.pop_back();
            r->swap_history.pop_back();
        }
    }
    if (!perm_is_identity(c)) return fail(VQE_ERR_INVALID, "qubit relabelling could not be undone");
    return VQE_OK;
}
static int need_caller_labelling(vqe_ctx* c) {
    if (c->world == 1 || c->swap_history.empty()) return VQE_OK;
    RankSet rs = group_of(c);
    return restore_labelling(rs);
}
static int need_caller_labelling(RankSet& rs) {
    if (rs.r.empty() || !rs.r[0] || rs.r[0]->world == 1 || rs.r[0]->swap_history.empty()) return VQE_OK;
    if (rs.r.size() == 1) return need_caller_labelling(rs.r[0]);
    return restore_labelling(rs);
}

// ---- tensor maps ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
        cudaGetLastError();
    }
    return fn;
}
struct TmaShape {
    bool ok = false;
    bool swizzled = false;
    TmaGeom tg;
    cuuint64_t gdim[5];
    cuuint64_t gstride[4];
    cuuint32_t box[5];
};
// Host-only: the tensor-map shape of a local tile plan (see TmaGeom).  ok = false: keep the per-segment bulk copies.
// rl: real layout (8-byte elements, one double per amplitude) instead of interleaved complex (two doubles per amplitude)
static TmaShape plan_tma(const TilePlan& tp, bool swizzle, bool rl = false) {
    TmaShape sh;
    memset(&sh.tg, 0, sizeof sh.tg);
    sh.swizzled = false;
    const int r1 = rl ? 1 : 0;
    if (tp.vbit || tp.bits.empty() || tp.nl > 36) return sh;
    if (swizzle && tp.lbits < 3 + r1) return sh;  // the 128-byte swizzle needs the three (real layout: four) lowest index bits in the tile
    // runs of adjacent tile bits; a run is cut where a box would exceed 256 elements (dimension 0 counts doubles)
    struct Run { int start, len; };
    std::vector<Run> runs;
    for (int b : tp.bits) {
        const bool adjacent = !runs.empty() && runs.back().start + runs.back().len == b;
        // the run at bit 0 is dimension 0 and counts doubles: at most 2^(7+1) of them; swizzled: exactly 8 amplitudes = 128 bytes
        // (real layout: one double per amplitude, so one index bit more in both cases)
        const int cap = (runs.size() == 1 && runs[0].start == 0) ? (swizzle ? 3 + r1 : 7 + r1) : 8;
        if (adjacent && runs.back().len < cap) runs.back().len++;
        else runs.push_back({b, 1});
    }
    sh.swizzled = swizzle;
    // dimensions: breakpoints at 0 and at the start of the first (up to) five runs
    std::vector<int> starts, blog;
    size_t used_runs = 0;
    if (runs[0].start != 0) { starts.push_back(0); blog.push_back(0); }
    while (used_runs < runs.size() && starts.size() < 5) {
        starts.push_back(runs[used_runs].start);
        blog.push_back(runs[used_runs].len);
        ++used_runs;
    }
    // tile bits of the remaining runs select the request
    std::vector<int> extra;
    for (size_t r = used_runs; r < runs.size(); ++r)
        for (int k = 0; k < runs[r].len; ++k) extra.push_back(runs[r].start + k);
    if (extra.size() > 4) return sh;
    const int nd = (int)starts.size();
    for (int i = 0; i < 5; ++i) {
        if (i < nd) {
            const int lo = starts[i], hi = (i + 1 < nd) ? starts[i + 1] : tp.nl;
            const int span = hi - lo;
            if (span > 31 || (i == 0 && span > 30)) {
                return sh;  // (a dimension holds at most 2^32 elements; such plans keep the bulk-copy path)
            }
            sh.tg.shift[i] = (uint32_t)lo;
            sh.tg.cmask[i] = (uint32_t)((1ull << span) - 1ull);
            sh.gdim[i] = (i == 0 && !rl) ? (2ull << span) : (1ull << span);
            sh.box[i] = (i == 0 && !rl) ? (2u << blog[i]) : (1u << blog[i]);
            if (i > 0) sh.gstride[i - 1] = (rl ? 8ull : 16ull) << lo;
        } else {
            sh.tg.shift[i] = 0;
            sh.tg.cmask[i] = 0;  // coordinate 0
            sh.gdim[i] = 1;
            sh.box[i] = 1;
            sh.gstride[i - 1] = (rl ? 8ull : 16ull) << tp.nl;
        }
    }
    sh.tg.n_req = 1u << extra.size();
    sh.tg.req_amps = (1u << tp.tbits) >> extra.size();
    for (uint32_t r = 0; r < sh.tg.n_req; ++r) {
        uint64_t bits = 0;
        for (size_t k = 0; k < extra.size(); ++k)
            if ((r >> k) & 1u) bits |= 1ull << extra[k];
        sh.tg.req_bits[r] = bits;
    }
    sh.ok = true;
    return sh;
}
// Fill g.tma / g.tg and encode the tensor map of buffer `ptr` for this plan; leaves g.tma = 0 when the plan has no
// tensor-map form, the driver entry point is missing, or VQE_TMA=0.
// would make_tmap(tp, ..., swizzle) succeed?  (plan-time decision of the lean Pauli-sum passes, whose entry tables are
// pre-swizzled)
static bool tma_available(const TilePlan& tp, bool swizzle, bool rl = false) {
    return env_int("VQE_TMA", 1) && tmap_encoder() && plan_tma(tp, swizzle, rl).ok;
}
// the shared-memory swizzle make_tmap(tp, ..., -1, true) will choose for a real-layout tile (0x70 or 0)
static uint32_t rl_tile_swz(const TilePlan& tp) {
    const bool want = env_int("VQE_SWIZZLE", 1) != 0;
    TmaShape sh = plan_tma(tp, want, true);
    if (!sh.ok && want) sh = plan_tma(tp, false, true);
    return sh.ok && sh.swizzled ? 0x70u : 0u;
}
// swizzle: -1 = swizzled when the plan allows it, 0 = natural layout, 1 = swizzled or nothing
static void make_tmap(const TilePlan& tp, double2* ptr, TileGeom& g, CUtensorMap* map, int swizzle = -1, bool rl = false) {
    memset(map, 0, sizeof *map);
    g.tma = 0;
    g.swz = 0;
    g.rl = rl ? 1u : 0u;
    memset(&g.tg, 0, sizeof g.tg);
    if (!env_int("VQE_TMA", 1) || !ptr) return;
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return;
    if (!env_int("VQE_SWIZZLE", 1) && swizzle < 0) swizzle = 0;
    TmaShape sh = plan_tma(tp, swizzle != 0, rl);
    if (!sh.ok && swizzle < 0) sh = plan_tma(tp, false, rl);
    if (!sh.ok) return;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void*)ptr, sh.gdim, sh.gstride, sh.box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sh.swizzled ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return;
    g.tg = sh.tg;
    g.tma = 1;
    g.swz = sh.swizzled ? 0x70u : 0u;
}

// Host-only check of the tensor-map form of a tile plan (no CUDA call; CPU tests): emulates the TMA box traversal
// (dimension 0 fastest, element = one double) for every request of up to `max_tiles` tiles and compares the global
// element index of every tile element with the gather address the per-segment path uses.  Returns the number of
// mismatching elements, or -1 when the plan has no tensor-map form.  *n_req / *rank_used describe the shape.
static int tma_check_impl(int n_local, uint64_t need_mask, int tile_bits, int low_bits, int max_tiles, int32_t* n_req,
                          int32_t* dims_used, bool rl);
extern "C" int vqe_debug_tma_check(int n_local, uint64_t need_mask, int tile_bits, int low_bits, int max_tiles, int32_t* n_req,
                                   int32_t* dims_used) {
    return tma_check_impl(n_local, need_mask, tile_bits, low_bits, max_tiles, n_req, dims_used, false);
}
// the same for the REAL LAYOUT of the state buffer (n_amp contiguous doubles, tile elements of 8 bytes)
extern "C" int vqe_debug_tma_check_rl(int n_local, uint64_t need_mask, int tile_bits, int low_bits, int max_tiles, int32_t* n_req,
                                      int32_t* dims_used) {
    return tma_check_impl(n_local, need_mask, tile_bits, low_bits, max_tiles, n_req, dims_used, true);
}
static int tma_check_impl(int n_local, uint64_t need_mask, int tile_bits, int low_bits, int max_tiles, int32_t* n_req,
                          int32_t* dims_used, bool rl) {
    if (n_local < 1 || n_local > 40) return fail(VQE_ERR_INVALID, "bad n_local");
    if (tile_bits < 1 || tile_bits > 12) tile_bits = 12;
    const bool want_swz = max_tiles < 0;  // negative max_tiles: check the 128-byte-swizzled shape
    max_tiles = max_tiles < 0 ? -max_tiles : max_tiles;
    const TilePlan tp = make_plan(n_local, need_mask, tile_bits, std::max(0, std::min(low_bits, tile_bits)), 0);
    const TmaShape sh = plan_tma(tp, want_swz, rl);
    if (!sh.ok) return -1;
    if (want_swz && sh.box[0] * 8 != 128) return 1 << 30;  // the swizzle atom is exactly 128 bytes wide
    if (n_req) *n_req = (int32_t)sh.tg.n_req;
    if (dims_used) {
        int d = 0;
        for (int i = 0; i < 5; ++i) d += sh.gdim[i] > 1 ? 1 : 0;
        *dims_used = d;
    }
    // constraints of cuTensorMapEncodeTiled
    for (int i = 0; i < 5; ++i) {
        if (sh.box[i] == 0 || sh.box[i] > 256 || sh.gdim[i] == 0 || sh.gdim[i] > (1ull << 32)) return 1 << 30;
        if (i > 0 && ((sh.gstride[i - 1] & 15ull) || sh.gstride[i - 1] >= (1ull << 40))) return 1 << 30;
    }
    if ((sh.box[0] * 8) % 16) return 1 << 30;
    const uint32_t ts = 1u << tp.tbits, lmask = (1u << tp.lbits) - 1u;
    int bad = 0;
    const uint64_t n_tiles = std::min<uint64_t>(tp.n_tiles, (uint64_t)std::max(1, max_tiles));
    for (uint64_t k = 0; k < n_tiles; ++k) {
        const uint64_t t = (n_tiles == tp.n_tiles) ? k : (k * 2654435761ull) % tp.n_tiles;  // spread the sample
        uint64_t base = 0, v = t, m = tp.comp_mask;
        while (m) {
            const uint64_t low = m & (0 - m);
            if (v & 1) base |= low;
            v >>= 1;
            m ^= low;
        }
        uint32_t local = 0;  // tile-local element (in doubles: 2 per amplitude)
        for (uint32_t r = 0; r < sh.tg.n_req; ++r) {
            const uint64_t idx = base | sh.tg.req_bits[r];
            int64_t c[5];
            for (int i = 0; i < 5; ++i) c[i] = (int64_t)((uint32_t)(idx >> sh.tg.shift[i]) & sh.tg.cmask[i]);
            if (!rl) c[0] <<= 1;
            for (uint32_t e4 = 0; e4 < sh.box[4]; ++e4)
                for (uint32_t e3 = 0; e3 < sh.box[3]; ++e3)
                    for (uint32_t e2 = 0; e2 < sh.box[2]; ++e2)
                        for (uint32_t e1 = 0; e1 < sh.box[1]; ++e1)
                            for (uint32_t e0 = 0; e0 < sh.box[0]; ++e0, ++local) {
                                const uint64_t byte = (uint64_t)(c[0] + e0) * 8ull + (uint64_t)(c[1] + e1) * sh.gstride[0] +
                                                      (uint64_t)(c[2] + e2) * sh.gstride[1] + (uint64_t)(c[3] + e3) * sh.gstride[2] +
                                                      (uint64_t)(c[4] + e4) * sh.gstride[3];
                                if ((uint64_t)(c[0] + e0) >= sh.gdim[0] || (uint64_t)(c[1] + e1) >= sh.gdim[1] ||
                                    (uint64_t)(c[2] + e2) >= sh.gdim[2] || (uint64_t)(c[3] + e3) >= sh.gdim[3] ||
                                    (uint64_t)(c[4] + e4) >= sh.gdim[4]) { ++bad; continue; }
                                const uint32_t kk = rl ? local : local >> 1;  // amplitude inside the tile
                                const uint64_t amp = base | tp.scat[kk >> tp.lbits] | (uint64_t)(kk & lmask);
                                const uint64_t want = rl ? amp * 8ull : amp * 16ull + (local & 1u) * 8ull;
                                if (byte != want) ++bad;
                            }
        }
        if (local != (rl ? ts : 2 * ts)) bad += 1000000;
    }
    return bad;
}

extern "C" int vqe_n_qubits(const vqe_ctx* c) { return c ? c->n : 0; }
extern "C" int vqe_state_layout(const vqe_ctx* c) { return (c && c->real_layout) ? 1 : 0; }
extern "C" uint64_t vqe_launch_count(const vqe_ctx* c) { return c ? c->launches : 0; }
extern "C" int vqe_profile_enable(vqe_ctx* c, int on) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    c->profiling = on != 0;
    return VQE_OK;
}
extern "C" int vqe_profile_read(vqe_ctx* c, int which, double* ms_total, uint64_t* launches, int reset) {
    if (!c || which < 0 || which > 5) return fail(VQE_ERR_INVALID, "bad profile slot");
    resolve_profile(c);
    if (ms_total) *ms_total = c->prof[which].ms;
    if (launches) *launches = c->prof[which].launches;
    if (reset) c->prof[which] = KernelProf();
    return VQE_OK;
}

extern "C" int vqe_timer_begin(vqe_ctx* c) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev0, c->stream));
    return VQE_OK;
}
extern "C" int vqe_timer_end(vqe_ctx* c, double* ms) {
    if (!c || !ms) return fail(VQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, c->ev0, c->ev1));
    *ms = f;
    return VQE_OK;
}
extern "C" int vqe_transfer_bytes(vqe_ctx* c, uint64_t* h2d, uint64_t* d2h, int reset) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    if (h2d) *h2d = c->h2d_bytes;
    if (d2h) *d2h = c->d2h_bytes;
    if (reset) c->h2d_bytes = c->d2h_bytes = 0;
    return VQE_OK;
}

extern "C" int vqe_peer_bytes(vqe_ctx* c, uint64_t* gathered, int reset) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    if (gathered) *gathered = c->gather_bytes;
    if (reset) c->gather_bytes = 0;
    return VQE_OK;
}

extern "C" int vqe_relabel_stats(vqe_ctx* c, uint64_t* swaps, uint64_t* swap_bytes, int reset) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    if (swaps) *swaps = c->n_swaps;
    if (swap_bytes) *swap_bytes = c->swap_bytes;
    if (reset) c->n_swaps = c->swap_bytes = 0;
    return VQE_OK;
}

static int grid_1d(const vqe_ctx* c, uint64_t n_amp, int threads) {
    uint64_t want = (n_amp + threads - 1) / threads;
    uint64_t cap = (uint64_t)c->sm_count * 8;
    return (int)std::max<uint64_t>(1, std::min(want, cap));
}

extern "C" int vqe_set_basis_state(vqe_ctx* c, uint64_t index) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    if (index >> c->n) return fail(VQE_ERR_INVALID, "basis index %llu >= 2^%d", (unsigned long long)index, c->n);
    CK(cudaSetDevice(c->device));
    // sharded: only the rank that owns the index holds the 1 (an out-of-range local index leaves the shard zero)
    const uint64_t local = ((int)(index >> c->nl) == c->rank) ? (index & (c->n_amp - 1)) : ~0ull;
    if (c->real_layout_ok)
        k_zero_set_real<<<grid_1d(c, c->n_amp, 256), 256, 0, c->stream>>>(reinterpret_cast<double*>(c->buf[0]), c->n_amp, local);
    else
        k_zero_set<<<grid_1d(c, c->n_amp, 256), 256, 0, c->stream>>>(c->buf[0], c->n_amp, local);
    c->walk_desc = false;   // written by ascending address
    c->real_layout = c->real_layout_ok;
    c->psi_real = true;
    perm_reset(c);  // a fresh state is labelled as the caller labels it
    c->launches++;
    CK(cudaGetLastError());
    return VQE_OK;
}

extern "C" int vqe_set_state(vqe_ctx* c, int b, const double* re_im) {
    if (!c || !re_im) return fail(VQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    int rc = ensure_buf(c, b);
    if (rc) return rc;
    c->h2d_bytes += c->n_amp * sizeof(double2);
    if (b == VQE_BUF_PSI) {
        c->psi_real = c->real_layout = false;  // overwritten as a whole, in complex form, in the caller's labelling
        perm_reset(c);
    }
    CK(cudaMemcpyAsync(c->buf[b], re_im, c->n_amp * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return VQE_OK;
}
extern "C" int vqe_get_state(vqe_ctx* c, int b, double* re_im) {
    if (!c || !re_im) return fail(VQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    int rc = ensure_buf(c, b);
    if (rc) return rc;
    if (b == VQE_BUF_PSI) {
        rc = need_caller_labelling(c);
        if (rc) return rc;
    }
    if (b == VQE_BUF_PSI && c->real_layout) {
        // real layout: the n_amp real parts are copied into the upper half of the caller's array and interleaved there
        double* tmp = re_im + c->n_amp;
        c->d2h_bytes += c->n_amp * sizeof(double);
        CK(cudaMemcpyAsync(tmp, c->buf[b], c->n_amp * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (uint64_t i = 0; i < c->n_amp; ++i) {  // ascending: slot 2i is below the unread reals n_amp + i.. for every i
            const double v = tmp[i];
            re_im[2 * i] = v;
            re_im[2 * i + 1] = 0.0;
        }
        return vqe_shard_status(c);
    }
    c->d2h_bytes += c->n_amp * sizeof(double2);
    CK(cudaMemcpyAsync(re_im, c->buf[b], c->n_amp * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return vqe_shard_status(c);
}
extern "C" int vqe_copy_buffer(vqe_ctx* c, int dst, int src) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    CK(cudaSetDevice(c->device));
    int rc = ensure_buf(c, dst);
    if (rc) return rc;
    rc = ensure_buf(c, src);
    if (rc) return rc;
    if (dst == src) return VQE_OK;
    if (src == VQE_BUF_PSI) {
        rc = need_caller_labelling(c);
        if (rc) return rc;
    }
    rc = ensure_complex(c, src);
    if (rc) return rc;
    if (dst == VQE_BUF_PSI) {
        c->psi_real = c->real_layout = false;
        perm_reset(c);  // overwritten with a vector in the caller's labelling
    }
    CK(cudaMemcpyAsync(c->buf[dst], c->buf[src], c->n_amp * sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
    return VQE_OK;
}
extern "C" int vqe_buffer_ptr(vqe_ctx* c, int b, void** p, uint64_t* n_amp) {
    if (!c || !p) return fail(VQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    if (b < 0 || b > 3) return fail(VQE_ERR_INVALID, "bad buffer id %d", b);
    int rc = ensure_buf(c, b);
    if (rc) return rc;
    rc = ensure_complex(c, b);
    if (rc == VQE_OK && b == VQE_BUF_PSI) rc = need_caller_labelling(c);
    if (rc) return rc;
    *p = c->buf[b];
    if (b == VQE_BUF_PSI) c->psi_real = false;  // the caller may write through the pointer
    if (n_amp) *n_amp = c->n_amp;
    return VQE_OK;
}
extern "C" int vqe_synchronize(vqe_ctx* c) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return vqe_shard_status(c);
}

// ---- generic op program ---------------------------------------------------------------------
struct HostOp {
    int kind = 0;
    uint64_t x = 0, z = 0;  // ROT: masks | GATE1: x = bit | CNOT: x = target bit, z = control bit
    double c = 0.0, s = 0.0;
    int ny = 0;
    double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double ang = 0.0;    // ROT: the rotation angle itself (collapsed runs add angles)
    // PLANE: tabulated plane rotation.  For every listed a-side pattern p (bits inside x, highest x bit clear) the
    // pairs (l, l ^ x) with (l & x) == p rotate:  a' = c a - s b,  b' = s a + c b.  Unlisted patterns are untouched.
    std::vector<uint64_t> ppat;
    std::vector<double> pcos, psin;
};

// Persistent grid of a pass kernel.  The CTAs stride over the tiles, so with `cap` resident CTAs a pass takes
// ceil(n_tiles / cap) rounds; the grid is then shrunk to ceil(n_tiles / rounds) so that every CTA works in every round
// (2048 tiles on 444 slots: 410 CTAs x 5 tiles instead of a last round with 39 % of the slots idle) with VQE_GRID_BALANCE=1.
// Measured: no gain for the HBM-bound rotation passes, 4 % slower expectation passes (they want every slot): off by default.
static int tile_grid(const vqe_ctx* c, uint64_t n_tiles, int ctas_per_sm = 0) {
    uint64_t cap = (uint64_t)c->sm_count * (ctas_per_sm ? ctas_per_sm : c->ctas_per_sm);
    if (n_tiles > cap && env_int("VQE_GRID_BALANCE", 0) != 0) {
        const uint64_t rounds = (n_tiles + cap - 1) / cap;
        cap = (n_tiles + rounds - 1) / rounds;
    }
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, cap));
}

// ---- planning of an ordered op list into tile passes (pure host code, no CUDA) -----------------
struct OpPass {
    TilePlan tp;
    size_t op_begin, op_end;            // in the dev op array
    size_t sup_begin = 0, sup_end = 0;  // fast passes: shared-memory round trips (orbits) ...
    size_t sub_begin = 0, sub_end = 0;  // ... and their sub-runs
    size_t col_begin = 0, col_end = 0;  // collapsed runs ...
    size_t ent_begin = 0, ent_end = 0;  // ... and their (pattern, cos, sin) entries
    bool fast = false;
    bool has_imag = false;              // some rotation of a fast pass has a +-i phase
    // gather-form peer pass (see GatherGeom): groups of 4 partner amplitudes each rank's results depend on
    bool gather = false;
    std::vector<uint16_t> need_lo, need_hi;  // needed by the lower / the higher rank of a pair
    double pass_scale = 1.0;            // fast passes: product of the cosines not yet applied by a run
};
struct OpPlan {
    std::vector<OpPass> passes;
    std::vector<DevOp> dops;
    std::vector<DevSuper> dsupers;
    std::vector<DevSub> dsubs;
    std::vector<DevCol> dcols;
    std::vector<DevColEntry> dents;
    std::vector<double> mats;
    // what a re-use of the plan with new angles needs (see refresh_plan)
    struct ColRecipe {
        uint32_t first_op, n_str;   // strings of the run: dops[first_op .. first_op + n_str)
        uint32_t pat_begin, n_pat;  // its a-side patterns in col_pats
    };
    struct ColPat {
        uint32_t signmask;          // bit q set: string q enters the pattern's angle with a minus sign
        int32_t ent;                // index into dents, or -1 when the angle vanishes on this pattern
    };
    std::vector<ColRecipe> recipes;
    std::vector<ColPat> col_pats;
    std::vector<double> rot_cos;    // cosine of every dop (fast passes)
    bool cacheable = true;
    int tile_bits = 12;             // 13: planned for the real layout of the state (tiles of 2^13 doubles = 64 KiB)
    // item table of the real-layout collapsed-run kernel (k_col_tab), built on first use (launch_plan) and kept on the
    // device per context while the plan lives: angle-independent, so a cached plan uploads it once
    mutable uint64_t gen = 0;                  // identity of the table (0: not built)
    mutable std::vector<uint32_t> coltab;      // all passes
    mutable std::vector<size_t> coltab_off;    // first word of every pass
    mutable std::vector<uint32_t> coltab_swz;  // the tile swizzle every pass was tabulated for
    mutable std::vector<uint16_t> coltab16;    // the same items in 16 bits (k_col_stab), every pass padded to a multiple of 8 items
    mutable std::vector<size_t> coltab16_off;
    mutable std::vector<uint32_t> coltab16_seg;  // segments per pass
    mutable bool coltab_has32 = false;           // the 32-bit words were built too (k_col_tab, host interpreter form 0)
};

static bool fast_eligible(const HostOp& h) {
    return (h.kind == OP_ROT && h.x != 0 && fabs(h.c) >= 0.3) || h.kind == OP_PLANE;
}

// Greedy, order-preserving: a pass takes consecutive ops while the union of their LOCAL X bits fits the tile and
// their GLOBAL X parts are 0 or one common pattern m (then the pass is a peer pass between ranks r and r^m and
// the tile loses one local bit to the virtual shard bit).
static int plan_ops(int n, int nl, int tile_bits, int low_bits, int threads_cfg, const std::vector<HostOp>& ops,
                    OpPlan& out) {
    std::vector<OpPass>& passes = out.passes;
    std::vector<DevOp>& dops = out.dops;
    std::vector<DevSuper>& dsupers = out.dsupers;
    std::vector<DevSub>& dsubs = out.dsubs;
    std::vector<DevCol>& dcols = out.dcols;
    std::vector<DevColEntry>& dents = out.dents;
    std::vector<double>& mats = out.mats;
    dops.reserve(ops.size());
    size_t i = 0;
    const uint64_t lfull = (1ull << nl) - 1ull;
    const int lb = std::min(low_bits, std::min(tile_bits, nl));
    const uint64_t lowmask = (1ull << lb) - 1ull;
    (void)n;
    const int lb_default = lb;
    const uint64_t lowmask_default = lowmask;
    while (i < ops.size()) {
        uint64_t need = 0, pat = 0, ctrl_glob = 0;
        size_t j = i;
        const bool fast0 = fast_eligible(ops[i]);
        // low-bit floor of this pass: the default unless the first operation alone needs more room
        int lb = fit_low_bits(ops[i].x & lfull, lb_default, tile_bits, nl, (ops[i].x >> nl) != 0);
        if (lb < 0)
            return fail(VQE_ERR_INVALID, "operation %zu touches %d local X-bits, more than a %d-bit tile can hold", i,
                        popc64(ops[i].x & lfull), std::min(tile_bits, nl));
        const uint64_t lowmask = lb == lb_default ? lowmask_default : ((1ull << lb) - 1ull);
        size_t tab_bytes = 0;  // shared-memory tables of the pass next to the 64 KiB tile (227 KiB per CTA at most)
        while (j < ops.size() && j - i < OPTAB_CAP) {
            if (fast_eligible(ops[j]) != fast0) break;  // a pass is either all-fast or general
            {
                // worst case per op: RotOp + its own segment descriptor + sub-run; a plane rotation brings its table
                const size_t op_bytes = ops[j].kind == OP_PLANE
                                            ? sizeof(RotOp) + sizeof(DevSuper) + sizeof(DevCol) + 4 + ops[j].ppat.size() * sizeof(DevColEntry)
                                            : sizeof(RotOp) + sizeof(DevSuper) + sizeof(DevSub);
                if (j > i && tab_bytes + op_bytes > 96 * 1024) break;
                tab_bytes += op_bytes;
            }
            const uint64_t xg = ops[j].x >> nl;
            uint64_t npat = pat;
            if (xg) {
                if (pat && pat != xg) break;
                npat = xg;
            }
            uint64_t nctrl = ctrl_glob;
            if (ops[j].kind == OP_CNOT) nctrl |= ops[j].z >> nl;
            // a global CNOT control inside the pattern must BE the pattern (then it is the virtual bit)
            if (npat && (nctrl & npat) && popc64(npat) > 1) break;
            if (npat && (nctrl & npat) && (nctrl & npat) != npat) break;
            const uint64_t u = need | (ops[j].x & lfull);
            // shrink the low-bit floor when the shard is tiny (tests): the planner only needs "fits"
            if (!plan_fits(u, lowmask, tile_bits, nl, npat != 0)) break;
            // 13-bit tiles exist only in the real layout, which is staged by tensor-map requests: the pass must keep one
            if (tile_bits > 12 && j > i && !plan_tma(make_plan(nl, u, tile_bits, lb, npat), false, true).ok) break;
            need = u;
            pat = npat;
            ctrl_glob = nctrl;
            ++j;
        }
        if (j == i)
            return fail(VQE_ERR_INVALID, "operation %zu touches %d local X-bits, more than a %d-bit tile can hold", i,
                        popc64(ops[i].x & lfull), std::min(tile_bits, nl));
        OpPass p;
        p.tp = make_plan(nl, need, tile_bits, lb, pat);
        p.op_begin = dops.size();
        for (size_t k = i; k < j; ++k) {
            const HostOp& h = ops[k];
            DevOp d;
            memset(&d, 0, sizeof d);
            d.kind = h.kind;
            if (h.kind == OP_ROT) {
                d.lx = plan_lx(h.x, p.tp);
                d.lz = plan_lz(h.z, p.tp);
                d.zout = plan_zout(h.z, p.tp);
                d.c = h.c;
                d.s = h.s;
                d.k4 = (uint32_t)((h.ny + 3) & 3);
                d.nyodd = h.ny & 1;
                d.hb = d.lx ? 31 - __builtin_clz(d.lx) : 0;
                d.run = 1;
            } else if (h.kind == OP_PLANE) {
                d.kind = OP_ROTF;          // lives in fast passes; emitted as a ready-made collapsed run
                d.lx = plan_lx(h.x, p.tp);
                d.hb = 31 - __builtin_clz(d.lx);
                d.c = 1.0;
                d.nyodd = 0xffffffffu;     // marker: tabulated plane rotation (see is_plane below)
            } else if (h.kind == OP_GATE1) {
                d.lx = plan_lx(h.x, p.tp);
                d.hb = 31 - __builtin_clz(d.lx);
                d.mat = (uint32_t)(mats.size() / 8);
                mats.insert(mats.end(), h.m, h.m + 8);
            } else {  // CNOT: x = target bit, z = control bit
                d.lx = plan_lx(h.x, p.tp);
                d.hb = 31 - __builtin_clz(d.lx);
                if ((h.z & lfull) & p.tp.tile_mask) d.lz = pext_local(h.z, p.tp);
                else if (p.tp.vbit && ((h.z >> nl) & p.tp.gpat)) d.lz = 1u << (p.tp.tbits - 1);
                else d.zout = h.z;
            }
            dops.push_back(d);
        }
        p.op_end = dops.size();
        // classify: small/moderate angles with lx != 0 take the tangent-form fast path
        const int threads_p = (int)std::min<uint64_t>(threads_cfg, std::max<uint64_t>(32, (1ull << p.tp.tbits) / 2));
        const int tshift = 31 - __builtin_clz((unsigned)threads_p);
        const int n_j = (int)std::max<uint64_t>(1, ((1ull << p.tp.tbits) / 2) / threads_p);
        for (size_t k = p.op_begin; k < p.op_end; ++k) {
            DevOp& d = dops[k];
            if (d.kind != OP_ROT || d.lx == 0 || fabs(d.c) < 0.3 || n_j > 16) continue;
            d.kind = OP_ROTF;
            d.imag = (d.k4 & 1u);
            double tn = d.s / d.c;
            if (d.k4 >> 1) tn = -tn;
            d.s = tn;  // tangent, unit-phase sign folded in
            uint32_t jm = 0;
            for (int jj = 0; jj < n_j; ++jj) {
                uint32_t pj = (uint32_t)jj << tshift;
                uint32_t uj = ((pj >> d.hb) << (d.hb + 1)) | (pj & ((1u << d.hb) - 1u));
                if (__builtin_popcount(uj & d.lz) & 1) jm |= 1u << jj;
            }
            d.jmask = jm;
        }
        std::vector<double> rot_cos(p.op_end - p.op_begin);  // per-rotation cosines (the run heads get overwritten below)
        for (size_t k = p.op_begin; k < p.op_end; ++k) rot_cos[k - p.op_begin] = dops[k].c;
        out.rot_cos.insert(out.rot_cos.end(), rot_cos.begin(), rot_cos.end());
        // run lengths of consecutive same-lx rotations of the same kind; fast runs carry prod(c) in the head
        for (size_t k = p.op_begin; k < p.op_end;) {
            if (dops[k].kind != OP_ROT && dops[k].kind != OP_ROTF) { ++k; continue; }
            size_t e = k + 1;
            while (e < p.op_end && dops[e].kind == dops[k].kind && dops[e].lx == dops[k].lx &&
                   (dops[k].kind != OP_ROTF || dops[e].imag == dops[k].imag)) ++e;
            dops[k].run = (uint32_t)(e - k);
            if (dops[k].kind == OP_ROTF) {
                double prod = 1.0;
                for (size_t q = k; q < e; ++q) prod *= dops[q].c;
                dops[k].c = prod;
            }
            k = e;
        }
        p.fast = true;
        bool any_plane = false;
        for (size_t k = p.op_begin; k < p.op_end; ++k) {
            if (dops[k].kind != OP_ROTF) p.fast = false;
            if (dops[k].nyodd == 0xffffffffu) any_plane = true;
        }
        auto is_plane = [&](size_t k) { return dops[k].nyodd == 0xffffffffu; };
        const uint32_t half_p = (1u << p.tp.tbits) >> 1;
        const bool four = half_p == 4u * (uint32_t)threads_p;
        if (!(four || half_p <= (uint32_t)threads_p)) p.fast = false;  // k_tile_rot holds 4 pairs per thread, or 1
        if (any_plane && !p.fast)
            return fail(VQE_ERR_INVALID, "tabulated plane rotations need the default tile configuration");
        if (p.fast) {
            p.sup_begin = dsupers.size();
            p.sub_begin = dsubs.size();
            p.col_begin = dcols.size();
            p.ent_begin = dents.size();
            double pending = 1.0;  // cosines are applied once per pass (tile store) unless the product gets tiny
            // Try to collapse the same-X-mask run [k, e) into one plane rotation with a tabulated angle (see DevCol).
            // Returns 0 = not collapsible / not worth it, 1 = emitted, 2 = the angles cancel everywhere (identity).
            auto try_collapse = [&](size_t k, size_t e) -> int {
                const size_t R = e - k;
                if (R < 2) return 0;
                uint32_t D = 0;
                for (size_t q = k; q < e; ++q) {
                    if (dops[q].zout != dops[k].zout) return 0;
                    D |= dops[q].lz ^ dops[k].lz;
                }
                const uint32_t hb = dops[k].hb;
                const uint32_t E = D | (1u << hb);
                const int ne = __builtin_popcount(E);
                if (ne > 6) return 0;
                std::vector<uint32_t> epos;
                for (int b2 = 0; b2 < p.tp.tbits; ++b2)
                    if ((E >> b2) & 1u) epos.push_back((uint32_t)b2);
                double scale = 0.0;
                for (size_t q = k; q < e; ++q) scale += fabs(ops[i + (q - p.op_begin)].ang);
                std::vector<DevColEntry> ent;
                std::vector<OpPlan::ColPat> pats;  // recipe of the run: sign masks of all a-side patterns
                if (R > 32) out.cacheable = false;
                for (uint32_t pi = 0; pi < (1u << ne); ++pi) {
                    uint32_t pat = 0;
                    for (int b2 = 0; b2 < ne; ++b2)
                        if ((pi >> b2) & 1u) pat |= 1u << epos[b2];
                    if ((pat >> hb) & 1u) continue;  // a-side only
                    double F = 0.0;
                    uint32_t signmask = 0;
                    for (size_t q = k; q < e; ++q) {
                        const bool neg = ((dops[q].k4 >> 1) & 1u) != (uint32_t)(__builtin_popcount(pat & (dops[q].lz ^ dops[k].lz)) & 1);
                        if (neg && q - k < 32) signmask |= 1u << (q - k);
                        F += (neg ? -1.0 : 1.0) * ops[i + (q - p.op_begin)].ang;
                    }
                    const bool active = fabs(F) > 1e-15 * scale;  // else: exact cancellation up to rounding, identity
                    pats.push_back({signmask, active ? (int32_t)ent.size() : -1});
                    if (!active) continue;
                    DevColEntry en;
                    memset(&en, 0, sizeof en);
                    en.c = cos(F);
                    en.s = sin(F);
                    en.pat = pat;
                    ent.push_back(en);
                }
                auto record = [&](bool emitted) {
                    OpPlan::ColRecipe rcp = {(uint32_t)k, (uint32_t)R, (uint32_t)out.col_pats.size(), (uint32_t)pats.size()};
                    for (OpPlan::ColPat cp : pats) {
                        if (cp.ent >= 0) cp.ent = emitted ? cp.ent + (int32_t)dents.size() : -2;  // -2: active but not emitted
                        out.col_pats.push_back(cp);
                    }
                    out.recipes.push_back(rcp);
                };
                if (ent.empty()) {
                    record(false);
                    return 2;
                }
                // shared-memory budget of the pass for collapsed-run tables (the tile itself takes 64 KiB of the ~113)
                const size_t need_bytes = sizeof(DevCol) + 4 + sizeof(DevSuper) + ent.size() * sizeof(DevColEntry);
                if ((dcols.size() - p.col_begin) * (sizeof(DevCol) + 4 + sizeof(DevSuper)) +
                        (dents.size() - p.ent_begin) * sizeof(DevColEntry) + need_bytes > 24 * 1024)
                    return 0;
                const double cost_col = (double)ent.size() * (double)(1u << (p.tp.tbits - ne)) * 35.0;
                const double cost_seq = (double)R * (double)half_p * 6.0;
                // patterns with an exactly vanishing angle are left untouched by the collapsed form (structural zeros
                // stay exact), so it is preferred up to twice the sequential cost
                const bool has_identity = ent.size() < (size_t)(1u << (ne - 1));
                if (cost_col >= (has_identity ? 2.0 : 1.0) * cost_seq) return 0;
                DevCol co;
                memset(&co, 0, sizeof co);
                co.zout = dops[k].zout;
                co.lx = dops[k].lx;
                co.lz = dops[k].lz;
                co.n_active = (uint32_t)ent.size();
                co.nd = (uint32_t)ne;
                for (int b2 = 0; b2 < ne; ++b2) co.dpos[b2] = ~((1u << epos[b2]) - 1u);
                co.ent_begin = (uint32_t)(dents.size() - p.ent_begin);
                co.free_log = (uint32_t)(p.tp.tbits - ne);
                co.imag = dops[k].imag;
                if (co.imag) p.has_imag = true;
                DevSuper su;
                memset(&su, 0, sizeof su);
                su.sub_begin = (uint32_t)(dcols.size() - p.col_begin);
                su.sub_count = 0xffffffffu;
                su.cscale = 1.0;
                record(true);
                dcols.push_back(co);
                dents.insert(dents.end(), ent.begin(), ent.end());
                dsupers.push_back(su);
                return 1;
            };
            auto run_end_of = [&](size_t k) {
                size_t e = k + 1;
                while (e < p.op_end && !is_plane(e) && dops[e].lx == dops[k].lx && dops[e].imag == dops[k].imag) ++e;
                return e;
            };
            // a tabulated plane rotation is a ready-made collapsed run (no parity factor: lz = 0, zout = 0)
            auto emit_plane = [&](size_t k) -> int {
                const HostOp& h = ops[i + (k - p.op_begin)];
                const uint32_t lx = dops[k].lx;
                const int ne = __builtin_popcount(lx);
                if (ne > 6 || ne > p.tp.tbits) return fail(VQE_ERR_INVALID, "plane rotation on %d qubits (max 6)", ne);
                DevCol co;
                memset(&co, 0, sizeof co);
                co.lx = lx;
                co.n_active = (uint32_t)h.ppat.size();
                co.nd = (uint32_t)ne;
                int b3 = 0;
                for (int b2 = 0; b2 < p.tp.tbits; ++b2)
                    if ((lx >> b2) & 1u) co.dpos[b3++] = ~((1u << b2) - 1u);
                co.ent_begin = (uint32_t)(dents.size() - p.ent_begin);
                co.free_log = (uint32_t)(p.tp.tbits - ne);
                for (size_t q = 0; q < h.ppat.size(); ++q) {
                    DevColEntry en;
                    memset(&en, 0, sizeof en);
                    en.c = h.pcos[q];
                    en.s = h.psin[q];
                    en.pat = plan_lx(h.ppat[q], p.tp);
                    if ((en.pat >> dops[k].hb) & 1u)  // keep the a-side convention "highest tile X bit clear"
                        return fail(VQE_ERR_INVALID, "plane rotation pattern must leave the highest X bit clear");
                    dents.push_back(en);
                }
                if (h.ppat.empty()) return VQE_OK;  // identity
                DevSuper su;
                memset(&su, 0, sizeof su);
                su.sub_begin = (uint32_t)(dcols.size() - p.col_begin);
                su.sub_count = 0xffffffffu;
                su.cscale = 1.0;
                dcols.push_back(co);
                dsupers.push_back(su);
                return VQE_OK;
            };
            auto close_scale = [&](DevSuper& su, size_t first_sub) {
                for (size_t q = first_sub; q < dsubs.size(); ++q)
                    for (uint32_t w = 0; w < dsubs[q].len; ++w) pending *= rot_cos[dsubs[q].begin + w];
                su.cscale = 1.0;
                if (fabs(pending) < 1e-30) {  // unnormalised amplitudes have grown by 1e30: rescale now
                    su.cscale = pending;
                    pending = 1.0;
                }
            };
            size_t k = p.op_begin;
            std::vector<char> col_state(p.op_end - p.op_begin, 0);  // per run head: 0 unknown, 1 tried and refused
            while (k < p.op_end) {
                if (is_plane(k)) {
                    const int prc = emit_plane(k);
                    if (prc) return prc;
                    ++k;
                    continue;
                }
                {
                    const size_t e = run_end_of(k);
                    const int cr = try_collapse(k, e);
                    if (cr != 0) {
                        k = e;
                        continue;
                    }
                    col_state[k - p.op_begin] = 1;
                }
                DevSuper su;
                memset(&su, 0, sizeof su);
                su.sub_begin = (uint32_t)(dsubs.size() - p.sub_begin);
                su.hb_log = 31 - __builtin_clz(half_p ? half_p : 1u);
                const size_t first_sub = dsubs.size();
                if (!four) {  // one pair per thread: every same-X-mask run is its own round trip
                    const size_t e = run_end_of(k);
                    dsubs.push_back({(uint32_t)(k - p.op_begin), (uint32_t)(e - k), 1u, dops[k].imag});
                    if (dops[k].imag) p.has_imag = true;
                    su.e0 = dops[k].hb;
                    su.off[1] = dops[k].lx;
                    su.sub_count = 1;
                    close_scale(su, first_sub);
                    dsupers.push_back(su);
                    k = e;
                    continue;
                }
                // orbit basis: v = X-masks as given, r = reduced forms (zero at the earlier pivots), T = coordinates of r in v
                uint32_t v[3] = {0, 0, 0}, rr[3] = {0, 0, 0}, T[3] = {0, 0, 0}, piv[3] = {0, 0, 0};
                int dim = 0;
                auto reduce = [&](uint32_t w, uint32_t& coord) {
                    coord = 0;
                    for (int a2 = 0; a2 < dim; ++a2)
                        if ((w >> piv[a2]) & 1u) {
                            w ^= rr[a2];
                            coord ^= T[a2];
                        }
                    return w;
                };
                auto pick_pivot = [&](uint32_t res) {
                    for (int b2 = p.tp.tbits - 1; b2 >= 5; --b2)
                        if ((res >> b2) & 1u) return (uint32_t)b2;  // bits >= 5 keep the shared-memory accesses conflict free
                    return (uint32_t)(31 - __builtin_clz(res));
                };
                while (k < p.op_end) {
                    if (is_plane(k)) break;  // emitted by the outer loop
                    size_t e = run_end_of(k);
                    if (!col_state[k - p.op_begin] && dsubs.size() > first_sub) {
                        // a collapsible run ends this round trip (it is emitted by the outer loop)
                        const size_t mark_c = dcols.size(), mark_e = dents.size(), mark_s = dsupers.size();
                        const size_t mark_r = out.recipes.size(), mark_p = out.col_pats.size();
                        const bool imag_before = p.has_imag;
                        const int cr = try_collapse(k, e);
                        if (cr != 0) {
                            dcols.resize(mark_c);
                            dents.resize(mark_e);
                            dsupers.resize(mark_s);
                            out.recipes.resize(mark_r);
                            out.col_pats.resize(mark_p);
                            p.has_imag = imag_before;
                            break;
                        }
                        col_state[k - p.op_begin] = 1;
                    }
                    uint32_t coord = 0;
                    const uint32_t res = reduce(dops[k].lx, coord);
                    uint32_t cpat = coord;
                    if (res != 0) {
                        if (dim == 3) break;  // a fourth independent X-mask: next round trip
                        v[dim] = dops[k].lx;
                        rr[dim] = res;
                        T[dim] = coord ^ (1u << dim);
                        piv[dim] = pick_pivot(res);
                        cpat = 1u << dim;
                        ++dim;
                    }
                    dsubs.push_back({(uint32_t)(k - p.op_begin), (uint32_t)(e - k), cpat, dops[k].imag});
                    if (dops[k].imag) p.has_imag = true;
                    k = e;
                }
                for (int b2 = p.tp.tbits - 1; b2 >= 0 && dim < 3; --b2) {  // fill up with free single bits
                    bool is_piv = false;
                    for (int a2 = 0; a2 < dim; ++a2) is_piv = is_piv || piv[a2] == (uint32_t)b2;
                    if (is_piv) continue;
                    uint32_t coord = 0;
                    const uint32_t res = reduce(1u << b2, coord);
                    if (res == 0) continue;
                    v[dim] = 1u << b2;
                    rr[dim] = res;
                    T[dim] = coord ^ (1u << dim);
                    piv[dim] = pick_pivot(res);
                    ++dim;
                }
                uint32_t ps3[3] = {piv[0], piv[1], piv[2]};
                std::sort(ps3, ps3 + 3);
                su.e0 = ps3[0]; su.e1 = ps3[1]; su.e2 = ps3[2];
                for (uint32_t b2 = 0; b2 < 8; ++b2)
                    su.off[b2] = ((b2 & 1u) ? v[0] : 0u) ^ ((b2 & 2u) ? v[1] : 0u) ^ ((b2 & 4u) ? v[2] : 0u);
                su.sub_count = (uint32_t)(dsubs.size() - first_sub);
                // sign step of every rotation along the three orbit directions (RotOp::mq)
                for (size_t q = first_sub; q < dsubs.size(); ++q)
                    for (uint32_t w = 0; w < dsubs[q].len; ++w) {
                        DevOp& d = dops[p.op_begin + dsubs[q].begin + w];
                        d.jmask = 0;
                        for (int a2 = 0; a2 < 3; ++a2)
                            if (__builtin_popcount(v[a2] & d.lz) & 1) d.jmask |= 1u << a2;
                    }
                close_scale(su, first_sub);
                dsupers.push_back(su);
            }
            p.pass_scale = pending;
            p.sup_end = dsupers.size();
            p.sub_end = dsubs.size();
            p.col_end = dcols.size();
            p.ent_end = dents.size();
            if (p.tp.vbit && p.tp.tbits >= 3 && env_int("VQE_PEER_GATHER", 1)) {
                // dependency closure, walking the segments backwards: which amplitudes of the partner half does a
                // rank need (before the pass) to compute the final values of its own half?
                const uint32_t tsz = 1u << p.tp.tbits, hbit = tsz >> 1;
                auto closure = [&](uint32_t own_half, std::vector<uint16_t>& out) {
                    std::vector<char> R(tsz, 0);
                    for (uint32_t l = 0; l < tsz; ++l) R[l] = ((l & hbit) ? 1u : 0u) == own_half;
                    auto pull = [&](uint32_t u, uint32_t v) {
                        if (R[u] | R[v]) R[u] = R[v] = 1;
                    };
                    for (size_t si = p.sup_end; si-- > p.sup_begin;) {
                        const DevSuper& su = dsupers[si];
                        if (su.sub_count == 0xffffffffu) {
                            const DevCol& co = dcols[p.col_begin + su.sub_begin];
                            for (uint32_t e2 = 0; e2 < co.n_active; ++e2) {
                                const uint32_t pat = dents[p.ent_begin + co.ent_begin + e2].pat;
                                for (uint32_t f = 0; f < (1u << co.free_log); ++f) {
                                    uint32_t l = f;
                                    for (uint32_t d = 0; d < co.nd; ++d) l += l & co.dpos[d];
                                    l |= pat;
                                    pull(l, l ^ co.lx);
                                }
                            }
                        } else {
                            for (uint32_t sb = su.sub_count; sb-- > 0;) {
                                const DevSub& sub = dsubs[p.sub_begin + su.sub_begin + sb];
                                const uint32_t lxs = four ? su.off[sub.c & 7u] : su.off[1];
                                for (uint32_t l = 0; l < tsz; ++l)
                                    if (!(l & (1u << (31 - __builtin_clz(lxs))))) pull(l, l ^ lxs);
                            }
                        }
                    }
                    out.clear();
                    const uint32_t par = own_half ? 0u : hbit;
                    for (uint32_t gidx = 0; gidx < hbit / 4; ++gidx) {
                        bool any = false;
                        for (uint32_t e2 = 0; e2 < 4; ++e2) any = any || R[par + gidx * 4 + e2];
                        if (any) out.push_back((uint16_t)gidx);
                    }
                };
                closure(0, p.need_lo);
                closure(1, p.need_hi);
                const size_t half_groups = hbit / 4;
                p.gather = !p.need_lo.empty() && !p.need_hi.empty() && 2 * p.need_lo.size() <= half_groups &&
                           2 * p.need_hi.size() <= half_groups;
                if (getenv("VQE_DEBUG_PLAN") && atoi(getenv("VQE_DEBUG_PLAN")) > 1)
                    fprintf(stderr, "[plan] peer pass: %zu segments, need_lo %zu need_hi %zu of %zu groups -> %s\n",
                            p.sup_end - p.sup_begin, p.need_lo.size(), p.need_hi.size(), half_groups, p.gather ? "gather" : "exchange");
            }
        }
        if (p.fast) {   // first word of every collapsed run in the pass's item table (k_col_tab)
            uint32_t running = 0;
            for (size_t q = p.col_begin; q < dcols.size(); ++q) {
                dcols[q].pad = running;
                running += dcols[q].n_active << dcols[q].free_log;
            }
        }
        passes.push_back(std::move(p));
        i = j;
    }
    return VQE_OK;
}

// plan + upload + launch an ordered op list on buffer 0 of every rank in the set
static int launch_plan(RankSet& rs, const OpPlan& plan, int buf = VQE_BUF_PSI);

static int run_ops(RankSet& rs, const std::vector<HostOp>& ops) {
    int rc = check_rankset(rs);
    if (rc) return rc;
    if (ops.empty()) return VQE_OK;
    vqe_ctx* c0 = rs.r[0];
    OpPlan plan;
    rc = plan_ops(c0->n, c0->nl, c0->tile_bits, c0->low_bits, c0->threads, ops, plan);
    if (rc) return rc;
    return launch_plan(rs, plan);
}

// upload + launch a planned op list on buffer 0 of every rank in the set
// Can every pass of the plan run on the REAL LAYOUT of the state (k_tile_col<true, true>)?  Purely real collapsed-run
// passes, local, with a real-layout tensor-map form and tables that fit next to the tile.
static bool plan_runs_in_real_layout(const OpPlan& plan) {
    if (env_int("VQE_COL_KERNEL", 1) == 0 || env_int("VQE_PIPE", 0) != 0) return false;
    for (const OpPass& ps : plan.passes) {
        const bool all_col = ps.fast && !ps.has_imag && ps.sub_end == ps.sub_begin && ps.col_end > ps.col_begin && ps.pass_scale == 1.0 &&
                             (ps.sup_end - ps.sup_begin) == (ps.col_end - ps.col_begin);
        const size_t smem_r = (8ull << ps.tp.tbits) + (size_t)(ps.col_end - ps.col_begin) * (sizeof(ColLite) + 8 + 4) +
                              (size_t)(ps.ent_end - ps.ent_begin) * sizeof(DevColEntry);
        if (!all_col || ps.tp.vbit || smem_r > 110 * 1024) return false;
        if (!(tma_available(ps.tp, true, true) || tma_available(ps.tp, false, true))) return false;
    }
    return true;
}
// Launch attribute: access-policy window over `bytes` of the state at `base` (see vqe_create).  Returns 1 when the attribute
// was filled.  Only for states up to a few times the set-aside: beyond that the resident fraction is not worth the
// L2 capacity taken from everything else.
static int l2_window_attr(const vqe_ctx* c, cudaLaunchAttribute* at, const void* base, size_t bytes) {
    if (!c->l2_persist_bytes || !base || bytes == 0 || bytes > 4 * c->l2_persist_bytes) return 0;
    const size_t win = std::min(bytes, c->l2_window_max);
    memset(at, 0, sizeof *at);
    at->id = cudaLaunchAttributeAccessPolicyWindow;
    at->val.accessPolicyWindow.base_ptr = const_cast<void*>(base);
    at->val.accessPolicyWindow.num_bytes = win;
    at->val.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)c->l2_persist_bytes * env_int("VQE_L2_HIT_PCT", 100) / 100.0 / (double)win);
    at->val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at->val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    return 1;
}
// Item table of k_col_tab for every pass of a real-layout collapsed-run plan (pure host code).  One word per item of a run:
// bits 0-15 byte offset of the a-side element in the pass's (swizzled) shared-memory tile, bit 16 parity(l & lz),
// bits 17- pattern number within the run.  The index arithmetic is the one k_tile_col does per thread and run (col_prep).
static int build_coltab(const OpPlan& plan, bool want32) {
    static std::atomic<uint64_t> next_gen{1};
    const std::vector<OpPass>& passes = plan.passes;
    plan.coltab.clear();
    plan.coltab_off.assign(passes.size(), 0);
    plan.coltab_swz.assign(passes.size(), 0);
    plan.coltab16.clear();
    plan.coltab16_off.assign(passes.size(), 0);
    plan.coltab16_seg.assign(passes.size(), 0);
    plan.coltab_has32 = want32;
    for (size_t p = 0; p < passes.size(); ++p) {
        const OpPass& ps = passes[p];
        const uint32_t swz = rl_tile_swz(ps.tp);
        plan.coltab_off[p] = plan.coltab.size();
        plan.coltab_swz[p] = swz;
        // 16-bit form (k_col_stab): per pass [first segment of every run: n_cols x uint32, padded to 16 bytes][segments]: a segment
        // is 512 slots, slot = (element index << 3) | parity or 0xffff = empty, stored FACTORED as 32 lane parts + 16 group parts
        // (slot s = LANE[s & 31] ^ GROUP[s >> 5]); a pattern of a run is ceil(2^free_log / 512) segments
        plan.coltab16_off[p] = plan.coltab16.size();
        const size_t n_cols = ps.col_end - ps.col_begin;
        if (ps.tp.tbits > 13) return fail(VQE_ERR_INVALID, "item table: %d-bit tile", ps.tp.tbits);
        plan.coltab16.resize(plan.coltab16.size() + ((2 * n_cols + 7) & ~size_t(7)), 0);
        uint32_t n_seg = 0, items_seen = 0;
        for (size_t q = ps.col_begin; q < ps.col_end; ++q) {
            const DevCol& co = plan.dcols[q];
            const uint32_t items = co.n_active << co.free_log;
            if (co.pad != items_seen) return fail(VQE_ERR_INVALID, "item table of pass %zu is inconsistent with its run descriptors", p);
            items_seen += items;
            // Which free tile position each bit of the item number drives.  Any bijection is a valid enumeration of the run's
            // pairs; the one chosen makes the 16 lanes of a half-warp (item bits 0-3) hit 16 different 8-byte bank pairs: a
            // 64-bit shared-memory access is served per half-warp, and the bank pair of an element is bits 0-3 of its swizzled
            // index, so bits 0-3 go to free positions whose images under the swizzle are linearly independent in those four
            // bits (the plain ascending order conflicts 2-way whenever the X-mask holds tile position 2 or 3 -- 29 % of the
            // wavefronts of a 25-run pass were replays).  What no order can avoid: a fixed position 0, or both positions that feed
            // one bank bit fixed (vqe_debug_coltab_banks counts both kinds; tests/test_coltab_cpu.py).
            std::vector<uint32_t> order;
            {
                uint32_t fixed = 0;
                for (uint32_t d = 0; d < co.nd && d < 6; ++d) fixed |= 1u << __builtin_popcount(~co.dpos[d]);
                std::vector<uint32_t> freep;
                for (uint32_t b2 = 0; b2 < (uint32_t)ps.tp.tbits; ++b2)
                    if (!((fixed >> b2) & 1u)) freep.push_back(b2);
                if (freep.size() != co.free_log) return fail(VQE_ERR_INVALID, "run descriptor of pass %zu: %zu free positions, free_log %u", p, freep.size(), co.free_log);
                std::vector<char> used(freep.size(), 0);
                uint32_t basis[4] = {0, 0, 0, 0};   // GF(2) basis of the chosen bank images, reduced by leading bit
                if (env_int("VQE_COL_LANE_ORDER", 1) != 0)
                    for (size_t k = 0; k < freep.size() && order.size() < 4; ++k) {
                        uint32_t v = swz_idx8(1u << freep[k], swz) & 0xfu;
                        for (int b3 = 3; b3 >= 0; --b3)
                            if (((v >> b3) & 1u) && basis[b3]) v ^= basis[b3];
                        if (!v) continue;
                        basis[31 - __builtin_clz(v)] = v;
                        used[k] = 1;
                        order.push_back(freep[k]);
                    }
                for (size_t k = 0; k < freep.size(); ++k)
                    if (!used[k]) order.push_back(freep[k]);
            }
            // index deposit, swizzle and parity are linear over XOR: the word of an item is the XOR of one word per set bit of
            // its number and the word of its pattern
            auto word_of = [&](uint32_t l) -> uint32_t {   // (byte offset in the swizzled tile) | parity in bit 0
                return (swz_idx8(l, swz) << 3) | ((uint32_t)__builtin_popcount(l & co.lz) & 1u);
            };
            uint32_t wbit[16];
            for (uint32_t k = 0; k < co.free_log; ++k) wbit[k] = word_of(1u << order[k]);
            auto word_free = [&](uint32_t it_in_pat) -> uint32_t {
                uint32_t w = 0;
                for (uint32_t k = 0; k < co.free_log; ++k)
                    if ((it_in_pat >> k) & 1u) w ^= wbit[k];
                return w;
            };
            if (want32)
                for (uint32_t it = 0; it < items; ++it) {
                    const uint32_t pi = it >> co.free_log;
                    const uint32_t w = word_free(it & ((1u << co.free_log) - 1u)) ^ word_of(plan.dents[ps.ent_begin + co.ent_begin + pi].pat);
                    plan.coltab.push_back((w & 0xfff8u) | ((w & 1u) << 16) | (pi << 17));
                }
            reinterpret_cast<uint32_t*>(plan.coltab16.data() + plan.coltab16_off[p])[q - ps.col_begin] = n_seg;
            const uint32_t per_pat_items = 1u << co.free_log;
            const uint32_t per_pat_seg = std::max<uint32_t>(1u, per_pat_items / COLSEG);
            uint16_t lane[32];
            for (uint32_t ln = 0; ln < 32; ++ln) lane[ln] = ln < per_pat_items ? (uint16_t)word_free(ln) : (uint16_t)0xffffu;
            for (uint32_t pi = 0; pi < co.n_active; ++pi) {
                const uint32_t wpat = word_of(plan.dents[ps.ent_begin + co.ent_begin + pi].pat);
                for (uint32_t k = 0; k < per_pat_seg; ++k, ++n_seg) {
                    plan.coltab16.insert(plan.coltab16.end(), lane, lane + 32);
                    for (uint32_t gj = 0; gj < 16; ++gj) {
                        const uint32_t it_in_pat = k * COLSEG + 32u * gj;
                        plan.coltab16.push_back(it_in_pat < per_pat_items ? (uint16_t)(word_free(it_in_pat) ^ wpat) : (uint16_t)0xffffu);
                    }
                }
            }
        }
        plan.coltab16_seg[p] = n_seg;
    }
    plan.gen = next_gen++;
    return VQE_OK;
}
// CTAs of a gather-form pass launch that fetch the NEXT chunk's partner amplitudes (the launch keeps its persistent grid:
// they are carved out of it and grown back when the grid would otherwise be too small to hold them)
static uint32_t gather_ctas(const vqe_ctx* c, bool wanted, int& grid) {
    if (!wanted) return 0;
    const int want = std::max(1, env_int("VQE_GATHER_CTAS", c->sm_count));
    const int n = std::min(want, std::max(1, grid / 2));
    if (grid - n < 1) grid = n + 1;
    return (uint32_t)n;
}
static int launch_plan(RankSet& rs, const OpPlan& plan, int buf) {
    int rc = VQE_OK;
    for (vqe_ctx* c : rs.r) {
        CK(cudaSetDevice(c->device));
        rc = ensure_buf(c, buf);
        if (rc) return rc;
    }
    const std::vector<OpPass>& passes = plan.passes;
    // upload: [ops][mats][runs][scat tables]
    size_t off_ops = 0, off_mats = plan.dops.size() * sizeof(DevOp);
    size_t off_runs = (off_mats + plan.mats.size() * sizeof(double) + 15) & ~size_t(15);
    size_t off_subs = off_runs + plan.dsupers.size() * sizeof(DevSuper);
    size_t off_cols = off_subs + plan.dsubs.size() * sizeof(DevSub);
    size_t off_ents = off_cols + plan.dcols.size() * sizeof(DevCol);
    size_t off_scat = off_ents + plan.dents.size() * sizeof(DevColEntry);
    off_scat = (off_scat + 15) & ~size_t(15);
    size_t total = off_scat;
    std::vector<size_t> scat_off(passes.size()), need_off(passes.size(), 0);
    for (size_t p = 0; p < passes.size(); ++p) {
        scat_off[p] = total;
        total += passes[p].tp.scat.size() * sizeof(uint64_t);
    }
    for (size_t p = 0; p < passes.size(); ++p)
        if (passes[p].gather) {  // [need_lo][need_hi], 16-byte aligned
            need_off[p] = total;
            total += ((passes[p].need_lo.size() + passes[p].need_hi.size()) * sizeof(uint16_t) + 15) & ~size_t(15);
        }
    for (vqe_ctx* c : rs.r) {
        CK(cudaSetDevice(c->device));
        rc = ensure_stage(c, total);
        if (rc) return rc;
        // the staging buffer may still be read by an earlier async copy
        CK(cudaStreamSynchronize(c->stream));
        memcpy(c->h_stage + off_ops, plan.dops.data(), plan.dops.size() * sizeof(DevOp));
        if (!plan.mats.empty()) memcpy(c->h_stage + off_mats, plan.mats.data(), plan.mats.size() * sizeof(double));
        if (!plan.dsupers.empty()) memcpy(c->h_stage + off_runs, plan.dsupers.data(), plan.dsupers.size() * sizeof(DevSuper));
        if (!plan.dsubs.empty()) memcpy(c->h_stage + off_subs, plan.dsubs.data(), plan.dsubs.size() * sizeof(DevSub));
        if (!plan.dcols.empty()) memcpy(c->h_stage + off_cols, plan.dcols.data(), plan.dcols.size() * sizeof(DevCol));
        if (!plan.dents.empty()) memcpy(c->h_stage + off_ents, plan.dents.data(), plan.dents.size() * sizeof(DevColEntry));
        for (size_t p = 0; p < passes.size(); ++p)
            memcpy(c->h_stage + scat_off[p], passes[p].tp.scat.data(), passes[p].tp.scat.size() * sizeof(uint64_t));
        for (size_t p = 0; p < passes.size(); ++p)
            if (passes[p].gather) {
                memcpy(c->h_stage + need_off[p], passes[p].need_lo.data(), passes[p].need_lo.size() * sizeof(uint16_t));
                memcpy(c->h_stage + need_off[p] + passes[p].need_lo.size() * sizeof(uint16_t), passes[p].need_hi.data(),
                       passes[p].need_hi.size() * sizeof(uint16_t));
            }
        c->h2d_bytes += total;
        CK(cudaMemcpyAsync(c->d_stage, c->h_stage, total, cudaMemcpyHostToDevice, c->stream));
    }
    // Purely real state (set_basis_state, then only +-1-phase fast rotations -- the UCC case): the passes skip the
    // imaginary halves.  All ranks of a sharded state see the same op list, so the flag stays consistent.
    std::vector<char> real_pass(passes.size(), 0);
    {
        bool real = buf == VQE_BUF_PSI;  // the flag is only tracked for the state buffer
        for (vqe_ctx* c : rs.r) real = real && c->psi_real;
        for (size_t p = 0; p < passes.size(); ++p) {
            real = real && passes[p].fast && !passes[p].has_imag;
            real_pass[p] = real ? 1 : 0;
        }
        if (buf == VQE_BUF_PSI)
            for (vqe_ctx* c : rs.r) c->psi_real = real;
    }
    // Real layout of the state buffer (unsharded contexts, see vqe_ctx::real_layout): kept when EVERY pass is a purely real
    // collapsed-run pass with a real-layout tensor-map form (every UCCSD / QUCCSD program); otherwise the buffer is expanded
    // to interleaved complex first and the passes run as usual.
    bool rl_plan = buf == VQE_BUF_PSI;
    for (vqe_ctx* c : rs.r) rl_plan = rl_plan && c->real_layout;
    for (size_t p = 0; p < passes.size() && rl_plan; ++p) rl_plan = real_pass[p] != 0;
    rl_plan = rl_plan && plan_runs_in_real_layout(plan);
    if (!rl_plan && plan.tile_bits > 12) return fail(VQE_ERR_INVALID, "a 13-bit tile plan needs the real layout of the state");
    if (!rl_plan && buf == VQE_BUF_PSI)
        for (vqe_ctx* c : rs.r) {
            CK(cudaSetDevice(c->device));
            rc = ensure_complex(c, buf);
            if (rc) return rc;
        }
    // item table of k_col_tab: built once per plan, uploaded once per context and plan
    // VQE_COL_TAB: 0 = k_tile_col (index arithmetic per thread and run), 1 = k_col_tab (32-bit items in global memory),
    // 2 = k_col_stab (16-bit items next to the tile in shared memory; passes whose table does not fit keep k_tile_col)
    const int tab_mode = rl_plan ? std::max(0, std::min(2, env_int("VQE_COL_TAB", 2))) : 0;
    const bool use_tab = tab_mode != 0;
    if (use_tab) {
        if (plan.gen == 0 || (tab_mode == 1 && !plan.coltab_has32)) {
            rc = build_coltab(plan, tab_mode == 1);
            if (rc) return rc;
        }
        const void* tab_src = tab_mode == 2 ? (const void*)plan.coltab16.data() : (const void*)plan.coltab.data();
        const size_t tab_bytes = tab_mode == 2 ? plan.coltab16.size() * sizeof(uint16_t) : plan.coltab.size() * sizeof(uint32_t);
        for (vqe_ctx* c : rs.r) {
            if (c->coltab_gen == plan.gen && c->coltab_mode == tab_mode) continue;
            CK(cudaSetDevice(c->device));
            if (c->coltab_cap < std::max<size_t>(16, tab_bytes)) {
                if (c->d_coltab) cudaFree(c->d_coltab);
                c->d_coltab = nullptr;
                c->coltab_cap = 0;
                if (cudaMalloc((void**)&c->d_coltab, std::max<size_t>(16, tab_bytes)) != cudaSuccess)
                    return fail(VQE_ERR_NOMEM, "item table of the rotation plan (%.1f MB)", tab_bytes / 1e6);
                c->coltab_cap = std::max<size_t>(16, tab_bytes);
            }
            if (tab_bytes) CK(cudaMemcpyAsync(c->d_coltab, tab_src, tab_bytes, cudaMemcpyHostToDevice, c->stream));
            c->h2d_bytes += tab_bytes;
            c->coltab_gen = plan.gen;
            c->coltab_mode = tab_mode;
        }
    }
    // one launch of the pass kernel on one rank (gg.n_need != 0: gather form over the tiles of the current chunk)
    auto launch_pass = [&](vqe_ctx* c, size_t p, const TileGeom& g, const Shards& sh, const GatherGeom& gg,
                           const GatherGeom* gnext = nullptr, uint64_t next_first = 0, uint64_t next_tiles = 0,
                           double2* next_stage = nullptr) -> int {
        const OpPass& ps = passes[p];
        const GatherGeom gn = gnext ? *gnext : GatherGeom{nullptr, nullptr, 0u, 0u};
        size_t smem = tile_smem(ps.tp.tbits, 1, false) +
                      (ps.fast ? (ps.op_end - ps.op_begin) * sizeof(RotOp) + (ps.sup_end - ps.sup_begin) * sizeof(DevSuper) +
                                     (ps.sub_end - ps.sub_begin) * sizeof(DevSub) + (ps.col_end - ps.col_begin) * (sizeof(DevCol) + 4) +
                                     (ps.ent_end - ps.ent_begin) * sizeof(DevColEntry)
                               : (ps.op_end - ps.op_begin) * sizeof(FastOp));
        int threads = (int)std::min<uint64_t>(c->threads, std::max<uint64_t>(32, (1ull << ps.tp.tbits) / 2));
        ProfScope prof(c, ps.tp.vbit ? 4 : 0);
        const bool all_col = ps.fast && ps.sub_end == ps.sub_begin && ps.col_end > ps.col_begin && gg.n_need == 0 &&
                             ps.pass_scale == 1.0 && (ps.sup_end - ps.sup_begin) == (ps.col_end - ps.col_begin) &&
                             env_int("VQE_COL_KERNEL", 1) != 0;
        if (all_col) {
            // every segment is a collapsed run / tabulated plane rotation: the lean kernel (256 threads, 3 CTAs per SM)
            const int n_cols = (int)(ps.col_end - ps.col_begin), n_ents = (int)(ps.ent_end - ps.ent_begin);
            const size_t smem_c = tile_smem(ps.tp.tbits, 1, false) + (size_t)n_cols * (sizeof(ColLite) + 8 + 4) + (size_t)n_ents * sizeof(DevColEntry);
            const int thr = (int)std::min<uint64_t>(256, std::max<uint64_t>(32, (1ull << ps.tp.tbits) / 2));
            const int grid = tile_grid(c, g.n_tiles, smem_c <= 74 * 1024 ? 3 : 2);
            const size_t smem_p = (size_t)NSLOT * tile_smem(ps.tp.tbits, 1, false) + (size_t)n_cols * (sizeof(ColLite) + 8 + 8) +
                                  (size_t)n_ents * sizeof(DevColEntry);
            TileGeom gt = g;
            CUtensorMap tmap;
            if (rl_plan) {
                // real layout: 8-byte tile elements, tensor-map staging only, four CTAs per SM
                make_tmap(ps.tp, sh.p0, gt, &tmap, -1, true);
                if (!gt.tma) return fail(VQE_ERR_CUDA, "real-layout tensor map of a rotation pass could not be encoded");
                const size_t smem_r = (8ull << ps.tp.tbits) + (size_t)n_cols * (sizeof(ColLite) + 8 + 4) + (size_t)n_ents * sizeof(DevColEntry);
                const int grid_r = tile_grid(c, g.n_tiles, smem_r <= 54 * 1024 ? 4 : (smem_r <= 74 * 1024 ? 3 : 2));
                const size_t n_seg = tab_mode == 2 ? plan.coltab16_seg[p] : 0;
                const size_t smem_s = (8ull << ps.tp.tbits) + n_seg * (sizeof(ColSub) + 8) + 16 + n_seg * COLFACT * sizeof(uint16_t);
                if (tab_mode == 2 && smem_s <= 110 * 1024) {
                    if (gt.swz != plan.coltab_swz[p]) return fail(VQE_ERR_CUDA, "tile swizzle of pass %zu differs from its item table", p);
                    const int ctas_s = smem_s <= 74 * 1024 ? 3 : 2;
                    const int grid_s = tile_grid(c, g.n_tiles, ctas_s);
                    // Consecutive passes walk the tiles in alternating directions: a pass reads and rewrites every line of the
                    // state once (the footprint never grows), so what the previous pass touched last is what the L2 still
                    // holds when the state is about its size (24 qubits in the real layout: 134 MB against 126 MB), and the
                    // next pass starts there.  H2O: rotation passes 65 -> 58 us on average.  (Ordering the tiles by the
                    // recency of their lines under the previous pass's own bit significance measured 2 us slower.)
                    if (buf == VQE_BUF_PSI && gt.tile_stride == 1 && gt.tile_first == 0 && env_int("VQE_ALT_ORDER", 1) != 0) {
                        c->walk_desc = !c->walk_desc;
                        if (c->walk_desc) {
                            gt.rev = 1;
                            gt.tile_first = gt.n_tiles - 1;
                        }
                    }
                    int thr_s = env_int(ctas_s == 3 ? "VQE_STAB_THREADS3" : "VQE_STAB_THREADS2", 512);
                    if (thr_s != 512) thr_s = 256;
                    cudaLaunchConfig_t cfg;
                    memset(&cfg, 0, sizeof cfg);
                    cfg.gridDim = dim3((unsigned)grid_s);
                    cfg.blockDim = dim3((unsigned)thr_s);
                    cfg.dynamicSmemBytes = smem_s;
                    cfg.stream = c->stream;
                    cudaLaunchAttribute at[2];
                    unsigned n_at = 0;
                    if (env_int("VQE_PDL", 1) != 0) {
                        at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                        at[n_at].val.programmaticStreamSerializationAllowed = 1;
                        ++n_at;
                    }
                    n_at += (unsigned)l2_window_attr(c, &at[n_at], sh.p0, (size_t)c->n_amp * sizeof(double));
                    cfg.attrs = at;
                    cfg.numAttrs = n_at;
                    auto kern = thr_s == 512 ? (ctas_s == 3 ? k_col_stab<512, 3> : k_col_stab<512, 2>) : k_col_stab<256, 3>;
                    CK(cudaLaunchKernelEx(&cfg, kern, tmap, gt, (const DevCol*)(c->d_stage + off_cols) + ps.col_begin, n_cols,
                                          (const DevColEntry*)(c->d_stage + off_ents) + ps.ent_begin, (int)n_seg,
                                          (const uint16_t*)c->d_coltab + plan.coltab16_off[p], env_int("VQE_DEBUG_SKELETON", 0), c->d_err));
                    c->launches++;
                    return VQE_OK;
                }
                if (tab_mode == 1) {
                    if (gt.swz != plan.coltab_swz[p]) return fail(VQE_ERR_CUDA, "tile swizzle of pass %zu differs from its item table", p);
                    const size_t smem_t = (8ull << ps.tp.tbits) + (size_t)n_cols * (sizeof(ColTabRun) + 8 + 4) + (size_t)n_ents * sizeof(double2);
                    const int grid_t = tile_grid(c, g.n_tiles, smem_t <= 54 * 1024 ? 4 : (smem_t <= 74 * 1024 ? 3 : 2));
                    cudaLaunchConfig_t cfg;
                    memset(&cfg, 0, sizeof cfg);
                    cfg.gridDim = dim3((unsigned)grid_t);
                    cfg.blockDim = dim3(256);
                    cfg.dynamicSmemBytes = smem_t;
                    cfg.stream = c->stream;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                    at[0].val.programmaticStreamSerializationAllowed = 1;
                    cfg.attrs = at;
                    cfg.numAttrs = env_int("VQE_PDL", 1) != 0 ? 1 : 0;
                    CK(cudaLaunchKernelEx(&cfg, k_col_tab, tmap, gt, (const DevCol*)(c->d_stage + off_cols) + ps.col_begin, n_cols,
                                          (const DevColEntry*)(c->d_stage + off_ents) + ps.ent_begin, n_ents,
                                          (const uint32_t*)c->d_coltab + plan.coltab_off[p], env_int("VQE_DEBUG_SKELETON", 0), c->d_err));
                    c->launches++;
                    return VQE_OK;
                }
                k_tile_col<true, true><<<grid_r, thr, smem_r, c->stream>>>(tmap, sh, gt, (const DevCol*)(c->d_stage + off_cols) + ps.col_begin, n_cols,
                                                                          (const DevColEntry*)(c->d_stage + off_ents) + ps.ent_begin, n_ents, c->d_err);
                c->launches++;
                CK(cudaGetLastError());
                return VQE_OK;
            }
            make_tmap(ps.tp, ps.tp.vbit ? nullptr : sh.p0, gt, &tmap);
            if ((gt.bulk || gt.tma) && smem_p <= 226 * 1024 && env_int("VQE_PIPE", 0) != 0) {
                // tile ring: one persistent CTA per SM, loads two tiles ahead, stores draining behind
                uint32_t max_items = 0;
                for (size_t q = ps.col_begin; q < ps.col_end; ++q)
                    max_items = std::max(max_items, plan.dcols[q].n_active << plan.dcols[q].free_log);
                const int thr_p = (int)std::min<uint64_t>(max_items <= 256 ? 256 : 512, std::max<uint64_t>(32, (1ull << ps.tp.tbits) / 2));
                const int grid_p = tile_grid(c, g.n_tiles, 1);
                if (env_int("VQE_DEBUG_SKELETON", 0))  // measurement aid: loads and stores only, no arithmetic (results are wrong)
                    k_col_pipe<true><<<grid_p, thr_p, smem_p, c->stream>>>(tmap, sh, gt, (const DevCol*)(c->d_stage + off_cols) + ps.col_begin, 0,
                                                                          (const DevColEntry*)(c->d_stage + off_ents) + ps.ent_begin, 0, c->d_err);
                else if (real_pass[p])
                    k_col_pipe<true><<<grid_p, thr_p, smem_p, c->stream>>>(tmap, sh, gt, (const DevCol*)(c->d_stage + off_cols) + ps.col_begin, n_cols,
                                                                          (const DevColEntry*)(c->d_stage + off_ents) + ps.ent_begin, n_ents, c->d_err);
                else
                    k_col_pipe<false><<<grid_p, thr_p, smem_p, c->stream>>>(tmap, sh, gt, (const DevCol*)(c->d_stage + off_cols) + ps.col_begin, n_cols,
                                                                           (const DevColEntry*)(c->d_stage + off_ents) + ps.ent_begin, n_ents, c->d_err);
            } else if (real_pass[p])
                k_tile_col<true, false><<<grid, thr, smem_c, c->stream>>>(tmap, sh, gt, (const DevCol*)(c->d_stage + off_cols) + ps.col_begin, n_cols,
                                                                  (const DevColEntry*)(c->d_stage + off_ents) + ps.ent_begin, n_ents, c->d_err);
            else
                k_tile_col<false, false><<<grid, thr, smem_c, c->stream>>>(tmap, sh, gt, (const DevCol*)(c->d_stage + off_cols) + ps.col_begin, n_cols,
                                                                   (const DevColEntry*)(c->d_stage + off_ents) + ps.ent_begin, n_ents, c->d_err);
        } else if (ps.fast && real_pass[p])
        {
            int grid_r = tile_grid(c, g.n_tiles, (smem <= 74 * 1024 && env_int("VQE_REAL_CTAS", 3) == 3) ? 3 : 0);
            const uint32_t n_gctas = gather_ctas(c, gnext != nullptr, grid_r);
            k_tile_rot<true><<<grid_r, threads, smem, c->stream>>>(
                sh, g, gg, (const DevOp*)(c->d_stage + off_ops) + ps.op_begin, (int)(ps.op_end - ps.op_begin),
                (const DevSuper*)(c->d_stage + off_runs) + ps.sup_begin, (int)(ps.sup_end - ps.sup_begin),
                (const DevSub*)(c->d_stage + off_subs) + ps.sub_begin, (int)(ps.sub_end - ps.sub_begin),
                (const DevCol*)(c->d_stage + off_cols) + ps.col_begin, (int)(ps.col_end - ps.col_begin),
                (const DevColEntry*)(c->d_stage + off_ents) + ps.ent_begin, (int)(ps.ent_end - ps.ent_begin), ps.pass_scale, c->d_err,
                gn, next_first, next_tiles, next_stage, n_gctas);
        }
        else if (ps.fast)
        {
            int grid_r = tile_grid(c, g.n_tiles);
            const uint32_t n_gctas = gather_ctas(c, gnext != nullptr, grid_r);
            k_tile_rot<false><<<grid_r, threads, smem, c->stream>>>(
                sh, g, gg, (const DevOp*)(c->d_stage + off_ops) + ps.op_begin, (int)(ps.op_end - ps.op_begin),
                (const DevSuper*)(c->d_stage + off_runs) + ps.sup_begin, (int)(ps.sup_end - ps.sup_begin),
                (const DevSub*)(c->d_stage + off_subs) + ps.sub_begin, (int)(ps.sub_end - ps.sub_begin),
                (const DevCol*)(c->d_stage + off_cols) + ps.col_begin, (int)(ps.col_end - ps.col_begin),
                (const DevColEntry*)(c->d_stage + off_ents) + ps.ent_begin, (int)(ps.ent_end - ps.ent_begin), ps.pass_scale, c->d_err,
                gn, next_first, next_tiles, next_stage, n_gctas);
        }
        else
            k_tile_ops<<<tile_grid(c, g.n_tiles), threads, smem, c->stream>>>(
                sh, g, (const DevOp*)(c->d_stage + off_ops) + ps.op_begin, (int)(ps.op_end - ps.op_begin),
                (const double*)(c->d_stage + off_mats));
        c->launches++;
        CK(cudaGetLastError());
        return VQE_OK;
    };
    const GatherGeom no_gather = {nullptr, nullptr, 0u, 0u};
    bool fenced = false;  // a cross-rank barrier separates the previous pass from the next one
    for (size_t p = 0; p < passes.size(); ++p) {
        const OpPass& ps = passes[p];
        if (ps.tp.vbit && !fenced) {
            rc = rank_barrier(rs);  // the partner's earlier writes to its shard are complete
            if (rc) return rc;
        }
        if (ps.tp.vbit && ps.gather && ps.fast) {
            // gather form: [gather the needed partner amplitudes of a chunk of tiles | barrier | run the pass on the
            // chunk, own half written back locally]; nothing is written remotely, so no barrier is needed afterwards
            const size_t n_need_max = std::max(ps.need_lo.size(), ps.need_hi.size());
            const uint64_t per_tile = (uint64_t)n_need_max * 4;                      // amplitudes per tile
            // two staging buffers (chunk k is consumed while chunk k + 1 is fetched), together at most 1/8 of the shard
            const uint64_t cap_amp = std::max<uint64_t>(per_tile, std::min<uint64_t>(
                (uint64_t)env_int("VQE_GATHER_STAGE_MB", 8192) * (1ull << 20) / sizeof(double2), (rs.r[0]->n_amp / 16) + per_tile));
            const uint64_t chunk_tiles = std::max<uint64_t>(1, std::min<uint64_t>(ps.tp.n_tiles, cap_amp / per_tile));
            const bool overlap = env_int("VQE_GATHER_OVERLAP", 0) != 0 && chunk_tiles < ps.tp.n_tiles;  // measured slower on 2 x B200 (profiles/r2_summary.md): off
            for (vqe_ctx* c : rs.r) {
                CK(cudaSetDevice(c->device));
                for (int sb = 0; sb < (overlap ? 2 : 1); ++sb) {
                    if (c->gstage_cap[sb] >= chunk_tiles * per_tile) continue;
                    if (c->gstage[sb]) cudaFree(c->gstage[sb]);
                    c->gstage[sb] = nullptr;
                    c->gstage_cap[sb] = 0;
                    cudaError_t e = cudaMalloc((void**)&c->gstage[sb], chunk_tiles * per_tile * sizeof(double2));
                    if (e != cudaSuccess)
                        return fail(VQE_ERR_NOMEM, "staging buffer of a gather-form peer pass (%.1f GB) failed: %s; set VQE_PEER_GATHER=0",
                                    chunk_tiles * per_tile * 16.0 / 1e9, cudaGetErrorString(e));
                    c->gstage_cap[sb] = chunk_tiles * per_tile;
                }
            }
            // launch geometry of chunk ci on rank k (its staged amplitudes live in buffer ci & 1 when chunks overlap)
            struct ChunkGeom {
                TileGeom g;
                Shards sh;
                GatherGeom gg;
                uint64_t bytes;
            };
            auto chunk_geom = [&](size_t k, uint64_t ci, ChunkGeom& cg) -> int {
                vqe_ctx* c = rs.r[k];
                const uint64_t t0 = ci * chunk_tiles;
                const uint64_t nt = std::min<uint64_t>(chunk_tiles, ps.tp.n_tiles - t0);
                int rc2 = make_geom(c, ps.tp, (const uint64_t*)(c->d_stage + scat_off[p]), buf, cg.g, cg.sh);
                if (rc2) return rc2;
                cg.g.n_tiles = nt;      // every rank walks ALL its tiles, chunk by chunk
                cg.g.tile_first = t0;
                cg.g.tile_stride = 1;
                const int partner = c->rank ^ (int)ps.tp.gpat;
                const bool is_hi = c->rank > partner;
                const std::vector<uint16_t>& need = is_hi ? ps.need_hi : ps.need_lo;
                cg.gg.stage = c->gstage[overlap ? (ci & 1) : 0];
                cg.gg.need = (const uint16_t*)(c->d_stage + need_off[p]) + (is_hi ? ps.need_lo.size() : 0);
                cg.gg.n_need = (uint32_t)need.size();
                cg.gg.own_half = is_hi ? 1u : 0u;
                cg.bytes = (uint64_t)need.size() * 4 * nt * sizeof(double2);
                return VQE_OK;
            };
            auto gather_alone = [&](uint64_t ci) -> int {
                for (size_t k = 0; k < rs.r.size(); ++k) {
                    vqe_ctx* c = rs.r[k];
                    CK(cudaSetDevice(c->device));
                    ChunkGeom cg;
                    int rc2 = chunk_geom(k, ci, cg);
                    if (rc2) return rc2;
                    const int blocks = (int)std::min<uint64_t>((cg.g.n_tiles + 7) / 8, (uint64_t)c->sm_count * 8);
                    ProfScope prof(c, 4);
                    c->gather_bytes += cg.bytes;
                    k_gather_need<<<std::max(1, blocks), 256, 0, c->stream>>>(cg.sh, cg.g, cg.gg, const_cast<double2*>(cg.gg.stage));
                    c->launches++;
                    CK(cudaGetLastError());
                }
                return VQE_OK;
            };
            const uint64_t n_chunks = (ps.tp.n_tiles + chunk_tiles - 1) / chunk_tiles;
            for (uint64_t ci = 0; ci < n_chunks; ++ci) {
                if (ci == 0 || !overlap) {
                    rc = gather_alone(ci);
                    if (rc) return rc;
                    rc = rank_barrier(rs);  // every rank has read what it needs before anyone overwrites its tiles
                    if (rc) return rc;
                }
                const bool prefetch = overlap && ci + 1 < n_chunks;
                for (size_t k = 0; k < rs.r.size(); ++k) {
                    vqe_ctx* c = rs.r[k];
                    CK(cudaSetDevice(c->device));
                    ChunkGeom cg, nx;
                    rc = chunk_geom(k, ci, cg);
                    if (rc) return rc;
                    if (prefetch) {
                        rc = chunk_geom(k, ci + 1, nx);
                        if (rc) return rc;
                        c->gather_bytes += nx.bytes;
                        rc = launch_pass(c, p, cg.g, cg.sh, cg.gg, &nx.gg, nx.g.tile_first, nx.g.n_tiles, const_cast<double2*>(nx.gg.stage));
                    } else {
                        rc = launch_pass(c, p, cg.g, cg.sh, cg.gg);
                    }
                    if (rc) return rc;
                }
                if (prefetch) {
                    // the partners' gather CTAs have read chunk ci + 1 of my shard before my next launch overwrites it
                    rc = rank_barrier(rs);
                    if (rc) return rc;
                }
            }
            fenced = false;
            continue;
        }
        for (vqe_ctx* c : rs.r) {
            CK(cudaSetDevice(c->device));
            TileGeom g;
            Shards sh;
            rc = make_geom(c, ps.tp, (const uint64_t*)(c->d_stage + scat_off[p]), buf, g, sh);
            if (rc) return rc;
            if (g.n_tiles == 0) continue;
            rc = launch_pass(c, p, g, sh, no_gather);
            if (rc) return rc;
        }
        fenced = false;
        if (ps.tp.vbit) {
            rc = rank_barrier(rs);  // my writes into the partner's shard are complete before anyone goes on
            if (rc) return rc;
            fenced = true;
        }
    }
    return VQE_OK;
}
static int run_ops(vqe_ctx* c, const std::vector<HostOp>& ops) {
    RankSet rs;
    rs.r.push_back(c);
    return run_ops(rs, ops);
}

// ---- plan cache ------------------------------------------------------------------------------------
// A VQE optimisation applies the SAME rotation program thousands of times with different angles.  The pass plan
// (tile bit sets, segments, collapsed runs and their pattern sign masks, orbit bases, gather sets) depends only on the
// masks and on which rotations are dropped (angle 0) or need the general path (|cos| < 0.3), so it is kept and only
// the angle-dependent numbers are recomputed: the signed tangents, the (cos, sin) of every collapsed pattern and the
// pending cosine products.  If a pattern whose angle vanished before does not vanish now (or vice versa) the plan is
// rebuilt.
struct PlanCache {
    std::vector<uint64_t> x, z;
    std::vector<int32_t> ny;
    std::vector<uint8_t> cls;   // per rotation: 0 = angle 0 (dropped), 1 = fast
    OpPlan plan;
    bool valid = false;
};
static void free_plan_cache(PlanCache* pc) { delete pc; }

// angles of the kept rotations (cos, sin, angle) -> angle-dependent fields of a cached plan; false: structure changed
static bool refresh_plan(OpPlan& plan, const std::vector<double>& ang, const std::vector<double>& cs, const std::vector<double>& sn) {
    if (plan.dops.size() != ang.size()) return false;
    for (size_t k = 0; k < plan.dops.size(); ++k) {
        DevOp& d = plan.dops[k];
        double tn = sn[k] / cs[k];
        if (d.k4 >> 1) tn = -tn;
        d.s = tn;
        plan.rot_cos[k] = cs[k];
    }
    for (const OpPlan::ColRecipe& rcp : plan.recipes) {
        double scale = 0.0;
        for (uint32_t q = 0; q < rcp.n_str; ++q) scale += fabs(ang[rcp.first_op + q]);
        for (uint32_t pi = 0; pi < rcp.n_pat; ++pi) {
            const OpPlan::ColPat& cp = plan.col_pats[rcp.pat_begin + pi];
            double F = 0.0;
            for (uint32_t q = 0; q < rcp.n_str; ++q) F += ((cp.signmask >> q) & 1u) ? -ang[rcp.first_op + q] : ang[rcp.first_op + q];
            const bool active = fabs(F) > 1e-15 * scale;
            if (active != (cp.ent >= 0)) return false;
            if (active) {
                plan.dents[cp.ent].c = cos(F);
                plan.dents[cp.ent].s = sin(F);
            }
        }
    }
    for (OpPass& p : plan.passes) {
        double pending = 1.0;
        for (size_t si = p.sup_begin; si < p.sup_end; ++si) {
            DevSuper& su = plan.dsupers[si];
            if (su.sub_count == 0xffffffffu) continue;
            for (uint32_t sb = 0; sb < su.sub_count; ++sb) {
                const DevSub& sub = plan.dsubs[p.sub_begin + su.sub_begin + sb];
                for (uint32_t w = 0; w < sub.len; ++w) pending *= plan.rot_cos[p.op_begin + sub.begin + w];
            }
            su.cscale = 1.0;
            if (fabs(pending) < 1e-30) {
                su.cscale = pending;
                pending = 1.0;
            }
        }
        p.pass_scale = pending;
    }
    return true;
}

static int rotations_core(RankSet& rs, int n_rot, const uint64_t* xmask, const uint64_t* zmask, const int32_t* ny,
                          const double* angle, int buf = VQE_BUF_PSI) {
    int rc0 = check_rankset(rs);
    if (rc0) return rc0;
    if (buf < 0 || buf > 3) return fail(VQE_ERR_INVALID, "bad buffer id %d", buf);
    vqe_ctx* c = rs.r[0];
    if (n_rot < 0 || (n_rot > 0 && (!xmask || !zmask || !ny || !angle))) return fail(VQE_ERR_INVALID, "null array");
    const uint64_t full = (1ull << c->n) - 1ull;
    // angle-dependent numbers of the kept rotations, and the class of every rotation
    std::vector<double> ang, cs, sn;
    std::vector<uint8_t> cls(n_rot);
    ang.reserve(n_rot);
    cs.reserve(n_rot);
    sn.reserve(n_rot);
    bool all_fast = true;
    for (int k = 0; k < n_rot; ++k) {
        if (angle[k] == 0.0) {  // exact identity
            cls[k] = 0;
            continue;
        }
        const double cv = cos(angle[k]);
        ang.push_back(angle[k]);
        cs.push_back(cv);
        sn.push_back(sin(angle[k]));
        cls[k] = (xmask[k] != 0 && fabs(cv) >= 0.3) ? 1 : 2;
        if (cls[k] == 2) all_fast = false;
    }
    PlanCache* pc = c->plan_cache;
    const bool use_cache = env_int("VQE_PLAN_CACHE", 1) != 0;
    // REAL LAYOUT: a tile of 2^13 doubles fills the same 64 KiB as 2^12 complex amplitudes, so a pass may use one more tile
    // bit (fewer sweeps over the state).  Only if the whole plan then runs in the real layout; else the usual 12-bit plan.
    bool rl_now = buf == VQE_BUF_PSI && all_fast && env_int("VQE_RL_TILE_BITS", 13) > 12 && c->nl >= 13;
    for (vqe_ctx* r : rs.r) rl_now = rl_now && r->real_layout;
    auto make_rot_plan = [&](const std::vector<HostOp>& hops, OpPlan& plan) -> int {
        if (rl_now) {
            plan = OpPlan();
            int rcp = plan_ops(c->n, c->nl, 13, c->low_bits, 1024, hops, plan);
            if (rcp == VQE_OK && plan_runs_in_real_layout(plan)) {
                plan.tile_bits = 13;
                return VQE_OK;
            }
        }
        plan = OpPlan();
        return plan_ops(c->n, c->nl, c->tile_bits, c->low_bits, c->threads, hops, plan);
    };
    if (pc && pc->valid && pc->plan.tile_bits > 12 && !rl_now) pc->valid = false;  // planned for a layout the state has left
    if (use_cache && all_fast && pc && pc->valid && (int)pc->cls.size() == n_rot &&
        memcmp(pc->cls.data(), cls.data(), n_rot) == 0 && memcmp(pc->x.data(), xmask, n_rot * sizeof(uint64_t)) == 0 &&
        memcmp(pc->z.data(), zmask, n_rot * sizeof(uint64_t)) == 0 && memcmp(pc->ny.data(), ny, n_rot * sizeof(int32_t)) == 0) {
        if (refresh_plan(pc->plan, ang, cs, sn)) return launch_plan(rs, pc->plan, buf);
        pc->valid = false;  // a pattern's angle (stopped) vanishing: rebuild below
    }
    std::vector<HostOp> ops;
    ops.reserve(ang.size());
    size_t j = 0;
    for (int k = 0; k < n_rot; ++k) {
        if ((xmask[k] | zmask[k]) & ~full) return fail(VQE_ERR_INVALID, "rotation %d: mask has bits >= n_qubits", k);
        if (popc64(xmask[k] & zmask[k]) != ny[k]) return fail(VQE_ERR_INVALID, "rotation %d: ny != popcount(x&z)", k);
        if (cls[k] == 0) continue;
        HostOp h = HostOp();
        h.kind = OP_ROT;
        h.x = xmask[k];
        h.z = zmask[k];
        h.ny = ny[k];
        h.c = cs[j];
        h.s = sn[j];
        h.ang = ang[j];
        ++j;
        ops.push_back(h);
    }
    if (ops.empty()) return VQE_OK;
    if (!(use_cache && all_fast)) {
        OpPlan plan;
        int rcp = make_rot_plan(ops, plan);
        if (rcp) return rcp;
        return launch_plan(rs, plan, buf);
    }
    if (!pc) pc = c->plan_cache = new PlanCache();
    pc->valid = false;
    int rc = make_rot_plan(ops, pc->plan);
    if (rc) return rc;
    bool fast_only = pc->plan.cacheable && pc->plan.rot_cos.size() == pc->plan.dops.size();
    for (const OpPass& p : pc->plan.passes) fast_only = fast_only && p.fast;
    if (fast_only) {
        pc->x.assign(xmask, xmask + n_rot);
        pc->z.assign(zmask, zmask + n_rot);
        pc->ny.assign(ny, ny + n_rot);
        pc->cls = cls;
        pc->valid = true;
    }
    return launch_plan(rs, pc->plan, buf);
}
// Rotations on a SHARDED state with qubit relabelling.  A rotation whose X-mask touches a global slot would run as a peer
// pass -- and so would every later rotation that flips the same qubit.  Instead the logical bit is moved into a local
// slot ONCE (swap_global_local: half a shard over NVLink, about the cost of one peer pass) and the local slot that gives
// way is the one whose logical bit is flipped furthest in the future (Belady's rule over the rest of the program; slots
// below VQE_RELABEL_FLOOR = 10 keep their bits so that swaps move contiguous runs of at least 8 KiB).  The masks of the following rotations
// are translated through the permutation; the permutation stays in force after the call (vqe_expectation evaluates a
// permuted twin of the Pauli sum; anything else that needs the caller's labelling undoes the swaps).  For the random C5
// program on 8 GPUs: 19 swaps instead of 68 peer passes.
static int rotations_impl(RankSet& rs, int n_rot, const uint64_t* xmask, const uint64_t* zmask, const int32_t* ny,
                          const double* angle, int buf = VQE_BUF_PSI) {
    int rc = check_rankset(rs);
    if (rc) return rc;
    vqe_ctx* c = rs.r[0];
    if (c->world == 1 || buf != VQE_BUF_PSI) return rotations_core(rs, n_rot, xmask, zmask, ny, angle, buf);
    if (n_rot < 0 || (n_rot > 0 && (!xmask || !zmask || !ny || !angle))) return fail(VQE_ERR_INVALID, "null array");
    const int RELABEL_FLOOR = std::max(2, std::min(env_int("VQE_RELABEL_FLOOR", 10), c->nl - 4));
    const bool relabel = env_int("VQE_RELABEL", 1) != 0 && c->nl >= 8;
    if (!relabel) {
        rc = need_caller_labelling(rs);
        if (rc) return rc;
        return rotations_core(rs, n_rot, xmask, zmask, ny, angle, buf);
    }
    const int n = c->n, nl = c->nl;
    const uint64_t full = (1ull << n) - 1ull;
    // for every logical bit: the rotations that flip it, ascending (zero-angle rotations are identities)
    std::vector<std::vector<int>> uses(n);
    for (int k = 0; k < n_rot; ++k) {
        if ((xmask[k] | zmask[k]) & ~full) return fail(VQE_ERR_INVALID, "rotation %d: mask has bits >= n_qubits", k);
        if (angle[k] == 0.0) continue;
        for (uint64_t m = xmask[k]; m; m &= m - 1) uses[__builtin_ctzll(m)].push_back(k);
    }
    std::vector<size_t> cursor(n, 0);
    auto next_use = [&](int bit, int k) {
        size_t& p = cursor[bit];
        while (p < uses[bit].size() && uses[bit][p] <= k) ++p;
        return p < uses[bit].size() ? uses[bit][p] : INT32_MAX;
    };
    std::vector<uint64_t> tx, tz;
    std::vector<int32_t> tny;
    std::vector<double> tang;
    auto flush = [&]() -> int {
        if (tx.empty()) return VQE_OK;
        int rcf = rotations_core(rs, (int)tx.size(), tx.data(), tz.data(), tny.data(), tang.data(), buf);
        tx.clear(); tz.clear(); tny.clear(); tang.clear();
        return rcf;
    };
    const uint64_t gmask = ((1ull << c->g) - 1ull) << nl;
    for (int k = 0; k < n_rot; ++k) {
        if (angle[k] == 0.0) continue;
        uint64_t px = perm_mask(c->perm, xmask[k]);
        while (px & gmask) {
            // the logical bit in the lowest touched global slot moves to the local slot whose bit is needed last
            const int gslot = __builtin_ctzll(px & gmask);
            int inv[64];
            for (int q = 0; q < n; ++q) inv[c->perm[q]] = q;
            int best = -1, best_use = -1;
            for (int slot = nl - 1; slot >= RELABEL_FLOOR; --slot) {
                const int lb = inv[slot];
                if ((xmask[k] >> lb) & 1ull) continue;
                const int nu = next_use(lb, k);
                if (nu > best_use) { best_use = nu; best = slot; }
            }
            if (best < 0) break;  // (cannot happen: a rotation flips at most tile_bits qubits) -> peer pass
            rc = flush();
            if (rc) return rc;
            rc = swap_global_local(rs, gslot, best);
            if (rc) return rc;
            px = perm_mask(c->perm, xmask[k]);
        }
        tx.push_back(px);
        tz.push_back(perm_mask(c->perm, zmask[k]));
        tny.push_back(ny[k]);
        tang.push_back(angle[k]);
    }
    return flush();
}
extern "C" int vqe_apply_pauli_rotations(vqe_ctx* c, int n_rot, const uint64_t* xmask, const uint64_t* zmask,
                                         const int32_t* ny, const double* angle) {
    RankSet rs;
    rs.r.push_back(c);
    return rotations_impl(rs, n_rot, xmask, zmask, ny, angle);
}
// the same ordered product applied to any of the context's buffers (adjoint gradient: the co-state lives in sigma)
extern "C" int vqe_apply_pauli_rotations_buf(vqe_ctx* c, int buf, int n_rot, const uint64_t* xmask, const uint64_t* zmask,
                                             const int32_t* ny, const double* angle) {
    RankSet rs;
    rs.r.push_back(c);
    return rotations_impl(rs, n_rot, xmask, zmask, ny, angle, buf);
}
// Host-only view of the pass planner (no CUDA call): how an ordered rotation list is cut into tile passes for
// a state of n_qubits with n_global rank bits.  Used by the CPU tests of the sharding logic and by bench.py to
// report local / peer pass counts.  pass_kind: 0 = local pass, 1 = peer pass in exchange form, 2 = peer pass in
// gather form (pattern in pass_pattern).
// Host-only interpreter of the item-table rotation kernel (no CUDA call; CPU test support): plans the rotation program
// for the real layout of an unsharded state exactly as vqe_apply_pauli_rotations does, builds the item table k_col_tab
// reads (build_coltab) and walks it tile by tile on a HOST state of 2^n doubles, with the kernel's own per-item
// arithmetic.  Returns VQE_ERR_INVALID when the program has no real-layout collapsed-run plan (then the GPU path does
// not use the table either).
extern "C" int vqe_debug_coltab_host(int n_qubits, int tile_bits, int low_bits, int form, int n_rot, const uint64_t* xmask,
                                     const uint64_t* zmask, const int32_t* ny, const double* angle, double* psi_re,
                                     int32_t* n_passes, int32_t* n_words) {
    if (n_qubits < 1 || n_qubits > 30) return fail(VQE_ERR_INVALID, "bad qubit count");
    if (tile_bits < 6 || tile_bits > 13) tile_bits = 13;
    tile_bits = std::min(tile_bits, n_qubits);
    if (low_bits < 0 || low_bits > tile_bits) low_bits = 4;
    if (n_rot < 0 || !psi_re || (n_rot > 0 && (!xmask || !zmask || !ny || !angle))) return fail(VQE_ERR_INVALID, "null array");
    std::vector<HostOp> ops;
    for (int k = 0; k < n_rot; ++k) {
        if (angle[k] == 0.0) continue;
        HostOp h = HostOp();
        h.kind = OP_ROT;
        h.x = xmask[k];
        h.z = zmask[k];
        h.ny = ny[k];
        h.c = cos(angle[k]);
        h.s = sin(angle[k]);
        h.ang = angle[k];
        if (h.x == 0 || fabs(h.c) < 0.3) return fail(VQE_ERR_INVALID, "rotation %d takes the general path", k);
        ops.push_back(h);
    }
    OpPlan plan;
    int rc = plan_ops(n_qubits, n_qubits, tile_bits, low_bits, tile_bits > 12 ? 1024 : 512, ops, plan);
    if (rc) return rc;
    for (const OpPass& ps : plan.passes) {
        const bool all_col = ps.fast && !ps.has_imag && ps.sub_end == ps.sub_begin && ps.col_end > ps.col_begin && ps.pass_scale == 1.0 &&
                             (ps.sup_end - ps.sup_begin) == (ps.col_end - ps.col_begin);
        if (!all_col || ps.tp.vbit) return fail(VQE_ERR_INVALID, "the program has a pass that is not a real collapsed-run pass");
    }
    rc = build_coltab(plan, true);
    if (rc) return rc;
    if (n_passes) *n_passes = (int32_t)plan.passes.size();
    if (n_words) *n_words = (int32_t)plan.coltab.size();
    for (size_t p = 0; p < plan.passes.size(); ++p) {
        const OpPass& ps = plan.passes[p];
        const uint32_t ts = 1u << ps.tp.tbits, swz = plan.coltab_swz[p];
        const uint32_t lmask = (1u << ps.tp.lbits) - 1u;
        const uint32_t* tab = plan.coltab.data() + plan.coltab_off[p];
        std::vector<double> tile(ts);
        std::vector<uint64_t> addr(ts);
        for (uint64_t t = 0; t < ps.tp.n_tiles; ++t) {
            uint64_t base = 0, v = t, m = ps.tp.comp_mask;
            while (m) {  // pdep
                const uint64_t low = m & (0 - m);
                if (v & 1) base |= low;
                v >>= 1;
                m ^= low;
            }
            for (uint32_t k = 0; k < ts; ++k) {
                addr[swz_idx8(k, swz)] = base | ps.tp.scat[k >> ps.tp.lbits] | (uint64_t)(k & lmask);
                tile[swz_idx8(k, swz)] = psi_re[base | ps.tp.scat[k >> ps.tp.lbits] | (uint64_t)(k & lmask)];
            }
            // the 16-bit segment table k_col_stab reads: [first segment of every run][segments of 512 slots]
            const uint16_t* t16 = plan.coltab16.data() + plan.coltab16_off[p];
            const size_t n_cols = ps.col_end - ps.col_begin;
            const uint16_t* seg0 = t16 + ((2 * n_cols + 7) & ~size_t(7));
            for (size_t q = ps.col_begin; q < ps.col_end; ++q) {
                const DevCol& co = plan.dcols[q];
                const uint32_t lxb = swz_idx8(co.lx, swz) << 3, items = co.n_active << co.free_log;
                const uint32_t cs = (uint32_t)__builtin_popcountll(base & co.zout) & 1u;
                if (form == 0) {  // 32-bit words (k_col_tab)
                    for (uint32_t it = 0; it < items; ++it) {
                        const uint32_t w = tab[co.pad + it];
                        const DevColEntry& en = plan.dents[ps.ent_begin + co.ent_begin + (co.n_active > 1u ? (w >> 17) : 0u)];
                        const double sn = (((w >> 16) ^ cs) & 1u) ? -en.s : en.s;
                        const uint32_t ia = (w & 0xffffu) >> 3, ib = ((w & 0xffffu) ^ lxb) >> 3;
                        const double a = tile[ia], b = tile[ib];
                        tile[ia] = fma(en.c, a, -sn * b);
                        tile[ib] = fma(en.c, b, sn * a);
                    }
                    continue;
                }
                uint32_t sg = reinterpret_cast<const uint32_t*>(t16)[q - ps.col_begin];
                const uint32_t per_pat = std::max<uint32_t>(1u, (1u << co.free_log) / COLSEG);
                uint32_t seen = 0;
                for (uint32_t pi = 0; pi < co.n_active; ++pi) {
                    const DevColEntry& en = plan.dents[ps.ent_begin + co.ent_begin + pi];
                    for (uint32_t k = 0; k < per_pat; ++k, ++sg)
                        for (uint32_t sl = 0; sl < COLSEG; ++sl) {
                            const uint32_t wl = seg0[(size_t)sg * COLFACT + (sl & 31u)], wg = seg0[(size_t)sg * COLFACT + 32u + (sl >> 5)];
                            if (wl == 0xffffu || wg == 0xffffu) continue;
                            const uint32_t w = wl ^ wg;
                            ++seen;
                            const double sn = ((w ^ cs) & 1u) ? -en.s : en.s;
                            const uint32_t ia = (w & 0xfff8u) >> 3, ib = ((w & 0xfff8u) ^ lxb) >> 3;
                            const double a = tile[ia], b = tile[ib];
                            tile[ia] = fma(en.c, a, -sn * b);
                            tile[ib] = fma(en.c, b, sn * a);
                        }
                }
                if (seen != items) return fail(VQE_ERR_INVALID, "segment table of pass %zu holds %u items of run %zu, expected %u", p, seen, q - ps.col_begin, items);
            }
            for (uint32_t k = 0; k < ts; ++k) psi_re[addr[k]] = tile[k];
        }
    }
    return VQE_OK;
}

// Host-only census of the shared-memory banks the item tables of k_col_stab address (no CUDA call; CPU test support).  A 64-bit
// shared-memory access is served per half-warp: 16 lanes x 8 bytes = one 128-byte wavefront when the 16 elements sit in 16
// different 8-byte bank pairs (bits 3-6 of the byte offset).  For every segment of every run of the plan the a-side offsets of
// the 32 half-warps are checked.  *accesses = half-warp accesses, *conflicting = those with two lanes in one bank pair,
// *unavoidable = those of runs whose free tile positions cannot reach all 16 bank pairs whatever the item order (tile position 0
// fixed, or both positions that feed one bank bit).
extern "C" int vqe_debug_coltab_banks(int n_qubits, int tile_bits, int low_bits, int n_rot, const uint64_t* xmask,
                                      const uint64_t* zmask, const int32_t* ny, const double* angle, int64_t* accesses,
                                      int64_t* conflicting, int64_t* unavoidable) {
    if (n_qubits < 1 || n_qubits > 30) return fail(VQE_ERR_INVALID, "bad qubit count");
    if (tile_bits < 6 || tile_bits > 13) tile_bits = 13;
    tile_bits = std::min(tile_bits, n_qubits);
    if (low_bits < 0 || low_bits > tile_bits) low_bits = 4;
    if (n_rot < 0 || !accesses || !conflicting || !unavoidable || (n_rot > 0 && (!xmask || !zmask || !ny || !angle)))
        return fail(VQE_ERR_INVALID, "null array");
    std::vector<HostOp> ops;
    for (int k = 0; k < n_rot; ++k) {
        if (angle[k] == 0.0) continue;
        HostOp h = HostOp();
        h.kind = OP_ROT;
        h.x = xmask[k];
        h.z = zmask[k];
        h.ny = ny[k];
        h.c = cos(angle[k]);
        h.s = sin(angle[k]);
        h.ang = angle[k];
        ops.push_back(h);
    }
    OpPlan plan;
    int rc = plan_ops(n_qubits, n_qubits, tile_bits, low_bits, tile_bits > 12 ? 1024 : 512, ops, plan);
    if (rc) return rc;
    rc = build_coltab(plan, false);
    if (rc) return rc;
    *accesses = *conflicting = *unavoidable = 0;
    for (size_t p = 0; p < plan.passes.size(); ++p) {
        const OpPass& ps = plan.passes[p];
        const size_t n_cols = ps.col_end - ps.col_begin;
        const uint16_t* t16 = plan.coltab16.data() + plan.coltab16_off[p];
        const uint16_t* seg0 = t16 + ((2 * n_cols + 7) & ~size_t(7));
        for (size_t q = ps.col_begin; q < ps.col_end; ++q) {
            const DevCol& co = plan.dcols[q];
            // unavoidable: the images of the run's FREE tile positions under the swizzle do not span the four bank-pair bits
            // (position 0 fixed, or both positions that feed one bank bit -- p and p + 3 for p = 1, 2, 3 -- fixed)
            bool fixes0 = false;
            {
                uint32_t fixed = 0, basis[4] = {0, 0, 0, 0};
                int rank = 0;
                for (uint32_t d = 0; d < co.nd && d < 6; ++d) fixed |= 1u << __builtin_popcount(~co.dpos[d]);
                for (uint32_t b2 = 0; b2 < (uint32_t)ps.tp.tbits; ++b2) {
                    if ((fixed >> b2) & 1u) continue;
                    uint32_t v = swz_idx8(1u << b2, plan.coltab_swz[p]) & 0xfu;
                    for (int b3 = 3; b3 >= 0 && v; --b3)
                        if (((v >> b3) & 1u) && basis[b3]) v ^= basis[b3];
                    if (v) {
                        basis[31 - __builtin_clz(v)] = v;
                        ++rank;
                    }
                }
                fixes0 = rank < 4;
            }
            uint32_t sg = reinterpret_cast<const uint32_t*>(t16)[q - ps.col_begin];
            const uint32_t per_pat = std::max<uint32_t>(1u, (1u << co.free_log) / COLSEG);
            for (uint32_t k = 0; k < co.n_active * per_pat; ++k, ++sg)
                for (uint32_t hw = 0; hw < COLSEG / 16u; ++hw) {
                    uint32_t seen = 0, lanes = 0;
                    for (uint32_t ln = 0; ln < 16; ++ln) {
                        const uint32_t sl = hw * 16u + ln;
                        const uint32_t wl = seg0[(size_t)sg * COLFACT + (sl & 31u)], wg = seg0[(size_t)sg * COLFACT + 32u + (sl >> 5)];
                        if (wl == 0xffffu || wg == 0xffffu) continue;
                        seen |= 1u << ((((wl ^ wg) & 0xfff8u) >> 3) & 15u);
                        ++lanes;
                    }
                    if (!lanes) continue;
                    ++*accesses;
                    if ((uint32_t)__builtin_popcount(seen) < lanes) {
                        ++*conflicting;
                        if (fixes0) ++*unavoidable;
                    }
                }
        }
    }
    return VQE_OK;
}

extern "C" int vqe_plan_rotations(int n_qubits, int n_global, int tile_bits, int low_bits, int n_rot,
                                  const uint64_t* xmask, const uint64_t* zmask, const int32_t* ny, const double* angle,
                                  int cap, int32_t* n_passes, int32_t* pass_kind, uint64_t* pass_pattern,
                                  int32_t* pass_n_ops, uint64_t* pass_tile_mask) {
    if (n_qubits < 1 || n_qubits > 40 || n_global < 0 || n_global > 6 || n_global >= n_qubits)
        return fail(VQE_ERR_INVALID, "bad qubit counts");
    if (tile_bits < 6 || tile_bits > 13) tile_bits = 12;  // 13: the real-layout plan of a purely real state (tiles of doubles)
    if (low_bits < 0 || low_bits > tile_bits) low_bits = 5;
    if (!n_passes || n_rot < 0 || (n_rot > 0 && (!xmask || !zmask || !ny || !angle))) return fail(VQE_ERR_INVALID, "null array");
    std::vector<HostOp> ops;
    for (int k = 0; k < n_rot; ++k) {
        if (angle[k] == 0.0) continue;
        HostOp h = HostOp();
        h.kind = OP_ROT;
        h.x = xmask[k];
        h.z = zmask[k];
        h.ny = ny[k];
        h.c = cos(angle[k]);
        h.s = sin(angle[k]);
        h.ang = angle[k];
        ops.push_back(h);
    }
    OpPlan plan;
    int rc = plan_ops(n_qubits, n_qubits - n_global, tile_bits, low_bits, tile_bits > 12 ? 1024 : 512, ops, plan);
    if (rc) return rc;
    *n_passes = (int32_t)plan.passes.size();
    for (size_t p = 0; p < plan.passes.size() && (int)p < cap; ++p) {
        if (pass_kind) pass_kind[p] = plan.passes[p].tp.vbit ? (plan.passes[p].gather ? 2 : 1) : 0;
        if (pass_pattern) pass_pattern[p] = plan.passes[p].tp.gpat;
        if (pass_n_ops) pass_n_ops[p] = (int32_t)(plan.passes[p].op_end - plan.passes[p].op_begin);
        if (pass_tile_mask) pass_tile_mask[p] = plan.passes[p].tp.tile_mask;
    }
    if (getenv("VQE_DEBUG_PLAN") && atoi(getenv("VQE_DEBUG_PLAN")) >= 3) {
        for (size_t pi = 0; pi < plan.passes.size(); ++pi) {
            const OpPass& p = plan.passes[pi];
            size_t ncol = 0, norb = 0, items = 0, nsub = 0, nrot_orb = 0;
            for (size_t si = p.sup_begin; si < p.sup_end; ++si) {
                const DevSuper& su = plan.dsupers[si];
                if (su.sub_count == 0xffffffffu) {
                    ++ncol;
                    const DevCol& co = plan.dcols[p.col_begin + su.sub_begin];
                    items += (size_t)co.n_active << co.free_log;
                } else {
                    ++norb;
                    nsub += su.sub_count;
                    for (uint32_t sb = 0; sb < su.sub_count; ++sb) nrot_orb += plan.dsubs[p.sub_begin + su.sub_begin + sb].len;
                }
            }
            fprintf(stderr, "[rotpass] %zu ops %zu supers %zu col %zu items %zu orb %zu subs %zu orbrots %zu lbits %d tbits %d scale %d\n", pi,
                    p.op_end - p.op_begin, p.sup_end - p.sup_begin, ncol, items, norb, nsub, nrot_orb, p.tp.lbits, p.tp.tbits,
                    p.pass_scale != 1.0);
        }
    }
    if (getenv("VQE_DEBUG_PLAN")) {
        size_t nruns = plan.dsubs.size(), nfast = 0, lens[9] = {0};
        for (const OpPass& p : plan.passes) nfast += p.fast ? 1 : 0;
        for (const DevSub& r : plan.dsubs) lens[std::min<uint32_t>(r.len, 8)]++;
        fprintf(stderr, "[plan] passes %zu (fast %zu) ops %zu segments %zu (collapsed runs %zu, %zu active patterns) sub-runs %zu; sub-run-length histogram 1..8+:",
                plan.passes.size(), nfast, plan.dops.size(), plan.dsupers.size(), plan.dcols.size(), plan.dents.size(), nruns);
        for (int k = 1; k <= 8; ++k) fprintf(stderr, " %zu", lens[k]);
        size_t npeer = 0, ngather = 0;
        double frac = 0.0;
        for (const OpPass& p : plan.passes) {
            if (!p.tp.vbit) continue;
            ++npeer;
            if (p.gather) {
                ++ngather;
                frac += (double)std::max(p.need_lo.size(), p.need_hi.size()) * 4.0 / (double)((1u << p.tp.tbits) >> 1);
            }
        }
        fprintf(stderr, "; peer passes %zu, gather form %zu (mean needed fraction of the partner half %.3f)\n", npeer, ngather,
                ngather ? frac / ngather : 0.0);
    }
    return VQE_OK;
}

// Tabulated plane rotations: the unitary of a gate template that only couples basis states differing by a fixed
// X-mask, given as (pattern, cos, sin) per coupled pair class.  Replaces the gate-by-gate execution of the QUCCSD
// excitation circuits (openvqe/common_files/circuit.py:13-93 as run at get_energy_qucc.py:50-55): the host derives
// the table of a template once from its 2- or 4-qubit unitary (openvqe_b200/common_files/circuit.py).
extern "C" int vqe_apply_plane_rotations(vqe_ctx* c, int n_ops, const uint64_t* xmask, const int32_t* tab_offsets,
                                         const uint64_t* pattern, const double* cosv, const double* sinv) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    if (n_ops < 0 || (n_ops > 0 && (!xmask || !tab_offsets))) return fail(VQE_ERR_INVALID, "null array");
    {
        int rcl = need_caller_labelling(c);
        if (rcl) return rcl;
    }
    const uint64_t full = (1ull << c->n) - 1ull;
    std::vector<HostOp> ops;
    ops.reserve(n_ops);
    for (int k = 0; k < n_ops; ++k) {
        if (xmask[k] == 0 || (xmask[k] & ~full)) return fail(VQE_ERR_INVALID, "plane rotation %d: bad X-mask", k);
        if (c->g && (xmask[k] >> c->nl))
            return fail(VQE_ERR_INVALID, "plane rotation %d touches a global qubit of a sharded state (use vqe_apply_gates)", k);
        const uint64_t top = 1ull << (63 - __builtin_clzll(xmask[k]));
        HostOp h = HostOp();
        h.kind = OP_PLANE;
        h.x = xmask[k];
        for (int q = tab_offsets[k]; q < tab_offsets[k + 1]; ++q) {
            if ((pattern[q] & ~xmask[k]) || (pattern[q] & top))
                return fail(VQE_ERR_INVALID, "plane rotation %d: pattern must lie inside the X-mask with its highest bit clear", k);
            if (sinv[q] == 0.0 && cosv[q] == 1.0) continue;  // identity on this pattern
            h.ppat.push_back(pattern[q]);
            h.pcos.push_back(cosv[q]);
            h.psin.push_back(sinv[q]);
        }
        if (h.ppat.empty()) continue;
        ops.push_back(h);
    }
    return run_ops(c, ops);
}

// psi <- (re + i im) psi  (global phase / scale; the QUCCSD single-excitation template carries e^{i pi/4})
extern "C" int vqe_scale_state(vqe_ctx* c, int b, double re, double im) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    CK(cudaSetDevice(c->device));
    int rc = ensure_buf(c, b);
    if (rc) return rc;
    rc = ensure_complex(c, b);
    if (rc) return rc;
    if (b == VQE_BUF_PSI && im != 0.0) c->psi_real = false;
    k_axpby<<<grid_1d(c, c->n_amp, 256), 256, 0, c->stream>>>(c->buf[b], c->buf[b], c->n_amp, re, im, 0.0, 0.0);
    c->launches++;
    CK(cudaGetLastError());
    return VQE_OK;
}

// dst = alpha * x + beta * dst on two buffers of the context (BLAS-1 helper of the Lanczos ground state and of the
// host-driven Taylor exponential on sharded states; purely local, no peer traffic)
extern "C" int vqe_axpby(vqe_ctx* c, int dst, int x, double a_re, double a_im, double b_re, double b_im) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    CK(cudaSetDevice(c->device));
    int rc = ensure_buf(c, dst);
    if (rc) return rc;
    rc = ensure_buf(c, x);
    if (rc) return rc;
    rc = ensure_complex(c, dst);
    if (rc == VQE_OK) rc = ensure_complex(c, x);
    if (rc == VQE_OK && (dst == VQE_BUF_PSI) != (x == VQE_BUF_PSI)) rc = need_caller_labelling(c);
    if (rc) return rc;
    if (dst == VQE_BUF_PSI) c->psi_real = false;
    k_axpby<<<grid_1d(c, c->n_amp, 256), 256, 0, c->stream>>>(c->buf[dst], c->buf[x], c->n_amp, a_re, a_im, b_re, b_im);
    c->launches++;
    CK(cudaGetLastError());
    return VQE_OK;
}

static RankSet rankset_of(vqe_ctx* const* ranks, int n_ranks) {
    RankSet rs;
    for (int k = 0; ranks && k < n_ranks; ++k) rs.r.push_back(ranks[k]);
    return rs;
}
extern "C" int vqe_group_apply_pauli_rotations(vqe_ctx* const* ranks, int n_ranks, int n_rot, const uint64_t* xmask,
                                               const uint64_t* zmask, const int32_t* ny, const double* angle) {
    RankSet rs = rankset_of(ranks, n_ranks);
    return rotations_impl(rs, n_rot, xmask, zmask, ny, angle);
}

static int gates_impl(RankSet& rs, int n_gates, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                      const double* angle) {
    int rc0 = check_rankset(rs);
    if (rc0 == VQE_OK) rc0 = need_caller_labelling(rs);
    if (rc0) return rc0;
    vqe_ctx* c = rs.r[0];
    if (n_gates < 0 || (n_gates > 0 && (!kind || !q0))) return fail(VQE_ERR_INVALID, "null array");
    std::vector<HostOp> ops;
    ops.reserve(n_gates);
    const double r2 = 0.70710678118654752440;
    for (int k = 0; k < n_gates; ++k) {
        if (q0[k] < 0 || q0[k] >= c->n) return fail(VQE_ERR_INVALID, "gate %d: qubit %d out of range", k, q0[k]);
        HostOp h = HostOp();
        h.x = 1ull << (c->n - 1 - q0[k]);
        double th = angle ? angle[k] : 0.0;
        double cs = cos(0.5 * th), sn = sin(0.5 * th);
        switch (kind[k]) {
            case VQE_GATE_X: h.kind = OP_GATE1; h.m[2] = 1; h.m[4] = 1; break;
            case VQE_GATE_H: h.kind = OP_GATE1; h.m[0] = r2; h.m[2] = r2; h.m[4] = r2; h.m[6] = -r2; break;
            case VQE_GATE_RX: h.kind = OP_GATE1; h.m[0] = cs; h.m[3] = -sn; h.m[5] = -sn; h.m[6] = cs; break;
            case VQE_GATE_RY: h.kind = OP_GATE1; h.m[0] = cs; h.m[2] = -sn; h.m[4] = sn; h.m[6] = cs; break;
            case VQE_GATE_RZ: h.kind = OP_GATE1; h.m[0] = cs; h.m[1] = -sn; h.m[6] = cs; h.m[7] = sn; break;
            case VQE_GATE_CNOT:
                if (!q1 || q1[k] < 0 || q1[k] >= c->n || q1[k] == q0[k])
                    return fail(VQE_ERR_INVALID, "gate %d: bad CNOT target", k);
                h.kind = OP_CNOT;
                h.z = 1ull << (c->n - 1 - q0[k]);  // control
                h.x = 1ull << (c->n - 1 - q1[k]);  // target
                break;
            default: return fail(VQE_ERR_INVALID, "gate %d: unknown kind %d", k, kind[k]);
        }
        ops.push_back(h);
    }
    return run_ops(rs, ops);
}
extern "C" int vqe_apply_gates(vqe_ctx* c, int n_gates, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                               const double* angle) {
    RankSet rs;
    rs.r.push_back(c);
    return gates_impl(rs, n_gates, kind, q0, q1, angle);
}
extern "C" int vqe_group_apply_gates(vqe_ctx* const* ranks, int n_ranks, int n_gates, const int32_t* kind,
                                     const int32_t* q0, const int32_t* q1, const double* angle) {
    RankSet rs = rankset_of(ranks, n_ranks);
    return gates_impl(rs, n_gates, kind, q0, q1, angle);
}

// ---- Pauli sums -----------------------------------------------------------------------------
struct PSPass {
    TilePlan tp;
    std::vector<DevGroup> groups;
    std::vector<DevTerm> terms_expect;  // pair weights (expectation)
    std::vector<DevTerm> terms_apply;   // c_k i^ny   (apply)
    std::vector<DevGCol> gcols;         // collapsed groups (expectation)
    std::vector<DevGColEntry> gents;
    std::vector<DevFlat> flats;
    std::vector<uint64_t> fzout;        // distinct outside-tile Z masks of the flat entries
    std::vector<DevAFlat> aflat;        // apply: flat entries of the collapsed groups, group after group
    std::vector<uint32_t> aoff;         // apply: groups.size() + 1 offsets into aflat (empty range: term loop)
    std::vector<uint64_t> azout;        // apply: distinct outside-tile Z masks
    DevAFlat* d_aflat = nullptr;
    uint32_t* d_aoff = nullptr;
    uint64_t* d_azout = nullptr;
    DevFlat* d_flats = nullptr;
    uint64_t* d_fzout = nullptr;
    DevGCol* d_gcols = nullptr;
    DevGColEntry* d_gents = nullptr;
    // device copies
    DevGroup* d_groups = nullptr;
    DevTerm* d_terms_expect = nullptr;
    DevTerm* d_terms_apply = nullptr;
    uint64_t* d_scat = nullptr;
    bool cplx = false;  // some expectation weight has a non-zero imaginary part
    // lean pass (see DevFlat2): every group is tabulated per occupation pattern; none of the tables above is used
    bool lean = false;
    bool swz = false;                   // the entries' per-j offsets and X-masks are stored in the 128-byte-swizzled layout
    std::vector<DevFlat2> flats2;       // entries, group after group
    std::vector<uint32_t> goff;         // n_lean_groups + 1 entry counts (statistics; the entries are re-ordered into batches)
    std::vector<uint32_t> boff;         // n_batches + 1 offsets into flats2: sigma = O psi walks the batches (see batch_lean_entries)
    std::vector<double> addtab;         // T_lo[32] + T_hi[...] of every additive pattern
    std::vector<DevAddPat> addpat;
    std::vector<DevAddOut> addout;
    size_t lean_terms = 0;              // Pauli strings folded into the entries (statistics)
    DevFlat2* d_flats2 = nullptr;
    DevFlat2* d_flats2_rl = nullptr;    // the entries in real-layout form (expectation on a state kept as n_amp doubles)
    bool rl_ok = false, rl_swz = false; // the pass has a real-layout tensor-map form | whose tile is swizzled
    bool diag_only = false;             // general pass that holds nothing but X-mask-0 groups (see vqe_paulisum::diag2)
    // REAL-LAYOUT TWIN with a LANE TABLE (expectation on a state kept as n_amp doubles; k_expect_rlp / k_expect_rl2): the same
    // groups lowered a second time, directly into real-layout offsets, with the free tile positions assigned so that the 16
    // lanes of a half-warp hit 16 different 8-byte bank pairs (see lower_lean_group, rl_order), and the per-lane part of an
    // entry (index deposit, pattern, swizzle, parity) tabulated: 32 words of (byte offset | parity) per entry.
    std::shared_ptr<PSPass> rlp;        // flats2 / addtab / addpat / addout / fzout of the twin; rlp->swz = the real-layout swizzle
    std::vector<uint16_t> lane_rl;      // (in the twin) 32 lane words per entry
    bool rl_lane = false;               // the twin is uploaded and in use (d_flats2_rl then holds ITS entries)
    uint16_t* d_lane_rl = nullptr;
    double* d_addtab_rl = nullptr;
    DevAddPat* d_addpat_rl = nullptr;
    DevAddOut* d_addout_rl = nullptr;
    uint64_t* d_fzout_rl = nullptr;
    uint32_t* d_goff = nullptr;
    double* d_addtab = nullptr;
    DevAddPat* d_addpat = nullptr;
    DevAddOut* d_addout = nullptr;
};
struct HTerm {
    uint64_t x, z;
    int ny;
    double cr, ci;
};
struct vqe_paulisum {
    int n = 0, nl = 0, device = 0, n_groups = 0;
    std::vector<PSPass> passes;
    // the strings as given (caller's labelling) and the planner settings: a relabelled sharded state (vqe_ctx::perm) is
    // evaluated with a twin of the sum whose masks are translated through the permutation, built on first use
    std::vector<HTerm> terms;
    int tile_bits = 12, low_bits = 5, threads = 512;
    std::map<std::string, vqe_paulisum*> variants;
    // The diagonal part (X-mask 0) as a quadratic form in the Z letters, when it is one (real weights, strings of at most
    // two Z letters) and sits alone in its general passes: evaluated on the real layout by k_expect_diag2_rl.
    struct Diag2 {
        bool on = false;
        int tb = 0;                     // chunk bits: min(13, nl)
        double c0 = 0.0;
        std::vector<double> a, b, t;    // a[n], b[n * n] (p < q), t[2^tb]
        double *d_a = nullptr, *d_b = nullptr, *d_t = nullptr;
    } diag2;
};

static void mul_i_pow(double& r, double& i, int k) {
    k &= 3;
    double a = r, b = i;
    if (k == 1) { r = -b; i = a; }
    else if (k == 2) { r = -a; i = -b; }
    else if (k == 3) { r = b; i = -a; }
}

// ---- lean groups (DevFlat2) -----------------------------------------------------------------------
#define LEAN_ENT_CAP 8192     // entries per lean pass (384 KiB of descriptors, streamed through L1/L2)
#define LEAN_PAT_CAP 1024     // additive patterns per lean pass (per-tile constants in shared memory: 8 KiB)
#define LEAN_TAB_CAP 60000    // doubles in the additive tables of a pass (16-bit offsets)

// pass-independent part of the eligibility test (full-width masks)
static bool lean_eligible(uint64_t x, const std::vector<HTerm>& terms, int nl, int tbits) {
    if (x == 0 || (x >> nl) != 0 || terms.empty()) return false;
    const int nx = popc64(x);
    if (nx > 4 || tbits - nx < 8) return false;
    for (const HTerm& t : terms) {
        if (t.ci != 0.0 || (t.ny & 1)) return false;
        if (popc64((t.z ^ terms[0].z) & ~x) > 1) return false;
    }
    return true;
}

// Lower one eligible group into entries of pass p (appended; nothing is appended when the pass capacities would be
// exceeded -> false).  See the comment block above DevFlat2 for the algebra.
// rl_order: lower into the REAL-LAYOUT twin (PSPass::rlp): offsets in bytes of 8-byte elements, swizzled with p.swz as the
// real-layout swizzle, 32 lane words per entry in p.lane_rl, and the free tile positions re-ordered -- entry bits 0-3 (the 16
// lanes of a half-warp, which a 64-bit shared-memory access is served for) go to free positions whose images under the
// swizzle are linearly independent in the four bank-pair bits, the rest ascending.  The plain ascending order conflicts
// 2-way whenever the X-mask holds tile position 0, 2 or 3: 36 % of the wavefronts of an H2O expectation pass were replays.
static bool lower_lean_group(PSPass& p, uint64_t x, const std::vector<HTerm>& terms, bool rl_order = false) {
    const TilePlan& tp = p.tp;
    const uint32_t lx = plan_lx(x, tp);
    const int nx = __builtin_popcount(lx);
    const int tb = tp.tbits, free_log = tb - nx;
    const uint32_t hb = 31 - __builtin_clz(lx);
    std::vector<uint32_t> xpos, fpos;
    for (int b = 0; b < tb; ++b) ((lx >> b) & 1u ? xpos : fpos).push_back((uint32_t)b);
    if (rl_order) {
        const uint32_t swr = p.swz ? 0x70u : 0u;
        std::vector<uint32_t> first, rest;
        uint32_t basis[4] = {0, 0, 0, 0};
        for (uint32_t b : fpos) {
            uint32_t v = first.size() < 4 ? (swz_idx8(1u << b, swr) & 0xfu) : 0u;
            for (int b3 = 3; b3 >= 0 && v; --b3)
                if (((v >> b3) & 1u) && basis[b3]) v ^= basis[b3];
            if (v) {
                basis[31 - __builtin_clz(v)] = v;
                first.push_back(b);
            } else rest.push_back(b);
        }
        fpos = first;
        fpos.insert(fpos.end(), rest.begin(), rest.end());
    }
    const uint32_t lz0 = plan_lz(terms[0].z, tp);
    const uint64_t zout0 = plan_zout(terms[0].z, tp);
    struct LT {
        double w;
        uint32_t dx;      // Z difference to the first string on the X positions
        int r;            // further in-tile Z letter (tile bit), or -1
        uint64_t dout;    // further outside-tile Z letter (mask), or 0
    };
    std::vector<LT> lt;
    double scale = 0.0;
    for (const HTerm& t : terms) {
        LT e;
        double wr = t.cr, wi = 0.0;
        mul_i_pow(wr, wi, t.ny);  // even ny: real
        e.w = wr;
        const uint32_t dl = plan_lz(t.z, tp) ^ lz0;
        e.dx = dl & lx;
        const uint32_t dr = dl & ~lx;
        e.dout = plan_zout(t.z, tp) ^ zout0;
        if (__builtin_popcount(dr) + popc64(e.dout) > 1) return false;  // cannot happen for eligible groups
        e.r = dr ? __builtin_ctz(dr) : -1;
        scale += fabs(e.w);
        lt.push_back(e);
    }
    auto pdep_free = [&](uint32_t v) {
        uint32_t o = 0;
        for (size_t k = 0; k < fpos.size(); ++k)
            if ((v >> k) & 1u) o |= 1u << fpos[k];
        return o;
    };
    std::vector<DevFlat2> ents;
    std::vector<uint16_t> lanes;
    std::vector<double> tabs;
    std::vector<DevAddPat> pats;
    std::vector<DevAddOut> outs;
    uint32_t zsel = 0;
    while (zsel < p.fzout.size() && p.fzout[zsel] != zout0) ++zsel;
    const uint32_t n_chunks = 1u << (free_log - 8);
    for (uint32_t pi = 0; pi < (1u << nx); ++pi) {
        uint32_t pat = 0;
        for (int b = 0; b < nx; ++b)
            if ((pi >> b) & 1u) pat |= 1u << xpos[b];
        if ((pat >> hb) & 1u) continue;  // a-side only
        double beta0 = 0.0;
        std::vector<double> beta_r(tb, 0.0);
        std::vector<std::pair<uint64_t, double>> bout;
        for (const LT& e : lt) {
            const double val = (__builtin_popcount(pat & e.dx) & 1) ? -e.w : e.w;
            if (e.r >= 0) beta_r[e.r] += val;
            else if (e.dout) {
                size_t q = 0;
                while (q < bout.size() && bout[q].first != e.dout) ++q;
                if (q == bout.size()) bout.push_back({e.dout, 0.0});
                bout[q].second += val;
            } else beta0 += val;
        }
        const double tiny = 1e-15 * scale;
        if (fabs(beta0) <= tiny) beta0 = 0.0;
        bool any_r = false;
        for (double& b : beta_r) {
            if (fabs(b) <= tiny) b = 0.0;
            any_r = any_r || b != 0.0;
        }
        std::vector<std::pair<uint64_t, double>> bout2;
        for (auto& bo : bout)
            if (fabs(bo.second) > tiny) bout2.push_back(bo);
        const bool additive = any_r || !bout2.empty();
        if (!additive && beta0 == 0.0) continue;  // the group does not couple this occupation pattern
        uint32_t tab = 0xffffu, bidx = 0;
        if (additive) {
            tab = (uint32_t)(p.addtab.size() + tabs.size());
            bidx = (uint32_t)(p.addpat.size() + pats.size());
            for (uint32_t v = 0; v < 32; ++v) {
                double acc = 0.0;
                for (int k = 0; k < 5 && k < free_log; ++k) acc += ((v >> k) & 1u) ? -beta_r[fpos[k]] : beta_r[fpos[k]];
                tabs.push_back(acc);
            }
            for (uint32_t u2 = 0; u2 < (1u << (free_log - 5)); ++u2) {
                double acc = 0.0;
                for (int k = 5; k < free_log; ++k) acc += ((u2 >> (k - 5)) & 1u) ? -beta_r[fpos[k]] : beta_r[fpos[k]];
                tabs.push_back(acc);
            }
            DevAddPat ap;
            ap.beta0 = beta0;
            ap.out_begin = (uint32_t)(p.addout.size() + outs.size());
            ap.out_count = (uint32_t)bout2.size();
            for (auto& bo : bout2) outs.push_back({bo.second, bo.first});
            pats.push_back(ap);
        }
        for (uint32_t ch = 0; ch < n_chunks; ++ch) {
            DevFlat2 fl;
            memset(&fl, 0, sizeof fl);
            const uint32_t sw = p.swz ? 0x70u : 0u;
            const uint32_t esh = rl_order ? 3u : 4u;   // log2 of the tile element size
            fl.fr = additive ? 2.0 : 2.0 * beta0;
            fl.lx16 = (uint16_t)swz_off(lx << esh, sw);
            fl.lz16 = (uint16_t)(lz0 << esh);
            fl.pat16 = (uint16_t)(pat << esh);
            fl.zsel = (uint16_t)zsel;
            // (the twin's lane part comes from its lane table: no deposit masks)
            for (int k = 0; k < 4; ++k) fl.hm16[k] = (k < nx && !rl_order) ? (uint16_t)(~((1u << (xpos[k] + 4)) - 1u) & 0xffffu) : 0;
            uint32_t js = 0;
            for (uint32_t j = 0; j < 8; ++j) {
                const uint32_t oj = pdep_free((ch << 8) | (j << 5));
                fl.o16[j] = (uint16_t)swz_off(oj << esh, sw);
                if (__builtin_popcount(oj & lz0) & 1) js |= 1u << j;
            }
            if (rl_order)
                for (uint32_t ln = 0; ln < 32; ++ln) {
                    const uint32_t l = pdep_free(ln) | pat;
                    lanes.push_back((uint16_t)(swz_off(l << 3, sw) | ((uint32_t)__builtin_popcount(l & lz0) & 1u)));
                }
            fl.jsign = (uint16_t)js;
            fl.tab = (uint16_t)tab;
            fl.hi0 = additive ? (uint16_t)(tab + 32 + (ch << 3)) : 0;
            fl.bidx = (uint16_t)bidx;
            ents.push_back(fl);
        }
    }
    if (!p.flats2.empty() && (p.flats2.size() + ents.size() > LEAN_ENT_CAP || p.addpat.size() + pats.size() > LEAN_PAT_CAP ||
                              p.addtab.size() + tabs.size() > LEAN_TAB_CAP))
        return false;
    if (p.addtab.size() + tabs.size() > LEAN_TAB_CAP || p.addpat.size() + pats.size() > 0xffffu || zsel > 0xfffeu) return false;
    if (zsel == p.fzout.size()) p.fzout.push_back(zout0);
    if (p.goff.empty()) p.goff.push_back(0);
    p.flats2.insert(p.flats2.end(), ents.begin(), ents.end());
    p.lane_rl.insert(p.lane_rl.end(), lanes.begin(), lanes.end());
    p.goff.push_back((uint32_t)p.flats2.size());
    p.addtab.insert(p.addtab.end(), tabs.begin(), tabs.end());
    p.addpat.insert(p.addpat.end(), pats.begin(), pats.end());
    p.addout.insert(p.addout.end(), outs.begin(), outs.end());
    p.lean_terms += terms.size();
    return true;
}

// sigma = O psi adds  G(l) psi[l ^ x]  into an accumulator tile; entries whose (a, b) element sets are disjoint can be worked
// on between the same two barriers without any thread touching another's accumulator element.  The entries of a pass are
// therefore re-ordered into conflict-free BATCHES (first fit, in entry order: deterministic), instead of one barrier
// interval per X-mask group: a group has ~2 entries = 563 items for 512 threads (55 % of the slots busy), a batch holds
// several, and there are fewer barriers.  The expectation kernel is indifferent to the entry order.
static void batch_lean_entries(PSPass& p) {
    const size_t ne = p.flats2.size();
    p.boff.clear();
    p.boff.push_back(0);
    if (ne == 0) return;
    const uint32_t ts = 1u << p.tp.tbits;
    const size_t words = (ts + 63) / 64;
    const uint32_t sw = p.swz ? 0x70u : 0u;
    std::vector<std::vector<uint64_t>> touched(ne, std::vector<uint64_t>(words, 0));
    for (size_t e = 0; e < ne; ++e) {
        uint4 q[3];
        memcpy(q, &p.flats2[e], sizeof(DevFlat2));
        for (uint32_t lane = 0; lane < 32; ++lane) {
            LeanUnit u;
            lean_decode(q[0], q[1], q[2], lane, u);
            for (uint32_t j = 0; j < 8; ++j) {
                const uint32_t off = swz_off(u.v, sw) ^ u.o[j], offb = off ^ u.lx16;
                touched[e][(off >> 4) >> 6] |= 1ull << ((off >> 4) & 63u);
                touched[e][(offb >> 4) >> 6] |= 1ull << ((offb >> 4) & 63u);
            }
        }
    }
    const size_t max_batch = (size_t)std::max(1, env_int("VQE_APPLY_BATCH", 16));
    std::vector<std::vector<size_t>> batches;
    std::vector<std::vector<uint64_t>> used;
    for (size_t e = 0; e < ne; ++e) {
        size_t b = 0;
        for (; b < batches.size(); ++b) {
            if (batches[b].size() >= max_batch) continue;
            bool clash = false;
            for (size_t w = 0; w < words && !clash; ++w) clash = (used[b][w] & touched[e][w]) != 0;
            if (!clash) break;
        }
        if (b == batches.size()) {
            batches.emplace_back();
            used.emplace_back(words, 0);
        }
        batches[b].push_back(e);
        for (size_t w = 0; w < words; ++w) used[b][w] |= touched[e][w];
    }
    std::vector<DevFlat2> re;
    re.reserve(ne);
    for (const auto& bt : batches) {
        for (size_t e : bt) re.push_back(p.flats2[e]);
        p.boff.push_back((uint32_t)re.size());
    }
    p.flats2.swap(re);
}

// Tile bits of a pass, chosen greedily for COVERAGE: seed with `need`, repeatedly add the bit that completes the most
// open groups; when no single bit completes one, take the open group that needs the fewest new bits.
static uint64_t cover_greedy(const std::vector<uint64_t>& xs, const std::vector<size_t>& open, uint64_t need, uint64_t lowmask_pass,
                             int cap_bits, int nl, uint64_t lfull) {
    for (;;) {
        const uint64_t have = need | lowmask_pass;
        if (popc64(have) >= cap_bits) break;
        int best_bit = -1;
        size_t best_gain = 0;
        for (int b2 = 0; b2 < nl; ++b2) {
            if ((have >> b2) & 1ull) continue;
            const uint64_t with = have | (1ull << b2);
            size_t gain = 0;
            for (size_t g : open) {
                const uint64_t xl = xs[g] & lfull;
                if ((xl & ~with) == 0 && (xl & ~have) != 0) ++gain;
            }
            if (gain > best_gain) { best_gain = gain; best_bit = b2; }
        }
        if (best_bit >= 0) {
            need |= 1ull << best_bit;
            continue;
        }
        size_t best_g = xs.size();
        int best_missing = 1 << 30;
        for (size_t g : open) {
            const uint64_t xl = xs[g] & lfull;
            const int missing = popc64(xl & ~have);
            if (missing == 0 || popc64(have | xl) > cap_bits) continue;
            if (missing < best_missing) { best_missing = missing; best_g = g; }
        }
        if (best_g == xs.size()) break;
        need |= xs[best_g] & lfull;
    }
    return need;
}

static int build_paulisum(vqe_paulisum* ps, int n, int nl, int tbits_max, int low_bits, int threads_cfg,
                          std::vector<HTerm> terms, bool host_only = false) {
    {   // the diagonal part as a quadratic form (see vqe_paulisum::Diag2); the passes that hold it are marked at the end
        vqe_paulisum::Diag2& d = ps->diag2;
        d = vqe_paulisum::Diag2();
        bool ok = env_int("VQE_DIAG2", 1) != 0 && n <= 62, any = false;
        d.a.assign(n, 0.0);
        d.b.assign((size_t)n * n, 0.0);
        for (const HTerm& t : terms) {
            if (t.x != 0 || !ok) continue;
            any = true;
            const int wz = popc64(t.z);
            if (t.ci != 0.0 || wz > 2) { ok = false; break; }
            if (wz == 0) d.c0 += t.cr;
            else if (wz == 1) d.a[__builtin_ctzll(t.z)] += t.cr;
            else {
                const int p0 = __builtin_ctzll(t.z), q0 = 63 - __builtin_clzll(t.z);
                d.b[(size_t)p0 * n + q0] += t.cr;
            }
        }
        d.on = ok && any;
        if (d.on) {
            d.tb = std::min(13, nl);
            d.t.assign((size_t)1 << d.tb, 0.0);
            for (uint32_t tt = 0; tt < (1u << d.tb); ++tt) {
                double acc = 0.0;
                for (int p0 = 0; p0 < d.tb; ++p0)
                    for (int q0 = p0 + 1; q0 < d.tb; ++q0) {
                        const double bq = d.b[(size_t)p0 * n + q0];
                        if (bq != 0.0) acc += (((tt >> p0) ^ (tt >> q0)) & 1u) ? -bq : bq;
                    }
                d.t[tt] = acc;
            }
        }
    }
    // group by x (stable: keep first-appearance order of groups, term order inside)
    std::vector<uint64_t> xs;
    std::vector<std::vector<HTerm>> grp;
    {
        std::vector<std::pair<uint64_t, size_t>> keyed;
        for (size_t k = 0; k < terms.size(); ++k) keyed.push_back({terms[k].x, k});
        std::stable_sort(keyed.begin(), keyed.end(),
                         [](const std::pair<uint64_t, size_t>& a, const std::pair<uint64_t, size_t>& b) {
                             return a.first < b.first;
                         });
        for (size_t k = 0; k < keyed.size(); ++k) {
            if (k == 0 || keyed[k].first != keyed[k - 1].first) {
                xs.push_back(keyed[k].first);
                grp.emplace_back();
            }
            grp.back().push_back(terms[keyed[k].second]);
        }
    }
    ps->n = n;
    ps->nl = nl;
    ps->n_groups = (int)xs.size();
    const int lb = std::min(low_bits, std::min(tbits_max, nl));
    const uint64_t lfull = (1ull << nl) - 1ull;
    // split oversize groups so that each fits in the term cache
    {
        std::vector<uint64_t> xs2;
        std::vector<std::vector<HTerm>> grp2;
        for (size_t g = 0; g < xs.size(); ++g) {
            for (size_t o = 0; o < grp[g].size(); o += TERM_CAP) {
                xs2.push_back(xs[g]);
                grp2.emplace_back(grp[g].begin() + o, grp[g].begin() + std::min(grp[g].size(), o + TERM_CAP));
            }
        }
        xs.swap(xs2);
        grp.swap(grp2);
    }
    std::vector<char> done(xs.size(), 0);
    size_t remaining = xs.size();
    // ---- lean passes first: every group that can be tabulated per occupation pattern (all two-body groups and the
    // number-operator-dressed one-body groups of a molecular Hamiltonian).  No per-pass term tables, so a pass takes
    // EVERY open group its tile bits cover.
    if (env_int("VQE_EXP_LEAN", 1) != 0) {
        const int tb_local = std::min(tbits_max, nl);
        std::vector<char> lean_ok(xs.size(), 0);
        size_t n_lean = 0;
        for (size_t g = 0; g < xs.size(); ++g) {
            lean_ok[g] = lean_eligible(xs[g], grp[g], nl, tb_local) ? 1 : 0;
            n_lean += lean_ok[g];
        }
        // (A) tile-bit sets of the passes, greedily for coverage (as the general passes below)
        struct LeanPlan {
            TilePlan tp;
            bool swz;
        };
        std::vector<LeanPlan> lplans;
        std::vector<char> covered(xs.size(), 0);
        for (;;) {
            size_t seed = xs.size();
            for (size_t g = 0; g < xs.size(); ++g)
                if (!done[g] && lean_ok[g] && !covered[g]) { seed = g; break; }
            if (seed == xs.size()) break;
            uint64_t need = xs[seed] & lfull;
            const int lb_pass = fit_low_bits(need, lb, tbits_max, nl, false);
            if (lb_pass < 0) { lean_ok[seed] = 0; continue; }
            std::vector<size_t> open;
            for (size_t g = 0; g < xs.size(); ++g)
                if (!done[g] && lean_ok[g] && !covered[g]) open.push_back(g);
            need = cover_greedy(xs, open, need, (1ull << lb_pass) - 1ull, tb_local, nl, lfull);
            LeanPlan lp;
            lp.tp = make_plan(nl, need, tbits_max, lb_pass, 0);
            lp.swz = env_int("VQE_SWIZZLE", 1) != 0 && (host_only ? plan_tma(lp.tp, true).ok : tma_available(lp.tp, true));
            // a group that cannot be lowered even into an empty pass takes the general path
            size_t n_cov = 0;
            for (size_t g : open) {
                if ((xs[g] & lfull & ~lp.tp.tile_mask) != 0) continue;
                PSPass scratch;
                scratch.lean = true;
                scratch.tp = lp.tp;
                scratch.swz = lp.swz;
                if (!lower_lean_group(scratch, xs[g], grp[g])) { lean_ok[g] = 0; continue; }
                covered[g] = 1;
                ++n_cov;
            }
            if (n_cov) lplans.push_back(lp);
        }
        // (B) every group goes to ONE of the passes whose tile bits cover it.  A pass costs max(HBM sweep, its entries), so
        // the groups are spread for equal entry counts (most constrained groups first, each to its least loaded candidate)
        // instead of piling up in the first pass that covers them: the heavy passes were 5-10 x the sweep time, the light
        // ones idle at the HBM floor.  VQE_EXP_BALANCE=0 restores first-fit.
        const bool balance = env_int("VQE_EXP_BALANCE", 1) != 0;
        std::vector<std::vector<size_t>> members(lplans.size());
        {
            struct Cand {
                size_t g;
                std::vector<int> passes;
                double cost;
            };
            std::vector<Cand> cands;
            for (size_t g = 0; g < xs.size(); ++g) {
                if (done[g] || !lean_ok[g] || !covered[g]) continue;
                Cand c;
                c.g = g;
                for (size_t i = 0; i < lplans.size(); ++i)
                    if ((xs[g] & lfull & ~lplans[i].tp.tile_mask) == 0) c.passes.push_back((int)i);
                // entries = active occupation patterns x 256-pair chunks; estimated from the pattern count bound
                PSPass scratch;
                scratch.lean = true;
                scratch.tp = lplans[c.passes[0]].tp;
                scratch.swz = lplans[c.passes[0]].swz;
                lower_lean_group(scratch, xs[g], grp[g]);
                double cost = 0.0;
                for (const DevFlat2& fl : scratch.flats2) cost += fl.tab == 0xffffu ? 1.0 : 1.5;  // additive entries read two tables
                c.cost = cost;
                cands.push_back(std::move(c));
            }
            std::vector<double> load(lplans.size(), 0.0);
            if (balance)
                std::stable_sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) {
                    if (a.passes.size() != b.passes.size()) return a.passes.size() < b.passes.size();
                    return a.cost > b.cost;
                });
            for (const Cand& c : cands) {
                int best = c.passes[0];
                if (balance)
                    for (int i : c.passes)
                        if (load[i] < load[best]) best = i;
                load[best] += c.cost;
                members[best].push_back(c.g);
            }
        }
        // (C) lower the groups of every pass (ascending group order: deterministic tables); a group that exceeds a pass's
        // table capacities moves on to its next candidate, and to the general path when none takes it
        std::vector<PSPass> lpass(lplans.size());
        for (size_t i = 0; i < lplans.size(); ++i) {
            lpass[i].lean = true;
            lpass[i].tp = lplans[i].tp;
            lpass[i].swz = lplans[i].swz;
            std::sort(members[i].begin(), members[i].end());
            // the real-layout twin with its lane table (see PSPass::rlp), when the pass has a real-layout tensor-map form
            if (env_int("VQE_EXP_LANE_TAB", 1) != 0 && env_int("VQE_EXP_RL2", 1) != 0 && !lplans[i].tp.vbit && lplans[i].tp.tbits >= 9) {
                const bool sw_ok = env_int("VQE_SWIZZLE", 1) != 0 &&
                                   (host_only ? plan_tma(lplans[i].tp, true, true).ok : tma_available(lplans[i].tp, true, true));
                const bool nat_ok = host_only ? plan_tma(lplans[i].tp, false, true).ok : tma_available(lplans[i].tp, false, true);
                if (sw_ok || nat_ok) {
                    lpass[i].rlp = std::make_shared<PSPass>();
                    lpass[i].rlp->lean = true;
                    lpass[i].rlp->tp = lplans[i].tp;
                    lpass[i].rlp->swz = sw_ok;
                }
            }
        }
        // lowers a group into a pass and, in the same order, into its real-layout twin (the twin is dropped if it ever refuses)
        auto lower_both = [&](PSPass& ps2, size_t g) -> bool {
            if (!lower_lean_group(ps2, xs[g], grp[g])) return false;
            if (ps2.rlp && !lower_lean_group(*ps2.rlp, xs[g], grp[g], true)) ps2.rlp.reset();
            return true;
        };
        std::vector<size_t> spill;
        for (size_t i = 0; i < lplans.size(); ++i)
            for (size_t g : members[i]) {
                if (lower_both(lpass[i], g)) done[g] = 1;
                else spill.push_back(g);
            }
        for (size_t g : spill)
            for (size_t i = 0; i < lplans.size() && !done[g]; ++i)
                if ((xs[g] & lfull & ~lplans[i].tp.tile_mask) == 0 && lower_both(lpass[i], g)) done[g] = 1;
        for (size_t i = 0; i < lplans.size(); ++i)
            if (!lpass[i].flats2.empty()) {
                batch_lean_entries(lpass[i]);
                ps->passes.push_back(std::move(lpass[i]));
            }
        remaining = 0;
        for (size_t g = 0; g < xs.size(); ++g) remaining += done[g] ? 0 : 1;
        (void)n_lean;
    }
    while (remaining) {
        // Choose the tile bits of this pass greedily for COVERAGE (the Pauli sum is uploaded once and evaluated
        // thousands of times, so every pass saved is a full sweep over the state saved per evaluation):
        // seed with the first open group, then repeatedly add the bit that completes the most open groups; when no
        // single bit completes one, take the open group that needs the fewest new bits.  Only groups that share the
        // seed's global X pattern can join (0 = local pass, m = peer pass between ranks r and r ^ m).
        uint64_t need = 0, pat = 0;
        size_t seed = xs.size();
        for (size_t g = 0; g < xs.size(); ++g)
            if (!done[g]) { seed = g; break; }
        pat = xs[seed] >> nl;
        need = xs[seed] & lfull;
        const int lb_pass = fit_low_bits(need, lb, tbits_max, nl, pat != 0);  // smaller floor for wide seeds
        if (lb_pass < 0)
            return fail(VQE_ERR_INVALID, "Pauli term with %d local X/Y letters exceeds the %d-bit tile",
                        popc64(need), std::min(tbits_max, nl));
        const uint64_t lowmask_pass = (1ull << lb_pass) - 1ull;
        std::vector<size_t> open;  // open groups of this pattern
        for (size_t g = 0; g < xs.size(); ++g)
            if (!done[g] && (xs[g] >> nl) == pat) open.push_back(g);
        const int cap_bits = std::min(tbits_max - (pat ? 1 : 0), nl);
        for (;;) {
            const uint64_t have = need | lowmask_pass;
            if (popc64(have) >= cap_bits) break;
            // gain of every single bit: open groups that become fully covered
            int best_bit = -1;
            size_t best_gain = 0;
            for (int b2 = 0; b2 < nl; ++b2) {
                if ((have >> b2) & 1ull) continue;
                const uint64_t with = have | (1ull << b2);
                size_t gain = 0;
                for (size_t g : open) {
                    const uint64_t xl = xs[g] & lfull;
                    if ((xl & ~with) == 0 && (xl & ~have) != 0) ++gain;
                }
                if (gain > best_gain) { best_gain = gain; best_bit = b2; }
            }
            if (best_bit >= 0) {
                need |= 1ull << best_bit;
                continue;
            }
            // no single bit completes a group: the open group with the fewest missing bits that still fits
            size_t best_g = xs.size();
            int best_missing = 1 << 30;
            for (size_t g : open) {
                const uint64_t xl = xs[g] & lfull;
                const int missing = popc64(xl & ~have);
                if (missing == 0 || popc64(have | xl) > cap_bits) continue;
                if (missing < best_missing) { best_missing = missing; best_g = g; }
            }
            if (best_g == xs.size()) break;
            need |= xs[best_g] & lfull;
        }
        PSPass p;
        p.tp = make_plan(nl, need, tbits_max, lb_pass, pat);
        std::vector<size_t> members;
        size_t n_terms_pass = 0;
        // everything inside the final tile mask (within the per-pass table capacity); the seed always goes first
        for (size_t g : open) {
            if ((xs[g] & lfull & ~p.tp.tile_mask) != 0) continue;
            if (!members.empty() && (n_terms_pass + grp[g].size() > TERM_CAP || members.size() >= GROUP_CAP)) continue;
            members.push_back(g);
            n_terms_pass += grp[g].size();
            done[g] = 1;
        }
        remaining -= members.size();
        std::vector<size_t> pending_groups;
        for (size_t g : members) {
            DevGroup dg;
            memset(&dg, 0, sizeof dg);
            dg.lx = plan_lx(xs[g], p.tp);
            dg.hb = dg.lx ? 31 - __builtin_clz(dg.lx) : 0;
            dg.t_begin = (uint32_t)p.terms_expect.size();
            struct Sorted {
                int key;  // parity * 8 + sign class
                DevTerm e, a;
            };
            std::vector<Sorted> sorted;
            for (int parity = 0; parity < 2; ++parity) {
                for (const HTerm& t : grp[g]) {
                    if ((t.ny & 1) != parity) continue;
                    DevTerm e, a;
                    memset(&e, 0, sizeof e);
                    e.lz = plan_lz(t.z, p.tp);
                    e.zout = plan_zout(t.z, p.tp);
                    a = e;
                    // apply weight: c * i^ny
                    a.ar = t.cr;
                    a.ai = t.ci;
                    mul_i_pow(a.ar, a.ai, t.ny);
                    // expectation pair weight: even ny: c * i^ny (multiplies 2Re w)
                    //                          odd  ny: c * i^ny * i = c * i^(ny+1)  (multiplies 2Im w)
                    e.ar = t.cr;
                    e.ai = t.ci;
                    mul_i_pow(e.ar, e.ai, parity ? t.ny + 1 : t.ny);
                    if (xs[g] == 0) { e.ar = t.cr; e.ai = t.ci; }
                    if (e.ai != 0.0) p.cplx = true;
                    {   // per-pair parity bits (see k_tile_expect)
                        const uint64_t tsz = 1ull << p.tp.tbits;
                        const uint64_t count = (xs[g] == 0) ? tsz : tsz / 2;
                        const int thr = (int)std::min<uint64_t>(threads_cfg, std::max<uint64_t>(32, tsz / 2));
                        const int tshift = 31 - __builtin_clz((unsigned)thr);
                        const int n_j = (int)std::max<uint64_t>(1, count / thr);
                        uint32_t jm = 0;
                        for (int jj = 0; jj < n_j && jj < 32; ++jj) {
                            uint32_t pj = (uint32_t)jj << tshift;
                            uint32_t uj = (xs[g] == 0) ? pj : (((pj >> dg.hb) << (dg.hb + 1)) | (pj & ((1u << dg.hb) - 1u)));
                            if (__builtin_popcount(uj & e.lz) & 1) jm |= 1u << jj;
                        }
                        e.jmask = jm;
                    }
                    // sign class = Z letters on the tile bits that enumerate a thread's pairs (k_tile_expect)
                    const int cls = (xs[g] == 0) ? (int)(((e.jmask >> 1) & 1u) | (((e.jmask >> 2) & 1u) << 1) | (((e.jmask >> 4) & 1u) << 2))
                                                 : (int)(((e.jmask >> 1) & 1u) | (((e.jmask >> 2) & 1u) << 1));
                    sorted.push_back({parity * 8 + cls, e, a});
                    if (parity) dg.n_odd++; else dg.n_even++;
                }
            }
            std::stable_sort(sorted.begin(), sorted.end(), [](const Sorted& u, const Sorted& v) { return u.key < v.key; });
            for (const Sorted& sd : sorted) {
                p.terms_expect.push_back(sd.e);
                p.terms_apply.push_back(sd.a);
                const int parity = sd.key >> 3, cls = sd.key & 7;
                dg.cnt[(xs[g] == 0) ? cls : parity * 4 + cls]++;
            }
            p.groups.push_back(dg);
            pending_groups.push_back(p.groups.size() - 1);
        }
        // collapse groups whose Z-variants differ on few tile bits (needs real weights: decided once p.cplx is known)
        for (size_t gi2 : pending_groups) {
            DevGroup& dg = p.groups[gi2];
            const uint32_t nt = dg.n_even + dg.n_odd;
            if (p.cplx || dg.lx == 0 || nt < 2 || p.gcols.size() >= GROUP_CAP) continue;
            const DevTerm* te = p.terms_expect.data() + dg.t_begin;
            uint32_t D = 0;
            bool same_out = true;
            for (uint32_t q = 0; q < nt; ++q) {
                D |= te[q].lz ^ te[0].lz;
                same_out = same_out && te[q].zout == te[0].zout;
            }
            const uint32_t E = D | (1u << dg.hb);
            const int ne = __builtin_popcount(E);
            if (!same_out || ne > 6) continue;
            std::vector<uint32_t> epos;
            for (int b2 = 0; b2 < p.tp.tbits; ++b2)
                if ((E >> b2) & 1u) epos.push_back((uint32_t)b2);
            double scale = 0.0;
            for (uint32_t q = 0; q < nt; ++q) scale += fabs(te[q].ar);
            std::vector<DevGColEntry> ent;
            for (uint32_t pi = 0; pi < (1u << ne); ++pi) {
                uint32_t pat = 0;
                for (int b2 = 0; b2 < ne; ++b2)
                    if ((pi >> b2) & 1u) pat |= 1u << epos[b2];
                if ((pat >> dg.hb) & 1u) continue;  // a-side only
                double fr = 0.0, fi = 0.0;
                for (uint32_t q = 0; q < nt; ++q) {
                    const double sg = (__builtin_popcount(pat & (te[q].lz ^ te[0].lz)) & 1) ? -1.0 : 1.0;
                    if (q < dg.n_even) fr += sg * te[q].ar;
                    else fi += sg * te[q].ar;
                }
                if (fabs(fr) <= 1e-15 * scale) fr = 0.0;
                if (fabs(fi) <= 1e-15 * scale) fi = 0.0;
                if (fr == 0.0 && fi == 0.0) continue;
                DevGColEntry en;
                memset(&en, 0, sizeof en);
                en.fr = fr;
                en.fi = fi;
                en.pat = pat;
                ent.push_back(en);
            }
            // cost model (instructions per thread): ~35 per active pair vs ~40 + 2 per term for every 4 pairs
            const double cost_col = (double)ent.size() * (double)(1u << (p.tp.tbits - ne)) * 35.0;
            const double cost_cls = (double)(1u << (p.tp.tbits - 1)) / 4.0 * (150.0 + 10.0 * nt);
            if (cost_col >= cost_cls) continue;
            const int free_log = p.tp.tbits - ne;
            bool any_fi = false;
            for (const DevGColEntry& en : ent) any_fi = any_fi || en.fi != 0.0;
            if (!any_fi && ne <= 4 && free_log >= 8 && p.flats.size() + (ent.size() << (free_log - 8)) <= FLAT_CAP) {
                // flat work list: one entry per 256 free indices of every active pattern
                uint32_t zsel = 0;
                while (zsel < p.fzout.size() && p.fzout[zsel] != te[0].zout) ++zsel;
                if (zsel == p.fzout.size()) p.fzout.push_back(te[0].zout);
                for (const DevGColEntry& en : ent)
                    for (uint32_t hi = 0; hi < (1u << (free_log - 8)); ++hi) {
                        DevFlat fl;
                        memset(&fl, 0, sizeof fl);
                        fl.fr = 2.0 * en.fr;
                        fl.zsel = zsel;
                        fl.lx = dg.lx;
                        fl.lz = te[0].lz;
                        uint32_t fhi = hi << 8;  // deposit the high free-index bits now
                        uint32_t d6 = 1u << 6, d7 = 1u << 7;
                        for (int b2 = 0; b2 < ne; ++b2) {
                            const uint32_t hm = ~((1u << epos[b2]) - 1u);
                            fl.himask[b2] = (uint16_t)(hm & 0xffffu);
                            fhi += fhi & hm;
                            d6 += d6 & hm;
                            d7 += d7 & hm;
                        }
                        fl.zsel = zsel | ((uint32_t)__builtin_ctz(d6) << 16) | ((uint32_t)__builtin_ctz(d7) << 24);
                        fl.pat = en.pat | fhi;
                        p.flats.push_back(fl);
                    }
                dg.pad = 0xffffffffu;
                continue;
            }
            if (p.gents.size() + ent.size() > GCOL_ENT_CAP) continue;
            DevGCol co;
            memset(&co, 0, sizeof co);
            co.zout = te[0].zout;
            co.lz = te[0].lz;
            co.n_active = (uint32_t)ent.size();
            co.nd = (uint32_t)ne;
            for (int b2 = 0; b2 < ne; ++b2) co.dpos[b2] = epos[b2];
            co.ent_begin = (uint32_t)p.gents.size();
            co.free_log = (uint32_t)(p.tp.tbits - ne);
            p.gcols.push_back(co);
            p.gents.insert(p.gents.end(), ent.begin(), ent.end());
            dg.pad = (uint32_t)p.gcols.size();  // index + 1
        }
        // apply (sigma = O psi): flat entries of the groups whose Z-variants differ on <= 4 tile bits
        {
            p.aoff.push_back(0);
            for (size_t gi2 = 0; gi2 < p.groups.size(); ++gi2) {
                const DevGroup& dg = p.groups[gi2];
                const uint32_t nt = dg.n_even + dg.n_odd;
                const DevTerm* ta = p.terms_apply.data() + dg.t_begin;
                uint32_t D = 0;
                bool same_out = nt > 0;
                for (uint32_t q = 0; q < nt; ++q) {
                    D |= ta[q].lz ^ ta[0].lz;
                    same_out = same_out && ta[q].zout == ta[0].zout;
                }
                const int nd = __builtin_popcount(D);
                const int free_log = p.tp.tbits - nd;
                if (same_out && nd <= 4 && free_log >= 8) {
                    std::vector<uint32_t> dpos;
                    for (int b2 = 0; b2 < p.tp.tbits; ++b2)
                        if ((D >> b2) & 1u) dpos.push_back((uint32_t)b2);
                    double scale = 0.0;
                    for (uint32_t q = 0; q < nt; ++q) scale += fabs(ta[q].ar) + fabs(ta[q].ai);
                    std::vector<DevAFlat> ent;
                    uint32_t zsel = 0;
                    while (zsel < p.azout.size() && p.azout[zsel] != ta[0].zout) ++zsel;
                    for (uint32_t pi = 0; pi < (1u << nd); ++pi) {
                        uint32_t pat = 0;
                        for (int b2 = 0; b2 < nd; ++b2)
                            if ((pi >> b2) & 1u) pat |= 1u << dpos[b2];
                        double fr = 0.0, fi = 0.0;
                        for (uint32_t q = 0; q < nt; ++q) {
                            const double sg = (__builtin_popcount(pat & (ta[q].lz ^ ta[0].lz)) & 1) ? -1.0 : 1.0;
                            fr += sg * ta[q].ar;
                            fi += sg * ta[q].ai;
                        }
                        if (fabs(fr) <= 1e-15 * scale) fr = 0.0;
                        if (fabs(fi) <= 1e-15 * scale) fi = 0.0;
                        if (fr == 0.0 && fi == 0.0) continue;
                        for (uint32_t hi = 0; hi < (1u << (free_log - 8)); ++hi) {
                            DevAFlat fl;
                            memset(&fl, 0, sizeof fl);
                            fl.fr = fr;
                            fl.fi = fi;
                            fl.lx = dg.lx;
                            fl.lz = ta[0].lz;
                            fl.zsel = zsel;
                            uint32_t fhi = hi << 8;
                            for (int b2 = 0; b2 < nd; ++b2) {
                                fl.himask[b2] = (uint16_t)(~((1u << dpos[b2]) - 1u) & 0xffffu);
                                fhi += fhi & ~((1u << dpos[b2]) - 1u);
                            }
                            fl.pat = pat | fhi;
                            ent.push_back(fl);
                        }
                    }
                    // an all-zero group contributes nothing; keep one zero entry so that it is not sent to the term loop
                    if (ent.empty()) {
                        DevAFlat fl;
                        memset(&fl, 0, sizeof fl);
                        fl.lx = dg.lx;
                        fl.zsel = zsel;
                        ent.push_back(fl);
                    }
                    if (p.aflat.size() + ent.size() <= AFLAT_CAP) {
                        if (zsel == p.azout.size()) p.azout.push_back(ta[0].zout);
                        p.aflat.insert(p.aflat.end(), ent.begin(), ent.end());
                    }
                }
                p.aoff.push_back((uint32_t)p.aflat.size());
            }
        }
        ps->passes.push_back(std::move(p));
    }
    if (ps->diag2.on) {   // the quadratic form replaces the X-mask-0 groups only if they sit alone in their passes
        for (PSPass& p : ps->passes) {
            if (p.lean) continue;
            size_t n0 = 0;
            for (const DevGroup& gr : p.groups) n0 += (gr.lx == 0 && !p.tp.vbit) ? 1 : 0;
            p.diag_only = n0 == p.groups.size() && n0 > 0;
            if (n0 != 0 && !p.diag_only) ps->diag2.on = false;
        }
        if (!ps->diag2.on)
            for (PSPass& p : ps->passes) p.diag_only = false;
    }
    if (getenv("VQE_DEBUG_PLAN")) {
        size_t ng = 0, nc = 0, ne = 0, nt = 0, nfl = 0;
        for (const PSPass& p : ps->passes) {
            ng += p.groups.size();
            nc += p.gcols.size();
            nfl += p.flats.size();
            ne += p.gents.size();
            nt += p.terms_expect.size();
        }
        fprintf(stderr, "[paulisum] passes %zu groups %zu terms %zu; collapsed groups (per-group path) %zu with %zu active patterns; flat entries %zu\n",
                ps->passes.size(), ng, nt, nc, ne, nfl);
    }
    return VQE_OK;
}

static void free_paulisum_device(vqe_paulisum* ps) {
    if (ps->diag2.d_a) cudaFree(ps->diag2.d_a);
    if (ps->diag2.d_b) cudaFree(ps->diag2.d_b);
    if (ps->diag2.d_t) cudaFree(ps->diag2.d_t);
    ps->diag2.d_a = ps->diag2.d_b = ps->diag2.d_t = nullptr;
    for (PSPass& p : ps->passes) {
        if (p.d_groups) cudaFree(p.d_groups);
        if (p.d_terms_expect) cudaFree(p.d_terms_expect);
        if (p.d_terms_apply) cudaFree(p.d_terms_apply);
        if (p.d_scat) cudaFree(p.d_scat);
        if (p.d_flats) cudaFree(p.d_flats);
        if (p.d_fzout) cudaFree(p.d_fzout);
        if (p.d_aflat) cudaFree(p.d_aflat);
        if (p.d_aoff) cudaFree(p.d_aoff);
        if (p.d_azout) cudaFree(p.d_azout);
        p.d_aflat = nullptr;
        p.d_aoff = nullptr;
        p.d_azout = nullptr;
        p.d_flats = nullptr;
        p.d_fzout = nullptr;
        if (p.d_flats2) cudaFree(p.d_flats2);
        if (p.d_flats2_rl) cudaFree(p.d_flats2_rl);
        p.d_flats2_rl = nullptr;
        if (p.d_lane_rl) cudaFree(p.d_lane_rl);
        if (p.d_addtab_rl) cudaFree(p.d_addtab_rl);
        if (p.d_addpat_rl) cudaFree(p.d_addpat_rl);
        if (p.d_addout_rl) cudaFree(p.d_addout_rl);
        if (p.d_fzout_rl) cudaFree(p.d_fzout_rl);
        p.d_lane_rl = nullptr;
        p.d_addtab_rl = nullptr;
        p.d_addpat_rl = nullptr;
        p.d_addout_rl = nullptr;
        p.d_fzout_rl = nullptr;
        if (p.d_goff) cudaFree(p.d_goff);
        if (p.d_addtab) cudaFree(p.d_addtab);
        if (p.d_addpat) cudaFree(p.d_addpat);
        if (p.d_addout) cudaFree(p.d_addout);
        p.d_flats2 = nullptr;
        p.d_goff = nullptr;
        p.d_addtab = nullptr;
        p.d_addpat = nullptr;
        p.d_addout = nullptr;
        if (p.d_gcols) cudaFree(p.d_gcols);
        if (p.d_gents) cudaFree(p.d_gents);
        p.d_gcols = nullptr;
        p.d_gents = nullptr;
        p.d_groups = nullptr;
        p.d_terms_expect = p.d_terms_apply = nullptr;
        p.d_scat = nullptr;
    }
}

template <class T>
static int upload_vec(T** dptr, const std::vector<T>& v) {
    CK(cudaMalloc((void**)dptr, std::max<size_t>(1, v.size()) * sizeof(T)));
    if (!v.empty()) CK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return VQE_OK;
}
static int upload_paulisum(vqe_ctx* c, vqe_paulisum* ps) {
    if (ps->diag2.on) {
        int rc = upload_vec(&ps->diag2.d_a, ps->diag2.a);
        if (rc == VQE_OK) rc = upload_vec(&ps->diag2.d_b, ps->diag2.b);
        if (rc == VQE_OK) rc = upload_vec(&ps->diag2.d_t, ps->diag2.t);
        if (rc) return rc;
    }
    for (PSPass& p : ps->passes) {
        if (p.lean) {
            int rc = upload_vec(&p.d_flats2, p.flats2);
            // real-layout twin of the entries (unsharded contexts keep a purely real state as n_amp doubles)
            p.rl_ok = false;
            p.rl_lane = false;
            if (rc == VQE_OK && c->real_layout_ok && !p.tp.vbit && p.tp.tbits >= 9) {
                p.rl_swz = env_int("VQE_SWIZZLE", 1) != 0 && tma_available(p.tp, true, true);
                p.rl_ok = p.rl_swz || tma_available(p.tp, false, true);
                if (p.rl_ok && p.rlp && p.rlp->swz == p.rl_swz && p.rlp->flats2.size() == p.flats2.size() &&
                    p.rlp->lane_rl.size() == 32 * p.flats2.size()) {
                    // the twin lowered for the real layout (lane table, conflict-free lane order)
                    rc = upload_vec(&p.d_flats2_rl, p.rlp->flats2);
                    if (rc == VQE_OK) rc = upload_vec(&p.d_lane_rl, p.rlp->lane_rl);
                    if (rc == VQE_OK) rc = upload_vec(&p.d_addtab_rl, p.rlp->addtab);
                    if (rc == VQE_OK) rc = upload_vec(&p.d_addpat_rl, p.rlp->addpat);
                    if (rc == VQE_OK) rc = upload_vec(&p.d_addout_rl, p.rlp->addout);
                    if (rc == VQE_OK) rc = upload_vec(&p.d_fzout_rl, p.rlp->fzout);
                    p.rl_lane = rc == VQE_OK;
                } else if (p.rl_ok) {
                    std::vector<DevFlat2> rl(p.flats2.size());
                    for (size_t k = 0; k < rl.size(); ++k) rl[k] = lean_entry_to_rl(p.flats2[k], p.swz, p.rl_swz);
                    rc = upload_vec(&p.d_flats2_rl, rl);
                }
            }
            if (rc == VQE_OK) rc = upload_vec(&p.d_goff, p.boff);
            if (rc == VQE_OK) rc = upload_vec(&p.d_addtab, p.addtab);
            if (rc == VQE_OK) rc = upload_vec(&p.d_addpat, p.addpat);
            if (rc == VQE_OK) rc = upload_vec(&p.d_addout, p.addout);
            if (rc == VQE_OK) rc = upload_vec(&p.d_fzout, p.fzout);
            if (rc == VQE_OK) rc = upload_vec(&p.d_scat, p.tp.scat);
            if (rc) return rc;
            continue;
        }
        CK(cudaMalloc((void**)&p.d_groups, p.groups.size() * sizeof(DevGroup)));
        CK(cudaMalloc((void**)&p.d_terms_expect, std::max<size_t>(1, p.terms_expect.size()) * sizeof(DevTerm)));
        CK(cudaMalloc((void**)&p.d_terms_apply, std::max<size_t>(1, p.terms_apply.size()) * sizeof(DevTerm)));
        CK(cudaMalloc((void**)&p.d_scat, p.tp.scat.size() * sizeof(uint64_t)));
        CK(cudaMemcpy(p.d_groups, p.groups.data(), p.groups.size() * sizeof(DevGroup), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(p.d_terms_expect, p.terms_expect.data(), p.terms_expect.size() * sizeof(DevTerm), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(p.d_terms_apply, p.terms_apply.data(), p.terms_apply.size() * sizeof(DevTerm), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(p.d_scat, p.tp.scat.data(), p.tp.scat.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
        CK(cudaMalloc((void**)&p.d_gcols, std::max<size_t>(1, p.gcols.size()) * sizeof(DevGCol)));
        CK(cudaMalloc((void**)&p.d_gents, std::max<size_t>(1, p.gents.size()) * sizeof(DevGColEntry)));
        CK(cudaMemcpy(p.d_gcols, p.gcols.data(), p.gcols.size() * sizeof(DevGCol), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(p.d_gents, p.gents.data(), p.gents.size() * sizeof(DevGColEntry), cudaMemcpyHostToDevice));
        CK(cudaMalloc((void**)&p.d_flats, std::max<size_t>(1, p.flats.size()) * sizeof(DevFlat)));
        CK(cudaMemcpy(p.d_flats, p.flats.data(), p.flats.size() * sizeof(DevFlat), cudaMemcpyHostToDevice));
        CK(cudaMalloc((void**)&p.d_aflat, std::max<size_t>(1, p.aflat.size()) * sizeof(DevAFlat)));
        CK(cudaMalloc((void**)&p.d_aoff, std::max<size_t>(1, p.aoff.size()) * sizeof(uint32_t)));
        CK(cudaMalloc((void**)&p.d_azout, std::max<size_t>(1, p.azout.size()) * sizeof(uint64_t)));
        CK(cudaMemcpy(p.d_aflat, p.aflat.data(), p.aflat.size() * sizeof(DevAFlat), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(p.d_aoff, p.aoff.data(), p.aoff.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(p.d_azout, p.azout.data(), p.azout.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
        CK(cudaMalloc((void**)&p.d_fzout, std::max<size_t>(1, p.fzout.size()) * sizeof(uint64_t)));
        CK(cudaMemcpy(p.d_fzout, p.fzout.data(), p.fzout.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    }
    return VQE_OK;
}

static int collect_terms(const vqe_ctx* c, int n_terms, const uint64_t* x, const uint64_t* z, const int32_t* ny,
                         const double* cre, const double* cim, std::vector<HTerm>& out) {
    if (n_terms < 0 || (n_terms > 0 && (!x || !z || !ny || !cre))) return fail(VQE_ERR_INVALID, "null array");
    const uint64_t full = (1ull << c->n) - 1ull;
    out.reserve(n_terms);
    for (int k = 0; k < n_terms; ++k) {
        if ((x[k] | z[k]) & ~full) return fail(VQE_ERR_INVALID, "term %d: mask has bits >= n_qubits", k);
        if (popc64(x[k] & z[k]) != ny[k]) return fail(VQE_ERR_INVALID, "term %d: ny != popcount(x&z)", k);
        HTerm t;
        t.x = x[k];
        t.z = z[k];
        t.ny = ny[k];
        t.cr = cre[k];
        t.ci = cim ? cim[k] : 0.0;
        if (t.cr == 0.0 && t.ci == 0.0) continue;
        out.push_back(t);
    }
    return VQE_OK;
}

extern "C" int vqe_paulisum_create(vqe_ctx* c, vqe_paulisum** out, int n_terms, const uint64_t* x,
                                   const uint64_t* z, const int32_t* ny, const double* cre, const double* cim) {
    if (!c || !out) return fail(VQE_ERR_INVALID, "null argument");
    *out = nullptr;
    CK(cudaSetDevice(c->device));
    std::vector<HTerm> terms;
    int rc = collect_terms(c, n_terms, x, z, ny, cre, cim, terms);
    if (rc) return rc;
    vqe_paulisum* ps = new vqe_paulisum();
    ps->device = c->device;
    if (c->world > 1) ps->terms = terms;  // kept for relabelled twins (sharded states only)
    ps->tile_bits = c->tile_bits;
    ps->low_bits = c->low_bits;
    ps->threads = c->threads;
    rc = build_paulisum(ps, c->n, c->nl, c->tile_bits, c->low_bits, c->threads, std::move(terms));
    if (rc == VQE_OK) rc = upload_paulisum(c, ps);
    if (rc) {
        free_paulisum_device(ps);
        delete ps;
        return rc;
    }
    *out = ps;
    return VQE_OK;
}
// Host-only view of the Pauli-sum planner (no CUDA call): how a Pauli sum is grouped by X-mask and packed into tile
// passes for a state of n_qubits with n_global rank bits.  Used by the CPU tests of the host logic and by the
// profiling notes (per-pass statistics on stderr with VQE_DEBUG_PLAN >= 3).
extern "C" int vqe_plan_paulisum(int n_qubits, int n_global, int tile_bits, int low_bits, int n_terms, const uint64_t* x,
                                 const uint64_t* z, const int32_t* ny, const double* cre, const double* cim,
                                 int32_t* n_groups, int32_t* n_passes, int cap, int32_t* pass_groups, int32_t* pass_terms,
                                 uint64_t* pass_tile_mask) {
    if (n_qubits < 1 || n_qubits > 40 || n_global < 0 || n_global > 6 || n_global >= n_qubits)
        return fail(VQE_ERR_INVALID, "bad qubit counts");
    if (tile_bits < 6 || tile_bits > 12) tile_bits = 12;
    if (low_bits < 0 || low_bits > tile_bits) low_bits = 5;
    vqe_ctx fake;
    fake.n = n_qubits;
    std::vector<HTerm> terms;
    int rc = collect_terms(&fake, n_terms, x, z, ny, cre, cim, terms);
    if (rc) return rc;
    vqe_paulisum ps;
    rc = build_paulisum(&ps, n_qubits, n_qubits - n_global, tile_bits, low_bits, 512, std::move(terms), true);
    if (rc) return rc;
    if (n_groups) *n_groups = ps.n_groups;
    if (n_passes) *n_passes = (int32_t)ps.passes.size();
    for (size_t p = 0; p < ps.passes.size(); ++p) {
        const PSPass& pp = ps.passes[p];
        if ((int)p < cap) {
            if (pass_groups) pass_groups[p] = pp.lean ? (int32_t)pp.goff.size() - 1 : (int32_t)pp.groups.size();
            if (pass_terms) pass_terms[p] = pp.lean ? (int32_t)pp.lean_terms : (int32_t)pp.terms_expect.size();
            if (pass_tile_mask) pass_tile_mask[p] = pp.tp.tile_mask;
        }
        if (getenv("VQE_DEBUG_PLAN") && atoi(getenv("VQE_DEBUG_PLAN")) >= 3 && pp.lean) {
            size_t n_add = 0;
            for (const DevFlat2& fl : pp.flats2) n_add += fl.tab != 0xffffu;
            fprintf(stderr, "[pspass] %zu LEAN groups %zu terms %zu entries %zu (additive %zu) batches %zu addpat %zu addout %zu tab %zu lbits %d tbits %d\n", p,
                    pp.goff.size() - 1, pp.lean_terms, pp.flats2.size(), n_add, pp.boff.size() - 1, pp.addpat.size(), pp.addout.size(), pp.addtab.size(),
                    pp.tp.lbits, pp.tp.tbits);
        } else if (getenv("VQE_DEBUG_PLAN") && atoi(getenv("VQE_DEBUG_PLAN")) >= 3) {
            size_t n_flatg = 0, n_colg = 0, n_clsg = 0, cls_terms = 0, diag_terms = 0, col_items = 0, aflat_groups = 0, aterm = 0;
            for (size_t gi = 0; gi < pp.groups.size(); ++gi) {
                const DevGroup& dg = pp.groups[gi];
                if (dg.pad == 0xffffffffu) ++n_flatg;
                else if (dg.pad) {
                    ++n_colg;
                    const DevGCol& co = pp.gcols[dg.pad - 1];
                    col_items += (size_t)co.n_active << co.free_log;
                } else if (dg.lx == 0) diag_terms += dg.n_even + dg.n_odd;
                else {
                    ++n_clsg;
                    cls_terms += dg.n_even + dg.n_odd;
                }
                if (pp.aoff[gi + 1] > pp.aoff[gi]) ++aflat_groups;
                else aterm += dg.n_even + dg.n_odd;
            }
            fprintf(stderr, "[pspass] %zu groups %zu terms %zu flatg %zu flats %zu colg %zu colitems %zu clsg %zu clsterms %zu diagterms %zu aflatg %zu aflats %zu aterms %zu lbits %d tbits %d vbit %d\n",
                    p, pp.groups.size(), pp.terms_expect.size(), n_flatg, pp.flats.size(), n_colg, col_items, n_clsg, cls_terms,
                    diag_terms, aflat_groups, pp.aflat.size(), aterm, pp.tp.lbits, pp.tp.tbits, (int)pp.tp.vbit);
        }
    }
    return VQE_OK;
}
// Host-only interpreter of the LEAN passes of a Pauli sum (no CUDA call): evaluates  sum over the lean groups of
// <psi|O_g|psi>  and, when sigma is given, accumulates  sigma += O_lean psi,  by walking the same entry tables with the
// same decode routine (lean_decode) the kernels k_expect_lean / k_apply_lean use.  Lets the CPU test-suite check the table
// construction (patterns, deposits, signs, additive tables, outside-tile constants) against the oracle without a GPU.
// psi / sigma: 2^(n_qubits - n_global) interleaved complex amplitudes of rank `rank`.
extern "C" int vqe_debug_lean_host(int n_qubits, int n_global, int rank, int tile_bits, int low_bits, int n_terms, const uint64_t* x,
                                   const uint64_t* z, const int32_t* ny, const double* cre, const double* cim,
                                   const double* psi_re_im, double* out_re, int32_t* n_lean_terms, int32_t* n_fat_terms,
                                   double* sigma_re_im) {
    if (n_qubits < 1 || n_qubits > 30 || n_global < 0 || n_global >= n_qubits) return fail(VQE_ERR_INVALID, "bad qubit counts");
    if (tile_bits < 6 || tile_bits > 12) tile_bits = 12;
    if (low_bits < 0 || low_bits > tile_bits) low_bits = 5;
    if (!psi_re_im || !out_re) return fail(VQE_ERR_INVALID, "null argument");
    vqe_ctx fake;
    fake.n = n_qubits;
    std::vector<HTerm> terms;
    int rc = collect_terms(&fake, n_terms, x, z, ny, cre, cim, terms);
    if (rc) return rc;
    vqe_paulisum ps;
    const int nl = n_qubits - n_global;
    rc = build_paulisum(&ps, n_qubits, nl, tile_bits, low_bits, 512, std::move(terms), true);
    if (rc) return rc;
    const double2* psi = reinterpret_cast<const double2*>(psi_re_im);
    double2* sigma = reinterpret_cast<double2*>(sigma_re_im);
    const uint64_t sign_base = (uint64_t)rank << nl;
    double total = 0.0;
    size_t lean_terms = 0, fat_terms = 0;
    for (const PSPass& p : ps.passes) {
        if (!p.lean) {
            fat_terms += p.terms_expect.size();
            continue;
        }
        lean_terms += p.lean_terms;
        const uint32_t ts = 1u << p.tp.tbits;
        std::vector<double2> tile(ts), acc(ts);
        std::vector<uint64_t> addr(ts);
        std::vector<double> beta(p.addpat.size());
        const uint32_t lmask = (1u << p.tp.lbits) - 1u;
        for (uint64_t t = 0; t < p.tp.n_tiles; ++t) {
            uint64_t base = 0, v = t, m = p.tp.comp_mask;
            while (m) {  // pdep
                const uint64_t low = m & (0 - m);
                if (v & 1) base |= low;
                v >>= 1;
                m ^= low;
            }
            const uint64_t sbase = base | sign_base;
            const uint32_t sw = p.swz ? 0x70u : 0u;   // the kernels see the tile in the TMA 128-byte swizzle
            for (uint32_t k = 0; k < ts; ++k) {
                addr[swz_idx(k, sw)] = base | p.tp.scat[k >> p.tp.lbits] | (uint64_t)(k & lmask);
                tile[swz_idx(k, sw)] = psi[base | p.tp.scat[k >> p.tp.lbits] | (uint64_t)(k & lmask)];
                acc[k] = make_double2(0.0, 0.0);
            }
            for (size_t k = 0; k < p.addpat.size(); ++k) {
                double b = p.addpat[k].beta0;
                for (uint32_t q = 0; q < p.addpat[k].out_count; ++q) {
                    const DevAddOut& ao = p.addout[p.addpat[k].out_begin + q];
                    b += (popc64(sbase & ao.zrel) & 1) ? -ao.c : ao.c;
                }
                beta[k] = b;
            }
            if (getenv("VQE_DEBUG_LEAN_RL") && atoi(getenv("VQE_DEBUG_LEAN_RL")) != 0) {
                // the REAL-LAYOUT twin of the entries (lean_entry_to_rl) on a tile of doubles holding Re(psi), swizzled the way
                // the real-layout tensor map delivers it: what k_expect_lean<true, T, true> evaluates (expectation only)
                const bool swz_r = plan_tma(p.tp, true, true).ok;
                if (!swz_r && !plan_tma(p.tp, false, true).ok) return fail(VQE_ERR_INVALID, "pass without a real-layout tensor-map form");
                const uint32_t swr = swz_r ? 0x70u : 0u;
                std::vector<double> tr(ts);
                for (uint32_t k = 0; k < ts; ++k) tr[swz_idx8(k, swr)] = psi[base | p.tp.scat[k >> p.tp.lbits] | (uint64_t)(k & lmask)].x;
                if (atoi(getenv("VQE_DEBUG_LEAN_RL")) == 2) {
                    // the twin lowered for the real layout with its lane table (PSPass::rlp): what k_expect_rlp / k_expect_rl2 read
                    if (!p.rlp || p.rlp->swz != swz_r || p.rlp->lane_rl.size() != 32 * p.rlp->flats2.size() || p.rlp->flats2.size() != p.flats2.size())
                        return fail(VQE_ERR_INVALID, "pass without a real-layout twin");
                    const PSPass& r = *p.rlp;
                    std::vector<double> rbeta(r.addpat.size());
                    for (size_t k = 0; k < r.addpat.size(); ++k) {
                        double b = r.addpat[k].beta0;
                        for (uint32_t q = 0; q < r.addpat[k].out_count; ++q) {
                            const DevAddOut& ao = r.addout[r.addpat[k].out_begin + q];
                            b += (popc64(sbase & ao.zrel) & 1) ? -ao.c : ao.c;
                        }
                        rbeta[k] = b;
                    }
                    for (size_t e = 0; e < r.flats2.size(); ++e) {
                        uint4 q[3];
                        memcpy(q, &r.flats2[e], sizeof(DevFlat2));
                        const uint32_t par_out = (uint32_t)popc64(sbase & r.fzout[q[0].w >> 16]);
                        LeanUnit u;
                        lean_decode_uniform(q[0], q[1], q[2], u);
                        for (uint32_t lane = 0; lane < 32; ++lane) {
                            const uint32_t lw = r.lane_rl[e * 32 + lane];
                            for (uint32_t j = 0; j < 8; ++j) {
                                const uint32_t off = (lw & 0xfff8u) ^ u.o[j], offb = off ^ u.lx16;
                                if ((off & 7u) || (offb & 7u) || (off >> 3) >= ts || (offb >> 3) >= ts) return fail(VQE_ERR_INVALID, "real-layout offset out of range");
                                double gw = 0.5 * u.fr;
                                if (u.tab != 0xffffu) gw *= rbeta[u.bidx] + r.addtab[u.tab + lane] + r.addtab[u.hi0 + j];
                                if (((lw & 1u) + (u.jsign >> j) + par_out) & 1u) gw = -gw;
                                total += 2.0 * gw * tr[off >> 3] * tr[offb >> 3];
                            }
                        }
                    }
                    continue;
                }
                for (size_t e = 0; e < p.flats2.size(); ++e) {
                    const DevFlat2 fr = lean_entry_to_rl(p.flats2[e], p.swz, swz_r);
                    uint4 q[3];
                    memcpy(q, &fr, sizeof(DevFlat2));
                    const uint32_t par_out = (uint32_t)popc64(sbase & p.fzout[q[0].w >> 16]);
                    for (uint32_t lane = 0; lane < 32; ++lane) {
                        LeanUnit u;
                        lean_decode(q[0], q[1], q[2], lane, u, 3);
                        for (uint32_t j = 0; j < 8; ++j) {
                            const uint32_t off = swz_off(u.v, swr) ^ u.o[j], offb = off ^ u.lx16;
                            if ((off & 7u) || (offb & 7u) || (off >> 3) >= ts || (offb >> 3) >= ts) return fail(VQE_ERR_INVALID, "real-layout offset out of range");
                            double gw = 0.5 * u.fr;
                            if (u.tab != 0xffffu) gw *= beta[u.bidx] + p.addtab[u.tab + lane] + p.addtab[u.hi0 + j];
                            if ((u.s0 + (u.jsign >> j) + par_out) & 1u) gw = -gw;
                            total += 2.0 * gw * tr[off >> 3] * tr[offb >> 3];
                        }
                    }
                }
                continue;
            }
            for (size_t e = 0; e < p.flats2.size(); ++e) {
                uint4 q[3];
                memcpy(q, &p.flats2[e], sizeof(DevFlat2));
                const uint32_t par_out = (uint32_t)popc64(sbase & p.fzout[q[0].w >> 16]);
                for (uint32_t lane = 0; lane < 32; ++lane) {
                    LeanUnit u;
                    lean_decode(q[0], q[1], q[2], lane, u);
                    for (uint32_t j = 0; j < 8; ++j) {
                        const uint32_t off = swz_off(u.v, sw) ^ u.o[j], offb = off ^ u.lx16;
                        const double2 a = tile[off >> 4], b = tile[offb >> 4];
                        double gw = 0.5 * u.fr;
                        if (u.tab != 0xffffu) gw *= beta[u.bidx] + p.addtab[u.tab + lane] + p.addtab[u.hi0 + j];
                        if ((u.s0 + (u.jsign >> j) + par_out) & 1u) gw = -gw;
                        total += 2.0 * gw * (b.x * a.x + b.y * a.y);
                        acc[offb >> 4].x += gw * a.x; acc[offb >> 4].y += gw * a.y;
                        acc[off >> 4].x += gw * b.x; acc[off >> 4].y += gw * b.y;
                    }
                }
            }
            if (sigma)
                for (uint32_t k = 0; k < ts; ++k) {
                    sigma[addr[k]].x += acc[k].x;
                    sigma[addr[k]].y += acc[k].y;
                }
        }
    }
    *out_re = total;
    if (n_lean_terms) *n_lean_terms = (int32_t)lean_terms;
    if (n_fat_terms) *n_fat_terms = (int32_t)fat_terms;
    return VQE_OK;
}
// Host-only interpreter of k_expect_diag2_rl (no CUDA call; CPU test support): builds the quadratic form of the diagonal part
// (X-mask 0) of the Pauli sum exactly as vqe_paulisum_create does and evaluates  sum_l psi_l^2 D(l)  on a HOST state of
// 2^n_local doubles of rank `rank`, chunk by chunk, with the kernel's decomposition  D = K(o) + A[t & 31] + B[t >> 5] + T(t).
// *is_form = 0 (and *out_re untouched) when the diagonal part is not a quadratic form (a string with more than two Z
// letters, a complex weight): the GPU path then takes the general pass.
extern "C" int vqe_debug_diag2_host(int n_qubits, int n_global, int rank, int n_terms, const uint64_t* x, const uint64_t* z,
                                    const int32_t* ny, const double* cre, const double* cim, const double* psi_re, double* out_re,
                                    int32_t* is_form) {
    if (n_qubits < 1 || n_qubits > 30 || n_global < 0 || n_global >= n_qubits) return fail(VQE_ERR_INVALID, "bad qubit counts");
    if (!psi_re || !out_re || !is_form) return fail(VQE_ERR_INVALID, "null argument");
    vqe_ctx fake;
    fake.n = n_qubits;
    std::vector<HTerm> terms;
    int rc = collect_terms(&fake, n_terms, x, z, ny, cre, cim, terms);
    if (rc) return rc;
    vqe_paulisum ps;
    const int n = n_qubits, nl = n_qubits - n_global;
    rc = build_paulisum(&ps, n, nl, 12, n_global ? 5 : 4, 512, std::move(terms), true);
    if (rc) return rc;
    const vqe_paulisum::Diag2& d = ps.diag2;
    *is_form = d.on ? 1 : 0;
    if (!d.on) return VQE_OK;
    const int tb = d.tb, lb = tb < 5 ? tb : 5;
    const uint64_t n_amp = 1ull << nl, n_chunks = n_amp >> tb, sign_base = (uint64_t)rank << nl;
    auto sgn = [](uint64_t o, int q, double v) { return ((o >> q) & 1ull) ? -v : v; };
    double total = 0.0;
    for (uint64_t ch = 0; ch < n_chunks; ++ch) {
        const uint64_t o = (ch << tb) | sign_base;
        std::vector<double> v(tb), A(32, 0.0), B((size_t)1 << (tb - lb), 0.0);
        for (int p0 = 0; p0 < tb; ++p0) {
            v[p0] = d.a[p0];
            for (int q0 = tb; q0 < n; ++q0) v[p0] += sgn(o, q0, d.b[(size_t)p0 * n + q0]);
        }
        double K = d.c0;
        for (int p0 = tb; p0 < n; ++p0) {
            double inner = d.a[p0];
            for (int q0 = p0 + 1; q0 < n; ++q0) inner += sgn(o, q0, d.b[(size_t)p0 * n + q0]);
            K += sgn(o, p0, inner);
        }
        for (uint32_t i = 0; i < 32; ++i)
            for (int p0 = 0; p0 < lb; ++p0) A[i] += ((i >> p0) & 1u) ? -v[p0] : v[p0];
        for (uint32_t j = 0; j < B.size(); ++j)
            for (int p0 = lb; p0 < tb; ++p0) B[j] += ((j >> (p0 - lb)) & 1u) ? -v[p0] : v[p0];
        for (uint32_t t = 0; t < (1u << tb); ++t) {
            const double xv = psi_re[(ch << tb) | t];
            total += xv * xv * (K + A[t & 31u] + B[t >> lb] + d.t[t]);
        }
    }
    *out_re = total;
    return VQE_OK;
}
extern "C" void vqe_paulisum_destroy(vqe_paulisum* ps) {
    if (!ps) return;
    cudaSetDevice(ps->device);
    for (auto& kv : ps->variants) {
        free_paulisum_device(kv.second);
        delete kv.second;
    }
    free_paulisum_device(ps);
    delete ps;
}
// A twin of `ps` for the current relabelling of context c (built, uploaded and cached on first use).  tag '*': the whole
// sum; 'A' / 'B': the two parts of a split evaluation (see expectation_impl), selected by `pick`.
static int paulisum_variant(vqe_ctx* c, const vqe_paulisum* ps, char tag, const std::vector<char>* pick, const vqe_paulisum** out) {
    *out = ps;
    if (c->world == 1 || (tag == '*' && perm_is_identity(c))) return VQE_OK;
    if (ps->terms.empty() && !ps->passes.empty())
        return fail(VQE_ERR_INVALID, "the Pauli sum was created on an unsharded context and cannot follow a relabelled state");
    vqe_paulisum* root = const_cast<vqe_paulisum*>(ps);
    std::string key((const char*)c->perm, (size_t)c->n);
    key.push_back(tag);
    if (pick) {  // the selection is part of the identity of the twin
        uint64_t h = 1469598103934665603ull;
        for (char v : *pick) h = (h ^ (uint64_t)(unsigned char)v) * 1099511628211ull;
        key.append((const char*)&h, sizeof h);
    }
    auto it = root->variants.find(key);
    if (it != root->variants.end()) {
        *out = it->second;
        return VQE_OK;
    }
    std::vector<HTerm> terms;
    terms.reserve(ps->terms.size());
    for (size_t k = 0; k < ps->terms.size(); ++k) {
        if (pick && !(*pick)[k]) continue;
        HTerm t = ps->terms[k];
        t.x = perm_mask(c->perm, t.x);
        t.z = perm_mask(c->perm, t.z);
        terms.push_back(t);
    }
    vqe_paulisum* v = new vqe_paulisum();
    v->device = ps->device;
    v->tile_bits = ps->tile_bits;
    v->low_bits = ps->low_bits;
    v->threads = ps->threads;
    CK(cudaSetDevice(c->device));
    int rc = build_paulisum(v, ps->n, ps->nl, ps->tile_bits, ps->low_bits, ps->threads, std::move(terms));
    if (rc == VQE_OK) rc = upload_paulisum(c, v);
    if (rc) {
        free_paulisum_device(v);
        delete v;
        return rc;
    }
    if (root->variants.size() >= 12) {  // an optimisation re-uses one relabelling; keep the cache small
        auto oldest = root->variants.begin();
        free_paulisum_device(oldest->second);
        delete oldest->second;
        root->variants.erase(oldest);
    }
    root->variants[key] = v;
    *out = v;
    return VQE_OK;
}
extern "C" int vqe_paulisum_groups(const vqe_paulisum* ps) { return ps ? ps->n_groups : 0; }
extern "C" int vqe_paulisum_passes(const vqe_paulisum* ps) { return ps ? (int)ps->passes.size() : 0; }

// <buf|O|buf> on every rank of the set.  Peer passes (groups whose X-mask flips global bits) read the partner's
// shard over NVLink; each super-tile is evaluated by exactly one rank of the pair.  out_per_rank[2r], [2r+1] =
// partial sum of rank rs.r[r]; the caller adds the partials in rank order.
static int expectation_core(RankSet& rs, int b, const vqe_paulisum* const* pss, double* out_per_rank) {
    int rc = check_rankset(rs);
    if (rc) return rc;
    if (!pss || !out_per_rank) return fail(VQE_ERR_INVALID, "null argument");
    const size_t nr = rs.r.size();
    for (size_t k = 0; k < nr; ++k) {
        vqe_ctx* c = rs.r[k];
        const vqe_paulisum* ps = pss[k];
        if (!ps) return fail(VQE_ERR_INVALID, "null Pauli sum");
        if (ps->n != c->n || ps->nl != c->nl)
            return fail(VQE_ERR_INVALID, "Pauli sum built for %d qubits (%d local), context has %d (%d local)", ps->n,
                        ps->nl, c->n, c->nl);
        if (ps->device != c->device) return fail(VQE_ERR_INVALID, "Pauli sum lives on device %d, context on %d", ps->device, c->device);
        if (ps->passes.size() != pss[0]->passes.size()) return fail(VQE_ERR_INVALID, "Pauli sums of the ranks differ");
        CK(cudaSetDevice(c->device));
        rc = ensure_buf(c, b);
        if (rc) return rc;
    }
    const size_t n_pass = pss[0]->passes.size();
    bool real_state = b == VQE_BUF_PSI;  // the purely-real flag is only tracked for the state buffer
    for (vqe_ctx* c : rs.r) real_state = real_state && c->psi_real;
    // Real layout of the state (unsharded): the lean passes read it as it is when every one of them has a real-layout form;
    // the buffer is expanded to interleaved complex before the first general pass (or right away otherwise).
    // (lean passes are always local passes, also on a shard; peer and general passes come after them)
    bool rl = b == VQE_BUF_PSI && real_state && env_int("VQE_PIPE", 0) == 0;
    for (size_t k = 0; k < nr && rl; ++k) {
        rl = rs.r[k]->real_layout;
        for (const PSPass& pp : pss[k]->passes)
            if (pp.lean && !pp.rl_ok) rl = false;
    }
    if (!rl && b == VQE_BUF_PSI)
        for (vqe_ctx* c : rs.r) {
            CK(cudaSetDevice(c->device));
            rc = ensure_complex(c, b);
            if (rc) return rc;
        }
    // per rank: grids and partial-sum layout (consecutive blocks of every pass this rank launches)
    std::vector<std::vector<dim3>> grids(nr, std::vector<dim3>(n_pass));
    std::vector<std::vector<TileGeom>> geoms(nr, std::vector<TileGeom>(n_pass));
    std::vector<std::vector<Shards>> shards(nr, std::vector<Shards>(n_pass));
    std::vector<size_t> total_blocks(nr, 0);
    for (size_t k = 0; k < nr; ++k) {
        vqe_ctx* c = rs.r[k];
        for (size_t p = 0; p < n_pass; ++p) {
            const PSPass& pp = pss[k]->passes[p];
            rc = make_geom(c, pp.tp, pp.d_scat, b, geoms[k][p], shards[k][p]);
            if (rc) return rc;
            if (geoms[k][p].n_tiles == 0) {
                grids[k][p] = dim3(0, 0, 1);
                continue;
            }
            if (rl && pp.diag_only && pss[k]->diag2.on) {
                // the diagonal part on the real layout (k_expect_diag2_rl): one launch for all X-mask-0 passes, no expansion
                bool first = true;
                for (size_t p2 = 0; p2 < p; ++p2) first = first && !pss[k]->passes[p2].diag_only;
                const uint64_t n_chunks = c->n_amp >> pss[k]->diag2.tb;
                grids[k][p] = first ? dim3((unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n_chunks, (uint64_t)c->sm_count * 8)), 1, 1) : dim3(0, 0, 1);
                total_blocks[k] += first ? grids[k][p].x : 0;
                continue;
            }
            const bool pipe = pp.lean && geoms[k][p].bulk && env_int("VQE_PIPE", 0) != 0 &&
                              3 * tile_smem(pp.tp.tbits, 1, false) + 2 * pp.addpat.size() * sizeof(double) <= 226 * 1024;
            // real layout: 32 KiB tiles; two slots per CTA and 3 CTAs per SM (k_expect_rl2), or one slot and 4 CTAs (64 registers)
            const bool rl2 = env_int("VQE_EXP_RL2", 1) != 0 && 2 * (8ull << pp.tp.tbits) + pp.addpat.size() * 16 <= 74 * 1024;
            const int lean_ctas = (rl && pp.lean) ? std::max(1, std::min(6, env_int("VQE_EXP_RL_CTAS", rl2 ? 3 : 4))) : 3;
            int gx = tile_grid(c, geoms[k][p].n_tiles, pipe ? 1 : (pp.lean ? lean_ctas : 0));
            int want = std::max(1, (c->sm_count * (pipe ? 1 : (pp.lean ? lean_ctas : c->ctas_per_sm))) / gx);
            int gy = std::max(1, std::min<int>(pp.lean ? (int)((pp.flats2.size() + 7) / 8) : (int)pp.groups.size(), want));
            grids[k][p] = dim3(gx, gy, 1);
            total_blocks[k] += (size_t)gx * gy;
        }
        CK(cudaSetDevice(c->device));
        rc = ensure_partial(c, std::max<size_t>(1, total_blocks[k]));
        if (rc) return rc;
    }
    std::vector<size_t> off(nr, 0);
    bool fenced = false;
    // the passes alternate their walk starting from the direction the state was last written in; the record is put back
    // afterwards (the passes only read), so a repeated evaluation of the same state takes the same directions -- and sums
    // its partials in the same order
    std::vector<char> walk0(nr, 0);
    for (size_t k = 0; k < nr; ++k) walk0[k] = rs.r[k]->walk_desc ? 1 : 0;
    for (size_t p = 0; p < n_pass; ++p) {
        const bool vbit = pss[0]->passes[p].tp.vbit;
        const bool diag_rl = rl && pss[0]->passes[p].diag_only && pss[0]->diag2.on;   // evaluated on the real layout
        if (rl && !pss[0]->passes[p].lean && !diag_rl)  // first general pass: every rank expands its shard BEFORE any partner reads it
            for (vqe_ctx* c : rs.r)
                if (c->real_layout) {
                    CK(cudaSetDevice(c->device));
                    rc = ensure_complex(c, b);
                    if (rc) return rc;
                }
        if (vbit && !fenced) {
            rc = rank_barrier(rs);
            if (rc) return rc;
        }
        for (size_t k = 0; k < nr; ++k) {
            vqe_ctx* c = rs.r[k];
            const PSPass& pp = pss[k]->passes[p];
            if (grids[k][p].x == 0) continue;
            CK(cudaSetDevice(c->device));
            size_t smem = tile_smem(pp.tp.tbits, 1, true);
            int threads = (int)std::min<uint64_t>(c->threads, std::max<uint64_t>(32, (1ull << pp.tp.tbits) / 2));
            ProfScope prof(c, vbit ? 5 : 1);
            CUtensorMap tmap;
            memset(&tmap, 0, sizeof tmap);
            if (diag_rl && pss[k]->diag2.on && c->real_layout) {
                const vqe_paulisum::Diag2& dg = pss[k]->diag2;
                if (env_int("VQE_ALT_ORDER", 1) != 0) c->walk_desc = !c->walk_desc;
                k_expect_diag2_rl<<<grids[k][p].x, 256, 0, c->stream>>>(reinterpret_cast<const double*>(shards[k][p].p0), c->n_amp, c->n, dg.tb,
                                                                       geoms[k][p].sign_base, dg.c0, dg.d_a, dg.d_b, dg.d_t, c->walk_desc ? 1 : 0,
                                                                       c->d_partial + off[k]);
                c->launches++;
                CK(cudaGetLastError());
                off[k] += (size_t)grids[k][p].x;
                continue;
            }
            if (rl && !pp.lean && c->real_layout) {  // first general pass: it reads the interleaved form
                rc = ensure_complex(c, b);
                if (rc) return rc;
            }
            const bool rl_pass = rl && pp.lean && c->real_layout;
            if (rl_pass) {
                make_tmap(pp.tp, shards[k][p].p0, geoms[k][p], &tmap, pp.rl_swz ? 1 : 0, true);
                if (!geoms[k][p].tma) return fail(VQE_ERR_CUDA, "real-layout tensor map of an expectation pass could not be encoded");
                const size_t smem_r = (8ull << pp.tp.tbits) + pp.addpat.size() * sizeof(double);
                const int thr_l = env_int("VQE_EXP_LEAN_THREADS", 256);
                // entries of the pass: the real-layout twin with its lane table (PSPass::rlp) when it was uploaded, else the
                // converted entries of the interleaved form (in-kernel index deposit)
                const bool lane = pp.rl_lane;
                const int n_fl = lane ? (int)pp.rlp->flats2.size() : (int)pp.flats2.size();
                const int n_ap = lane ? (int)pp.rlp->addpat.size() : (int)pp.addpat.size();
                const uint64_t* fz = lane ? pp.d_fzout_rl : pp.d_fzout;
                const double* atab = lane ? pp.d_addtab_rl : pp.d_addtab;
                const DevAddPat* apat = lane ? pp.d_addpat_rl : pp.d_addpat;
                const DevAddOut* aout = lane ? pp.d_addout_rl : pp.d_addout;
                const uint16_t* ltab = lane ? pp.d_lane_rl : nullptr;
                const size_t smem_2 = 2 * (8ull << pp.tp.tbits) + 2 * (size_t)n_ap * sizeof(double);
                if (b == VQE_BUF_PSI && geoms[k][p].tile_stride == 1 && geoms[k][p].tile_first == 0 && env_int("VQE_ALT_ORDER", 1) != 0) {
                    c->walk_desc = !c->walk_desc;   // alternating walk, as for the rotation passes
                    if (c->walk_desc) {
                        geoms[k][p].rev = 1;
                        geoms[k][p].tile_first = geoms[k][p].n_tiles - 1;
                    }
                }
                if (env_int("VQE_EXP_RL2", 1) != 0 && smem_2 <= 74 * 1024 && (thr_l == 256 || thr_l == 384)) {
                    cudaLaunchConfig_t cfg;
                    memset(&cfg, 0, sizeof cfg);
                    cfg.gridDim = grids[k][p];
                    cfg.blockDim = dim3((unsigned)thr_l);
                    cfg.dynamicSmemBytes = smem_2;
                    cfg.stream = c->stream;
                    cudaLaunchAttribute at[2];
                    unsigned n_at = 0;
                    if (env_int("VQE_PDL", 1) != 0) {
                        at[n_at].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                        at[n_at].val.programmaticStreamSerializationAllowed = 1;
                        ++n_at;
                    }
                    n_at += (unsigned)l2_window_attr(c, &at[n_at], shards[k][p].p0, (size_t)c->n_amp * sizeof(double));
                    cfg.attrs = at;
                    cfg.numAttrs = n_at;
                    if (thr_l == 384)
                        CK(cudaLaunchKernelEx(&cfg, k_expect_rl2<384>, tmap, geoms[k][p], (const DevFlat2*)pp.d_flats2_rl, n_fl, fz, atab, apat, n_ap,
                                              aout, ltab, c->d_partial + off[k], c->d_err));
                    else if (env_int("VQE_EXP_PAIR", 1) != 0)
                        CK(cudaLaunchKernelEx(&cfg, k_expect_rlp<256>, tmap, geoms[k][p], (const DevFlat2*)pp.d_flats2_rl, n_fl, fz, atab, apat, n_ap,
                                              aout, ltab, c->d_partial + off[k], c->d_err));
                    else
                        CK(cudaLaunchKernelEx(&cfg, k_expect_rl2<256>, tmap, geoms[k][p], (const DevFlat2*)pp.d_flats2_rl, n_fl, fz, atab, apat, n_ap,
                                              aout, ltab, c->d_partial + off[k], c->d_err));
                    c->launches++;
                    off[k] += (size_t)grids[k][p].x * grids[k][p].y;
                    continue;
                }
                if (lane) return fail(VQE_ERR_INVALID, "the Pauli sum was lowered with a lane table (VQE_EXP_LANE_TAB), which only the two-slot "
                                                       "real-layout kernels read: VQE_EXP_RL2 / VQE_EXP_LEAN_THREADS changed after vqe_paulisum_create?");
#define LAUNCH_EXPECT_RL(T)                                                                                                    \
    k_expect_lean<true, T, true><<<grids[k][p], T, smem_r, c->stream>>>(tmap, shards[k][p], geoms[k][p], pp.d_flats2_rl,       \
                                                                       (int)pp.flats2.size(), pp.d_fzout, pp.d_addtab, pp.d_addpat, \
                                                                       (int)pp.addpat.size(), pp.d_addout, c->d_partial + off[k], c->d_err)
                if (thr_l == 512) LAUNCH_EXPECT_RL(512);
                else if (thr_l == 384) LAUNCH_EXPECT_RL(384);
                else LAUNCH_EXPECT_RL(256);
#undef LAUNCH_EXPECT_RL
                c->launches++;
                CK(cudaGetLastError());
                off[k] += (size_t)grids[k][p].x * grids[k][p].y;
                continue;
            }
            if (pp.lean) {
                make_tmap(pp.tp, vbit ? nullptr : shards[k][p].p0, geoms[k][p], &tmap, pp.swz ? 1 : 0);
                if (pp.swz && !geoms[k][p].tma)
                    return fail(VQE_ERR_CUDA, "the Pauli sum was lowered for tensor-map (TMA) tile loads, which are not available now "
                                              "(VQE_TMA / VQE_SWIZZLE changed after vqe_paulisum_create?)");
            }
            const bool pipe = pp.lean && (geoms[k][p].bulk || geoms[k][p].tma) && env_int("VQE_PIPE", 0) != 0 &&
                              3 * tile_smem(pp.tp.tbits, 1, false) + 2 * pp.addpat.size() * sizeof(double) <= 226 * 1024;
            if (pipe) {
                const size_t smem_p = 3 * tile_smem(pp.tp.tbits, 1, false) + 2 * pp.addpat.size() * sizeof(double);
                int thr_p = env_int("VQE_EXP_THREADS", 512);
                if (thr_p < 32 || thr_p > 512 || (thr_p & 31)) thr_p = 512;
                if (env_int("VQE_DEBUG_SKELETON", 0))  // measurement aid: loads only (results are wrong)
                    k_expect_pipe<true><<<grids[k][p], thr_p, smem_p, c->stream>>>(tmap, shards[k][p], geoms[k][p], pp.d_flats2, 0,
                                                                                  pp.d_fzout, pp.d_addtab, pp.d_addpat, 0,
                                                                                  pp.d_addout, c->d_partial + off[k], c->d_err);
                else if (real_state)
                    k_expect_pipe<true><<<grids[k][p], thr_p, smem_p, c->stream>>>(tmap, shards[k][p], geoms[k][p], pp.d_flats2, (int)pp.flats2.size(),
                                                                                  pp.d_fzout, pp.d_addtab, pp.d_addpat, (int)pp.addpat.size(),
                                                                                  pp.d_addout, c->d_partial + off[k], c->d_err);
                else
                    k_expect_pipe<false><<<grids[k][p], thr_p, smem_p, c->stream>>>(tmap, shards[k][p], geoms[k][p], pp.d_flats2, (int)pp.flats2.size(),
                                                                                   pp.d_fzout, pp.d_addtab, pp.d_addpat, (int)pp.addpat.size(),
                                                                                   pp.d_addout, c->d_partial + off[k], c->d_err);
            } else if (pp.lean) {
                const size_t smem_l = tile_smem(pp.tp.tbits, 1, false) + pp.addpat.size() * sizeof(double);
                const int thr_l = env_int("VQE_EXP_LEAN_THREADS", 256);
#define LAUNCH_EXPECT_LEAN(R, T)                                                                                              \
    k_expect_lean<R, T><<<grids[k][p], T, smem_l, c->stream>>>(tmap, shards[k][p], geoms[k][p], pp.d_flats2, (int)pp.flats2.size(), \
                                                              pp.d_fzout, pp.d_addtab, pp.d_addpat, (int)pp.addpat.size(),    \
                                                              pp.d_addout, c->d_partial + off[k], c->d_err)
                if (real_state) {
                    if (thr_l == 512) LAUNCH_EXPECT_LEAN(true, 512);
                    else if (thr_l == 384) LAUNCH_EXPECT_LEAN(true, 384);
                    else LAUNCH_EXPECT_LEAN(true, 256);
                } else {
                    if (thr_l == 512) LAUNCH_EXPECT_LEAN(false, 512);
                    else if (thr_l == 384) LAUNCH_EXPECT_LEAN(false, 384);
                    else LAUNCH_EXPECT_LEAN(false, 256);
                }
#undef LAUNCH_EXPECT_LEAN
            } else if (pp.cplx)
                k_tile_expect<true><<<grids[k][p], threads, smem, c->stream>>>(shards[k][p], geoms[k][p], pp.d_groups,
                                                                              (int)pp.groups.size(), pp.d_terms_expect,
                                                                              pp.d_gcols, (int)pp.gcols.size(), pp.d_gents,
                                                                              (int)pp.gents.size(), pp.d_flats, (int)pp.flats.size(),
                                                                              pp.d_fzout, c->d_partial + off[k], c->d_err);
            else
                k_tile_expect<false><<<grids[k][p], threads, smem, c->stream>>>(shards[k][p], geoms[k][p], pp.d_groups,
                                                                               (int)pp.groups.size(), pp.d_terms_expect,
                                                                               pp.d_gcols, (int)pp.gcols.size(), pp.d_gents,
                                                                               (int)pp.gents.size(), pp.d_flats, (int)pp.flats.size(),
                                                                               pp.d_fzout, c->d_partial + off[k], c->d_err);
            c->launches++;
            CK(cudaGetLastError());
            off[k] += (size_t)grids[k][p].x * grids[k][p].y;
        }
        fenced = false;
        if (vbit) {  // nobody may modify its shard while a partner still reads it
            rc = rank_barrier(rs);
            if (rc) return rc;
            fenced = true;
        }
    }
    for (size_t k = 0; k < nr; ++k) rs.r[k]->walk_desc = walk0[k] != 0;
    for (size_t k = 0; k < nr; ++k) {
        vqe_ctx* c = rs.r[k];
        out_per_rank[2 * k] = out_per_rank[2 * k + 1] = 0.0;
        if (total_blocks[k] == 0) continue;
        CK(cudaSetDevice(c->device));
        k_reduce_partials<<<1, 1024, 0, c->stream>>>(c->d_partial, (int)total_blocks[k], 1, 1, c->d_result);
        c->launches++;
        c->d2h_bytes += sizeof(double2);
        CK(cudaMemcpyAsync(c->h_result, c->d_result, sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    }
    for (size_t k = 0; k < nr; ++k) {
        vqe_ctx* c = rs.r[k];
        if (total_blocks[k] == 0) continue;
        CK(cudaSetDevice(c->device));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaGetLastError());
        out_per_rank[2 * k] = c->h_result[0].x;
        out_per_rank[2 * k + 1] = c->h_result[0].y;
    }
    for (vqe_ctx* c : rs.r) {
        rc = vqe_shard_status(c);
        if (rc) return rc;
    }
    return VQE_OK;
}

// <buf|O|buf>.  On a (relabelled) sharded state the sum is evaluated through its relabelled twin; when strings would
// still flip a global qubit the evaluation is SPLIT: part A -- the X-mask groups that are local as the state stands --
// first, then up to n_global qubit swaps that move the qubits part B flips out of the global slots (into local slots whose
// qubits part B never flips), then part B (all the rest, diagonal group included) on local passes.  A peer pass reads
// half a shard over NVLink per global pattern (8 x 102 ms at 36 qubits on 8 GPUs); a swap costs half of that once.
static int expectation_impl(RankSet& rs, int b, const vqe_paulisum* const* pss_in, double* out_per_rank) {
    int rc = check_rankset(rs);
    if (rc) return rc;
    if (!pss_in || !out_per_rank) return fail(VQE_ERR_INVALID, "null argument");
    const size_t nr = rs.r.size();
    for (size_t k = 0; k < nr; ++k)
        if (!pss_in[k]) return fail(VQE_ERR_INVALID, "null Pauli sum");
    vqe_ctx* c0 = rs.r[0];
    std::vector<const vqe_paulisum*> eff(pss_in, pss_in + nr);
    if (b != VQE_BUF_PSI || c0->world == 1) return expectation_core(rs, b, eff.data(), out_per_rank);
    const vqe_paulisum* ps0 = pss_in[0];
    const uint64_t gmask = ((1ull << c0->g) - 1ull) << c0->nl;
    const bool can_split = env_int("VQE_RELABEL", 1) != 0 && c0->nl >= 8 && !ps0->terms.empty();
    std::vector<char> in_a(ps0->terms.size(), 0), in_b(ps0->terms.size(), 0);
    bool any_a = false, any_peer = false;
    if (can_split)
        for (size_t k = 0; k < ps0->terms.size(); ++k) {
            const uint64_t px = perm_mask(c0->perm, ps0->terms[k].x);
            if (px & gmask) any_peer = true;
            in_a[k] = (ps0->terms[k].x != 0 && !(px & gmask)) ? 1 : 0;
            in_b[k] = in_a[k] ? 0 : 1;
            any_a = any_a || in_a[k];
        }
    if (!can_split || !any_peer) {
        for (size_t k = 0; k < nr; ++k) {
            rc = paulisum_variant(rs.r[k], pss_in[k], '*', nullptr, &eff[k]);
            if (rc) return rc;
        }
        return expectation_core(rs, b, eff.data(), out_per_rank);
    }
    for (size_t k = 0; k < nr; ++k)
        if (pss_in[k]->terms.size() != ps0->terms.size()) return fail(VQE_ERR_INVALID, "Pauli sums of the ranks differ");
    std::vector<double> part(2 * nr, 0.0);
    for (size_t k = 0; k < 2 * nr; ++k) out_per_rank[k] = 0.0;
    if (any_a) {
        for (size_t k = 0; k < nr; ++k) {
            rc = paulisum_variant(rs.r[k], pss_in[k], 'A', &in_a, &eff[k]);
            if (rc) return rc;
        }
        rc = expectation_core(rs, b, eff.data(), part.data());
        if (rc) return rc;
        for (size_t k = 0; k < 2 * nr; ++k) out_per_rank[k] += part[k];
    }
    // qubits part B flips, per logical bit
    const int n = c0->n, nl = c0->nl;
    std::vector<int> busy(n, 0);
    for (size_t k = 0; k < ps0->terms.size(); ++k)
        if (in_b[k])
            for (uint64_t m = ps0->terms[k].x; m; m &= m - 1) busy[__builtin_ctzll(m)]++;
    const int floor_slot = std::max(2, std::min(env_int("VQE_RELABEL_FLOOR", 10), nl - 4));
    for (int gs = nl; gs < n; ++gs) {
        int inv[64];
        for (int q = 0; q < n; ++q) inv[c0->perm[q]] = q;
        if (busy[inv[gs]] == 0) continue;
        int best = -1;
        for (int slot = nl - 1; slot >= floor_slot; --slot)
            if (best < 0 || busy[inv[slot]] < busy[inv[best]]) best = slot;
        if (best < 0 || busy[inv[best]] >= busy[inv[gs]]) continue;
        rc = swap_global_local(rs, gs, best);
        if (rc) return rc;
    }
    for (size_t k = 0; k < nr; ++k) {
        rc = paulisum_variant(rs.r[k], pss_in[k], 'B', &in_b, &eff[k]);
        if (rc) return rc;
    }
    rc = expectation_core(rs, b, eff.data(), part.data());
    if (rc) return rc;
    for (size_t k = 0; k < 2 * nr; ++k) out_per_rank[k] += part[k];
    return VQE_OK;
}
extern "C" int vqe_expectation(vqe_ctx* c, int b, const vqe_paulisum* ps, double* out) {
    if (!c || !ps || !out) return fail(VQE_ERR_INVALID, "null argument");
    RankSet rs;
    rs.r.push_back(c);
    return expectation_impl(rs, b, &ps, out);
}
extern "C" int vqe_group_expectation(vqe_ctx* const* ranks, int n_ranks, int b, const vqe_paulisum* const* ps, double* out) {
    if (!out) return fail(VQE_ERR_INVALID, "null argument");
    RankSet rs = rankset_of(ranks, n_ranks);
    std::vector<double> per(2 * std::max(1, n_ranks), 0.0);
    int rc = expectation_impl(rs, b, ps, per.data());
    if (rc) return rc;
    out[0] = out[1] = 0.0;
    for (int k = 0; k < n_ranks; ++k) {  // fixed order: rank 0, 1, ...
        out[0] += per[2 * k];
        out[1] += per[2 * k + 1];
    }
    return VQE_OK;
}

// dst <- O src on every rank of the set (buffer ids).  Peer passes read src and read-modify-write dst of the
// partner's shard; every (super-)tile of dst is written by exactly one CTA per pass.
static int apply_paulisum_rs(RankSet& rs, int dst, int src, const vqe_paulisum* const* pss) {
    const size_t nr = rs.r.size();
    if (dst == VQE_BUF_PSI || src == VQE_BUF_PSI) {
        int rcl = need_caller_labelling(rs);
        if (rcl) return rcl;
    }
    for (size_t k = 0; k < nr; ++k) {
        vqe_ctx* c = rs.r[k];
        CK(cudaSetDevice(c->device));
        int rc = ensure_buf(c, dst);
        if (rc) return rc;
        rc = ensure_buf(c, src);
        if (rc) return rc;
    }
    for (vqe_ctx* c : rs.r) {
        CK(cudaSetDevice(c->device));
        int rc = ensure_complex(c, src);
        if (rc == VQE_OK) rc = ensure_complex(c, dst);
        if (rc) return rc;
    }
    bool real_src = src == VQE_BUF_PSI;  // lean passes have real weights: a purely real source gives a purely real sigma
    for (vqe_ctx* c : rs.r) real_src = real_src && c->psi_real;
    if (dst == VQE_BUF_PSI)
        for (vqe_ctx* c : rs.r) c->psi_real = false;
    const size_t n_pass = pss[0]->passes.size();
    if (n_pass == 0) {
        for (vqe_ctx* c : rs.r) {
            CK(cudaSetDevice(c->device));
            CK(cudaMemsetAsync(c->buf[dst], 0, c->n_amp * sizeof(double2), c->stream));
        }
        return VQE_OK;
    }
    bool fenced = false;
    for (size_t p = 0; p < n_pass; ++p) {
        const bool vbit = pss[0]->passes[p].tp.vbit;
        if (vbit && !fenced) {
            int rc = rank_barrier(rs);
            if (rc) return rc;
        }
        for (size_t k = 0; k < nr; ++k) {
            vqe_ctx* c = rs.r[k];
            const PSPass& pp = pss[k]->passes[p];
            CK(cudaSetDevice(c->device));
            TileGeom g;
            Shards ssrc, sdst;
            int rc = make_geom(c, pp.tp, pp.d_scat, src, g, ssrc);
            if (rc == VQE_OK) rc = make_geom(c, pp.tp, pp.d_scat, dst, g, sdst);
            if (rc) return rc;
            if (g.n_tiles == 0) continue;
            size_t smem = 2 * (16ull << pp.tp.tbits) + TERM_CAP * (sizeof(double2) + 4) +
                          AFLAT_CAP * (sizeof(DevAFlat) + sizeof(double2)) + (GROUP_CAP + 1) * 4;
            uint64_t ts = 1ull << pp.tp.tbits;
            int threads = (int)std::min<uint64_t>(1024, std::max<uint64_t>(32, ts / 2));
            ProfScope prof(c, 2);
            if (pp.lean) {
                CUtensorMap tmap;
                make_tmap(pp.tp, vbit ? nullptr : ssrc.p0, g, &tmap, pp.swz ? 1 : 0);
                if (pp.swz && !g.tma)
                    return fail(VQE_ERR_CUDA, "the Pauli sum was lowered for tensor-map (TMA) tile loads, which are not available now");
                const int n_lg = (int)pp.boff.size() - 1;  // conflict-free batches of entries
                const int thr = (int)std::min<uint64_t>(512, std::max<uint64_t>(32, ts / 2));
                if (real_src) {
                    const size_t smem_l = (16ull << pp.tp.tbits) + (8ull << pp.tp.tbits) + pp.addpat.size() * sizeof(double);
                    k_apply_lean<true><<<tile_grid(c, g.n_tiles, 2), thr, smem_l, c->stream>>>(
                        tmap, ssrc, sdst, g, pp.d_flats2, pp.d_goff, n_lg, pp.d_fzout, pp.d_addtab, pp.d_addpat, (int)pp.addpat.size(),
                        pp.d_addout, p == 0 ? 0 : 1, c->d_err);
                } else {
                    const size_t smem_l = 2 * (16ull << pp.tp.tbits) + pp.addpat.size() * sizeof(double);
                    k_apply_lean<false><<<tile_grid(c, g.n_tiles, 1), thr, smem_l, c->stream>>>(
                        tmap, ssrc, sdst, g, pp.d_flats2, pp.d_goff, n_lg, pp.d_fzout, pp.d_addtab, pp.d_addpat, (int)pp.addpat.size(),
                        pp.d_addout, p == 0 ? 0 : 1, c->d_err);
                }
                c->launches++;
                CK(cudaGetLastError());
                continue;
            }
            k_tile_apply<<<tile_grid(c, g.n_tiles, 1), threads, smem, c->stream>>>(ssrc, sdst, g, pp.d_groups,
                                                                                  (int)pp.groups.size(), pp.d_terms_apply,
                                                                                  pp.d_aflat, pp.d_aoff, pp.d_azout,
                                                                                  p == 0 ? 0 : 1);
            c->launches++;
            CK(cudaGetLastError());
        }
        fenced = false;
        if (vbit) {
            int rc = rank_barrier(rs);
            if (rc) return rc;
            fenced = true;
        }
    }
    return VQE_OK;
}
static int apply_paulisum_bufs(vqe_ctx* c, int dst, int src, const vqe_paulisum* ps) {
    RankSet rs;
    rs.r.push_back(c);
    return apply_paulisum_rs(rs, dst, src, &ps);
}

static int check_apply_args(RankSet& rs, int dst, int src, const vqe_paulisum* const* pss) {
    int rc = check_rankset(rs);
    if (rc) return rc;
    if (!pss) return fail(VQE_ERR_INVALID, "null argument");
    if (dst == src) return fail(VQE_ERR_INVALID, "dst and src buffers must differ");
    for (size_t k = 0; k < rs.r.size(); ++k) {
        if (!pss[k]) return fail(VQE_ERR_INVALID, "null Pauli sum");
        if (pss[k]->n != rs.r[k]->n || pss[k]->nl != rs.r[k]->nl)
            return fail(VQE_ERR_INVALID, "Pauli sum built for %d qubits, context has %d", pss[k]->n, rs.r[k]->n);
        if (pss[k]->passes.size() != pss[0]->passes.size()) return fail(VQE_ERR_INVALID, "Pauli sums of the ranks differ");
    }
    return VQE_OK;
}
extern "C" int vqe_apply_paulisum(vqe_ctx* c, int dst, int src, const vqe_paulisum* ps) {
    if (!c || !ps) return fail(VQE_ERR_INVALID, "null argument");
    RankSet rs;
    rs.r.push_back(c);
    int rc = check_apply_args(rs, dst, src, &ps);
    if (rc) return rc;
    return apply_paulisum_rs(rs, dst, src, &ps);
}
extern "C" int vqe_group_apply_paulisum(vqe_ctx* const* ranks, int n_ranks, int dst, int src, const vqe_paulisum* const* ps) {
    RankSet rs = rankset_of(ranks, n_ranks);
    int rc = check_apply_args(rs, dst, src, ps);
    if (rc) return rc;
    return apply_paulisum_rs(rs, dst, src, ps);
}

// ---- pool sweep -------------------------------------------------------------------------------
// out_per_rank: nr x (2 * n_ops) doubles, partial overlaps of every rank in the set.  An operator whose strings
// carry different global X patterns is split into one sub-operator per pattern (the overlap is linear in the
// strings); sub-operators with pattern m != 0 run in peer passes between ranks r and r ^ m.
static int pool_impl(RankSet& rs, int bra, int ket, int n_ops, const int32_t* op_offsets, const uint64_t* x,
                     const uint64_t* z, const int32_t* ny, const double* cre, const double* cim, double* out_per_rank) {
    int rc = check_rankset(rs);
    if (rc) return rc;
    if (!out_per_rank) return fail(VQE_ERR_INVALID, "null argument");
    if (n_ops < 0 || (n_ops > 0 && !op_offsets)) return fail(VQE_ERR_INVALID, "null offsets");
    const size_t nr = rs.r.size();
    vqe_ctx* c0 = rs.r[0];
    if (bra == VQE_BUF_PSI || ket == VQE_BUF_PSI) {
        rc = need_caller_labelling(rs);
        if (rc) return rc;
    }
    for (vqe_ctx* c : rs.r) {
        CK(cudaSetDevice(c->device));
        rc = ensure_buf(c, bra);
        if (rc) return rc;
        rc = ensure_buf(c, ket);
        if (rc) return rc;
        rc = ensure_complex(c, bra);
        if (rc == VQE_OK) rc = ensure_complex(c, ket);
        if (rc) return rc;
    }
    for (size_t k = 0; k < nr * 2 * (size_t)n_ops; ++k) out_per_rank[k] = 0.0;
    if (n_ops == 0) return VQE_OK;
    const int nl = c0->nl;
    const uint64_t full = (1ull << c0->n) - 1ull, lfull = (1ull << nl) - 1ull;
    // two tiles in shared memory: 2 * 16 * 2^tb <= 227 KB -> tb <= 12
    const int tbm = std::min(c0->tile_bits, 12);
    const int lb = std::min(c0->low_bits, std::min(tbm, nl));
    const uint64_t lowmask = (1ull << lb) - 1ull;
    const int n_terms = op_offsets[n_ops];
    for (int k = 0; k < n_terms; ++k) {
        if ((x[k] | z[k]) & ~full) return fail(VQE_ERR_INVALID, "pool term %d: mask has bits >= n_qubits", k);
        if (popc64(x[k] & z[k]) != ny[k]) return fail(VQE_ERR_INVALID, "pool term %d: ny != popcount(x&z)", k);
    }
    struct SubOp {
        int op;
        uint64_t pat, need;
    };
    std::vector<SubOp> subs;
    for (int o = 0; o < n_ops; ++o) {
        const size_t first = subs.size();
        for (int k = op_offsets[o]; k < op_offsets[o + 1]; ++k) {
            if (cre[k] == 0.0 && (!cim || cim[k] == 0.0)) continue;  // identically-zero strings: overlap exactly 0
            const uint64_t pat = x[k] >> nl;
            size_t q = first;
            while (q < subs.size() && subs[q].pat != pat) ++q;
            if (q == subs.size()) subs.push_back({o, pat, 0});
            subs[q].need |= x[k] & lfull;
        }
        for (size_t q = first; q < subs.size(); ++q)
            if (fit_low_bits(subs[q].need, lb, tbm, nl, subs[q].pat != 0) < 0)
                return fail(VQE_ERR_INVALID, "pool operator %d spans %d local X-bits, more than the tile holds", o,
                            popc64(subs[q].need));
    }
    std::vector<char> done(subs.size(), 0);
    size_t remaining = subs.size();
    while (remaining) {
        uint64_t acc = 0, pat = 0;
        std::vector<size_t> members;
        int lb_pass = lb;
        uint64_t lowmask_pass = lowmask;
        for (size_t q = 0; q < subs.size(); ++q) {
            if (done[q]) continue;
            if (members.empty()) {
                pat = subs[q].pat;
                lb_pass = fit_low_bits(subs[q].need, lb, tbm, nl, pat != 0);  // >= 0, checked above
                lowmask_pass = (1ull << lb_pass) - 1ull;
            } else if (subs[q].pat != pat) continue;
            uint64_t u = acc | subs[q].need;
            if (!plan_fits(u, lowmask_pass, tbm, nl, pat != 0)) continue;
            acc = u;
            members.push_back(q);
            done[q] = 1;
        }
        TilePlan tp = make_plan(nl, acc, tbm, lb_pass, pat);
        for (size_t q = 0; q < subs.size(); ++q) {
            if (done[q] || subs[q].pat != pat) continue;
            if ((subs[q].need & ~tp.tile_mask) == 0) {
                members.push_back(q);
                done[q] = 1;
            }
        }
        remaining -= members.size();
        std::vector<DevPoolOp> pops;
        std::vector<DevPoolTerm> pterms;
        std::vector<DevGCol> pcols;
        std::vector<DevGColEntry> pents;
        for (size_t q : members) {
            const int o = subs[q].op;
            DevPoolOp po;
            po.t_begin = (uint32_t)pterms.size();
            po.out_index = (uint32_t)o;
            po.col = 0;
            const size_t terms_mark = pterms.size();
            for (int k = op_offsets[o]; k < op_offsets[o + 1]; ++k) {
                double cr = cre[k], ci = cim ? cim[k] : 0.0;
                if ((cr == 0.0 && ci == 0.0) || (x[k] >> nl) != pat) continue;
                DevPoolTerm t;
                t.lx = plan_lx(x[k], tp);
                t.lz = plan_lz(z[k], tp);
                t.zout = plan_zout(z[k], tp);
                mul_i_pow(cr, ci, ny[k]);
                t.ar = cr;
                t.ai = ci;
                pterms.push_back(t);
            }
            po.n_terms = (uint32_t)pterms.size() - po.t_begin;
            // collapsed form (see DevPoolOp::col)
            if (po.n_terms >= 2) {
                const DevPoolTerm* te = pterms.data() + terms_mark;
                uint32_t D = 0;
                bool same = true;
                for (uint32_t k2 = 0; k2 < po.n_terms; ++k2) {
                    same = same && te[k2].lx == te[0].lx && te[k2].zout == te[0].zout;
                    D |= te[k2].lz ^ te[0].lz;
                }
                const int ne = __builtin_popcount(D);
                if (same && ne <= 5) {
                    std::vector<uint32_t> epos;
                    for (int b2 = 0; b2 < tp.tbits; ++b2)
                        if ((D >> b2) & 1u) epos.push_back((uint32_t)b2);
                    double scale = 0.0;
                    for (uint32_t k2 = 0; k2 < po.n_terms; ++k2) scale += fabs(te[k2].ar) + fabs(te[k2].ai);
                    DevGCol co;
                    memset(&co, 0, sizeof co);
                    co.zout = te[0].zout;
                    co.lz = te[0].lz;
                    co.pad[0] = te[0].lx;
                    co.nd = (uint32_t)ne;
                    for (int b2 = 0; b2 < ne; ++b2) co.dpos[b2] = ~((1u << epos[b2]) - 1u);
                    co.ent_begin = (uint32_t)pents.size();
                    co.free_log = (uint32_t)(tp.tbits - ne);
                    for (uint32_t pi = 0; pi < (1u << ne); ++pi) {
                        uint32_t pat2 = 0;
                        for (int b2 = 0; b2 < ne; ++b2)
                            if ((pi >> b2) & 1u) pat2 |= 1u << epos[b2];
                        double fr = 0.0, fi = 0.0;
                        for (uint32_t k2 = 0; k2 < po.n_terms; ++k2) {
                            const double sg = (__builtin_popcount(pat2 & (te[k2].lz ^ te[0].lz)) & 1) ? -1.0 : 1.0;
                            fr += sg * te[k2].ar;
                            fi += sg * te[k2].ai;
                        }
                        if (fabs(fr) <= 1e-15 * scale) fr = 0.0;
                        if (fabs(fi) <= 1e-15 * scale) fi = 0.0;
                        if (fr == 0.0 && fi == 0.0) continue;
                        DevGColEntry en;
                        memset(&en, 0, sizeof en);
                        en.fr = fr;
                        en.fi = fi;
                        en.pat = pat2;
                        pents.push_back(en);
                    }
                    co.n_active = (uint32_t)pents.size() - co.ent_begin;
                    pcols.push_back(co);
                    po.col = (uint32_t)pcols.size();
                    pterms.resize(terms_mark);  // the strings are folded into the table
                    po.n_terms = 0;
                }
            }
            pops.push_back(po);
        }
        const int np = (int)pops.size();
        size_t off_ops = 0, off_terms = (np * sizeof(DevPoolOp) + 15) & ~size_t(15);
        size_t off_pcols = (off_terms + pterms.size() * sizeof(DevPoolTerm) + 15) & ~size_t(15);
        size_t off_pents = off_pcols + pcols.size() * sizeof(DevGCol);
        size_t off_scat = (off_pents + pents.size() * sizeof(DevGColEntry) + 15) & ~size_t(15);
        size_t total = off_scat + tp.scat.size() * sizeof(uint64_t);
        if (tp.vbit) {
            rc = rank_barrier(rs);
            if (rc) return rc;
        }
        std::vector<char> launched(nr, 0);
        for (size_t k = 0; k < nr; ++k) {
            vqe_ctx* c = rs.r[k];
            CK(cudaSetDevice(c->device));
            TileGeom g;
            Shards sbra, sket;
            rc = make_geom(c, tp, nullptr, bra, g, sbra);
            if (rc == VQE_OK) rc = make_geom(c, tp, nullptr, ket, g, sket);
            if (rc) return rc;
            if (g.n_tiles == 0) continue;
            int gx = std::min(tile_grid(c, g.n_tiles), c->sm_count);  // 1 CTA per SM (two tiles of shared memory)
            int gy = std::max(1, std::min((np + 15) / 16, std::max(1, (c->sm_count) / gx)));
            rc = ensure_stage(c, total);
            if (rc) return rc;
            rc = ensure_partial(c, (size_t)gx * np);
            if (rc) return rc;
            rc = ensure_result(c, np);
            if (rc) return rc;
            CK(cudaStreamSynchronize(c->stream));
            memcpy(c->h_stage + off_ops, pops.data(), np * sizeof(DevPoolOp));
            memcpy(c->h_stage + off_terms, pterms.data(), pterms.size() * sizeof(DevPoolTerm));
            if (!pcols.empty()) memcpy(c->h_stage + off_pcols, pcols.data(), pcols.size() * sizeof(DevGCol));
            if (!pents.empty()) memcpy(c->h_stage + off_pents, pents.data(), pents.size() * sizeof(DevGColEntry));
            memcpy(c->h_stage + off_scat, tp.scat.data(), tp.scat.size() * sizeof(uint64_t));
            c->h2d_bytes += total;
            CK(cudaMemcpyAsync(c->d_stage, c->h_stage, total, cudaMemcpyHostToDevice, c->stream));
            g.scat = (const uint64_t*)(c->d_stage + off_scat);
            size_t smem = tile_smem(tp.tbits, 2, false);
            int threads = (int)std::min<uint64_t>(c->threads, std::max<uint64_t>(32, (1ull << tp.tbits)));
            {
                ProfScope prof(c, 3);
                k_tile_pool<<<dim3(gx, gy, 1), threads, smem, c->stream>>>(
                    sbra, sket, g, (const DevPoolOp*)(c->d_stage + off_ops), np,
                    (const DevPoolTerm*)(c->d_stage + off_terms), (const DevGCol*)(c->d_stage + off_pcols),
                    (const DevGColEntry*)(c->d_stage + off_pents), c->d_partial);
                c->launches++;
            }
            k_reduce_partials<<<np, 64, 0, c->stream>>>(c->d_partial, gx, np, np, c->d_result);
            c->launches++;
            c->d2h_bytes += np * sizeof(double2);
            CK(cudaMemcpyAsync(c->h_result, c->d_result, np * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaGetLastError());
            launched[k] = 1;
        }
        if (tp.vbit) {
            rc = rank_barrier(rs);
            if (rc) return rc;
        }
        for (size_t k = 0; k < nr; ++k) {
            if (!launched[k]) continue;
            vqe_ctx* c = rs.r[k];
            CK(cudaSetDevice(c->device));
            CK(cudaStreamSynchronize(c->stream));
            CK(cudaGetLastError());
            double* out = out_per_rank + k * 2 * (size_t)n_ops;
            for (int i = 0; i < np; ++i) {
                out[2 * pops[i].out_index] += c->h_result[i].x;
                out[2 * pops[i].out_index + 1] += c->h_result[i].y;
            }
        }
    }
    for (vqe_ctx* c : rs.r) {
        rc = vqe_shard_status(c);
        if (rc) return rc;
    }
    return VQE_OK;
}

extern "C" int vqe_pool_overlaps(vqe_ctx* c, int bra, int ket, int n_ops, const int32_t* op_offsets,
                                 const uint64_t* x, const uint64_t* z, const int32_t* ny, const double* cre,
                                 const double* cim, double* out) {
    if (!c || !out) return fail(VQE_ERR_INVALID, "null argument");
    RankSet rs;
    rs.r.push_back(c);
    return pool_impl(rs, bra, ket, n_ops, op_offsets, x, z, ny, cre, cim, out);
}
extern "C" int vqe_group_pool_overlaps(vqe_ctx* const* ranks, int n_ranks, int bra, int ket, int n_ops,
                                       const int32_t* op_offsets, const uint64_t* x, const uint64_t* z,
                                       const int32_t* ny, const double* cre, const double* cim, double* out) {
    if (!out) return fail(VQE_ERR_INVALID, "null argument");
    RankSet rs = rankset_of(ranks, n_ranks);
    std::vector<double> per((size_t)std::max(1, n_ranks) * 2 * std::max(0, n_ops), 0.0);
    int rc = pool_impl(rs, bra, ket, n_ops, op_offsets, x, z, ny, cre, cim, per.data());
    if (rc) return rc;
    for (int k = 0; k < 2 * n_ops; ++k) {
        double acc = 0.0;
        for (int r = 0; r < n_ranks; ++r) acc += per[(size_t)r * 2 * n_ops + k];  // fixed order
        out[k] = acc;
    }
    return VQE_OK;
}

// ---- reductions ---------------------------------------------------------------------------------
static int inner_bufs(vqe_ctx* c, const double2* a, const double2* b, double* out) {
    int blocks = grid_1d(c, c->n_amp, 256);
    int rc = ensure_partial(c, blocks);
    if (rc) return rc;
    k_inner<<<blocks, 256, 0, c->stream>>>(a, b, c->n_amp, c->d_partial);
    k_reduce_partials<<<1, 1024, 0, c->stream>>>(c->d_partial, blocks, 1, 1, c->d_result);
    c->launches += 2;
    c->d2h_bytes += sizeof(double2);
    CK(cudaMemcpyAsync(c->h_result, c->d_result, sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    out[0] = c->h_result[0].x;
    out[1] = c->h_result[0].y;
    return VQE_OK;
}

extern "C" int vqe_inner(vqe_ctx* c, int a, int b, double* out) {
    if (!c || !out) return fail(VQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    int rc = ensure_buf(c, a);
    if (rc) return rc;
    rc = ensure_buf(c, b);
    if (rc) return rc;
    rc = ensure_complex(c, a);
    if (rc == VQE_OK) rc = ensure_complex(c, b);
    if (rc == VQE_OK && (a == VQE_BUF_PSI) != (b == VQE_BUF_PSI)) rc = need_caller_labelling(c);  // <psi|psi> does not care
    if (rc) return rc;
    return inner_bufs(c, c->buf[a], c->buf[b], out);
}
extern "C" int vqe_norm2(vqe_ctx* c, int b, double* out) {
    double tmp[2];
    int rc = vqe_inner(c, b, b, tmp);
    if (rc) return rc;
    *out = tmp[0];
    return VQE_OK;
}
extern "C" int vqe_overlap_host(vqe_ctx* c, int b, const double* vec, double* out) {
    if (!c || !vec || !out) return fail(VQE_ERR_INVALID, "null argument");
    CK(cudaSetDevice(c->device));
    int rc = ensure_buf(c, b);
    if (rc) return rc;
    int tmp = (b == VQE_BUF_WORK) ? VQE_BUF_SIGMA : VQE_BUF_WORK;
    rc = ensure_buf(c, tmp);
    if (rc) return rc;
    rc = ensure_complex(c, b);
    if (rc == VQE_OK && b == VQE_BUF_PSI) rc = need_caller_labelling(c);
    if (rc) return rc;
    CK(cudaMemcpyAsync(c->buf[tmp], vec, c->n_amp * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    return inner_bufs(c, c->buf[tmp], c->buf[b], out);
}

// ---- exact exponential of an anti-Hermitian Pauli sum --------------------------------------------
extern "C" int vqe_apply_exp_paulisum(vqe_ctx* c, int n_terms, const uint64_t* x, const uint64_t* z,
                                      const int32_t* ny, const double* cre, const double* cim, double theta) {
    if (!c) return fail(VQE_ERR_INVALID, "ctx is null");
    CK(cudaSetDevice(c->device));
    std::vector<HTerm> terms;
    int rc = collect_terms(c, n_terms, x, z, ny, cre, cim, terms);
    if (rc) return rc;
    if (terms.empty() || theta == 0.0) return VQE_OK;
    rc = need_caller_labelling(c);
    if (rc) return rc;
    // anti-Hermitian  <=>  every coefficient purely imaginary
    bool antiherm = true, commute = true;
    for (const HTerm& t : terms)
        if (t.cr != 0.0) antiherm = false;
    for (size_t a = 0; a < terms.size() && commute; ++a)
        for (size_t b = a + 1; b < terms.size(); ++b)
            if ((popc64(terms[a].x & terms[b].z) + popc64(terms[a].z & terms[b].x)) & 1) {
                commute = false;
                break;
            }
    if (antiherm && commute) {
        // exp(theta * i*ci*P) = exp(-i (-theta ci) P): exact product of rotations
        std::vector<HostOp> ops;
        for (const HTerm& t : terms) {
            HostOp h = HostOp();
            h.kind = OP_ROT;
            h.x = t.x;
            h.z = t.z;
            h.ny = t.ny;
            double ang = -theta * t.ci;
            h.c = cos(ang);
            h.s = sin(ang);
            h.ang = ang;
            ops.push_back(h);
        }
        return run_ops(c, ops);
    }
    // general case: scaled Taylor series  psi <- (sum_m (theta A / s)^m / m!)^s psi
    vqe_paulisum ps;
    ps.device = c->device;
    rc = build_paulisum(&ps, c->n, c->nl, c->tile_bits, c->low_bits, c->threads, terms);
    if (rc == VQE_OK) rc = upload_paulisum(c, &ps);
    if (rc) { free_paulisum_device(&ps); return rc; }
    rc = ensure_buf(c, VQE_BUF_SIGMA);
    if (rc == VQE_OK) rc = ensure_buf(c, VQE_BUF_WORK);
    if (rc) { free_paulisum_device(&ps); return rc; }
    double bound = 0.0;
    for (const HTerm& t : terms) bound += sqrt(t.cr * t.cr + t.ci * t.ci);
    bound *= fabs(theta);
    int scale = std::max(1, (int)ceil(bound / 0.5));
    if (c->world > 1) {
        free_paulisum_device(&ps);
        return fail(VQE_ERR_INVALID, "exact exponential of non-commuting strings is not available on a sharded state");
    }
    rc = ensure_complex(c, VQE_BUF_PSI);
    if (rc) { free_paulisum_device(&ps); return rc; }
    c->psi_real = false;
    double2* term = c->buf[VQE_BUF_SIGMA];
    double2* next = c->buf[VQE_BUF_WORK];
    double2* psi = c->buf[VQE_BUF_PSI];
    const int threads = 256, blocks = grid_1d(c, c->n_amp, threads);
    for (int s = 0; s < scale && rc == VQE_OK; ++s) {
        CK(cudaMemcpyAsync(term, psi, c->n_amp * sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
        for (int m = 1; m <= 40; ++m) {
            rc = apply_paulisum_bufs(c, VQE_BUF_WORK, VQE_BUF_SIGMA, &ps);  // next = A term
            if (rc) break;
            // term = (theta / (scale m)) next ; psi += term
            k_axpby<<<blocks, threads, 0, c->stream>>>(term, next, c->n_amp, theta / (scale * (double)m), 0.0, 0.0, 0.0);
            k_axpby<<<blocks, threads, 0, c->stream>>>(psi, term, c->n_amp, 1.0, 0.0, 1.0, 0.0);
            c->launches += 2;
            double nrm[2];
            rc = inner_bufs(c, term, term, nrm);
            if (rc) break;
            if (nrm[0] < 1e-34) break;  // ||term|| < 1e-17
        }
    }
    free_paulisum_device(&ps);
    return rc;
}
