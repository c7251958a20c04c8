/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference hot path (CPU oracle / CPU baseline).
 *
 * Mirrors the cost structure of the reference's simulator back-end (myQLM CLinalg, reached from
 * openvqe/ucc_family/get_energy_ucc.py:42-48): one full sweep over the 2^n complex128 amplitudes per
 * Pauli rotation / gate, then one sweep per Hamiltonian term for <H>.  OpenMP over amplitudes.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load this.
 *
 * Conventions: qubit q <-> index bit n-1-q; masks are in index-bit space;
 *   P|i> = i^ny (-1)^popcount(i & z) |i ^ x>.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int orc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU baseline sets its thread count explicitly */
int orc_set_threads(int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    return omp_get_max_threads();
#else
    (void)n_threads;
    return 1;
#endif
}

void orc_basis_state(double* psi, int n, uint64_t index) {
    const uint64_t dim = 1ull << n;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < dim; ++i) { psi[2 * i] = 0.0; psi[2 * i + 1] = 0.0; }
    psi[2 * index] = 1.0;
}

/* psi <- exp(-i angle P) psi = cos(angle) psi - i sin(angle) P psi   (build_ucc_ansatz semantics,
 * SURVEY.md Appendix A V5; reference get_energy_ucc.py:42-45) */
void orc_pauli_rotation(double* psi, int n, uint64_t x, uint64_t z, int ny, double angle) {
    const uint64_t dim = 1ull << n;
    const double c = cos(angle), s = sin(angle);
    const int k = (ny + 3) & 3; /* (-i) * i^ny = i^k */
    if (x == 0) {
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < dim; ++i) {
            double sg = (__builtin_popcountll(i & z) & 1) ? -s : s;
            double re = psi[2 * i], im = psi[2 * i + 1];
            psi[2 * i] = c * re + sg * im;
            psi[2 * i + 1] = c * im - sg * re;
        }
        return;
    }
    const int hb = 63 - __builtin_clzll(x);
    const uint64_t half = dim >> 1, low = (1ull << hb) - 1;
#pragma omp parallel for schedule(static)
    for (uint64_t p = 0; p < half; ++p) {
        uint64_t i = ((p >> hb) << (hb + 1)) | (p & low), j = i ^ x;
        int pa = __builtin_popcountll(i & z) & 1, pb = __builtin_popcountll(j & z) & 1;
        double ar = psi[2 * i], ai = psi[2 * i + 1], br = psi[2 * j], bi = psi[2 * j + 1];
        double ubr, ubi, uar, uai;
        switch (k) {
            case 0: ubr = br; ubi = bi; uar = ar; uai = ai; break;
            case 1: ubr = -bi; ubi = br; uar = -ai; uai = ar; break;
            case 2: ubr = -br; ubi = -bi; uar = -ar; uai = -ai; break;
            default: ubr = bi; ubi = -br; uar = ai; uai = -ar; break;
        }
        double sb = pb ? -s : s, sa = pa ? -s : s;
        psi[2 * i] = c * ar + sb * ubr;
        psi[2 * i + 1] = c * ai + sb * ubi;
        psi[2 * j] = c * br + sa * uar;
        psi[2 * j + 1] = c * bi + sa * uai;
    }
}

/* out = <psi| P |psi>  (one term of the OBS job, reference get_energy_ucc.py:47-48) */
void orc_pauli_expectation(const double* psi, int n, uint64_t x, uint64_t z, int ny, double* out) {
    const uint64_t dim = 1ull << n;
    double re = 0.0, im = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : re, im)
    for (uint64_t i = 0; i < dim; ++i) {
        uint64_t j = i ^ x;
        double sg = (__builtin_popcountll(i & z) & 1) ? -1.0 : 1.0;
        /* conj(psi[j]) * psi[i] */
        double wr = psi[2 * j] * psi[2 * i] + psi[2 * j + 1] * psi[2 * i + 1];
        double wi = psi[2 * j] * psi[2 * i + 1] - psi[2 * j + 1] * psi[2 * i];
        re += sg * wr;
        im += sg * wi;
    }
    double r = re, m = im;
    switch (ny & 3) {
        case 1: re = -m; im = r; break;
        case 2: re = -r; im = -m; break;
        case 3: re = m; im = -r; break;
        default: break;
    }
    out[0] = re;
    out[1] = im;
}

/* sum_k c_k <P_k>, term by term (real part) */
double orc_expectation(const double* psi, int n, int n_terms, const uint64_t* x, const uint64_t* z,
                       const int32_t* ny, const double* cre, const double* cim) {
    double acc = 0.0;
    for (int k = 0; k < n_terms; ++k) {
        double e[2];
        if (cre[k] == 0.0 && (!cim || cim[k] == 0.0)) continue;
        orc_pauli_expectation(psi, n, x[k], z[k], ny[k], e);
        acc += cre[k] * e[0] - (cim ? cim[k] * e[1] : 0.0);
    }
    return acc;
}

/* E(theta) of the Trotterised UCC ansatz: basis state, ordered rotations, <H>
 * (reference get_energy_ucc.py:8-50) */
double orc_ucc_energy(double* psi, int n, uint64_t hf_index, int n_rot, const uint64_t* rx, const uint64_t* rz,
                      const int32_t* rny, const double* angle, int n_terms, const uint64_t* hx,
                      const uint64_t* hz, const int32_t* hny, const double* hre, const double* him) {
    orc_basis_state(psi, n, hf_index);
    for (int k = 0; k < n_rot; ++k)
        if (angle[k] != 0.0) orc_pauli_rotation(psi, n, rx[k], rz[k], rny[k], angle[k]);
    return orc_expectation(psi, n, n_terms, hx, hz, hny, hre, him);
}

/* one-qubit gate (row-major 2x2 complex, 8 doubles) on reference qubit q */
void orc_gate1(double* psi, int n, int q, const double* m) {
    const int b = n - 1 - q;
    const uint64_t half = 1ull << (n - 1), low = (1ull << b) - 1, bit = 1ull << b;
#pragma omp parallel for schedule(static)
    for (uint64_t p = 0; p < half; ++p) {
        uint64_t i = ((p >> b) << (b + 1)) | (p & low), j = i | bit;
        double ar = psi[2 * i], ai = psi[2 * i + 1], br = psi[2 * j], bi = psi[2 * j + 1];
        psi[2 * i] = m[0] * ar - m[1] * ai + m[2] * br - m[3] * bi;
        psi[2 * i + 1] = m[0] * ai + m[1] * ar + m[2] * bi + m[3] * br;
        psi[2 * j] = m[4] * ar - m[5] * ai + m[6] * br - m[7] * bi;
        psi[2 * j + 1] = m[4] * ai + m[5] * ar + m[6] * bi + m[7] * br;
    }
}

void orc_cnot(double* psi, int n, int control, int target) {
    const int b = n - 1 - target;
    const uint64_t cbit = 1ull << (n - 1 - control);
    const uint64_t half = 1ull << (n - 1), low = (1ull << b) - 1, bit = 1ull << b;
#pragma omp parallel for schedule(static)
    for (uint64_t p = 0; p < half; ++p) {
        uint64_t i = ((p >> b) << (b + 1)) | (p & low), j = i | bit;
        if (i & cbit) {
            double tr = psi[2 * i], ti = psi[2 * i + 1];
            psi[2 * i] = psi[2 * j];
            psi[2 * i + 1] = psi[2 * j + 1];
            psi[2 * j] = tr;
            psi[2 * j + 1] = ti;
        }
    }
}

/* gate list: kind 0=X 1=H 2=RX 3=RY 4=RZ 5=CNOT (myQLM conventions, SURVEY.md Appendix A V9;
 * reference circuit.py:13-106 executed gate by gate) */
void orc_apply_gates(double* psi, int n, int n_gates, const int32_t* kind, const int32_t* q0, const int32_t* q1,
                     const double* angle) {
    const double r2 = 0.70710678118654752440;
    for (int k = 0; k < n_gates; ++k) {
        double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        double c = cos(0.5 * angle[k]), s = sin(0.5 * angle[k]);
        switch (kind[k]) {
            case 0: m[2] = 1; m[4] = 1; break;
            case 1: m[0] = r2; m[2] = r2; m[4] = r2; m[6] = -r2; break;
            case 2: m[0] = c; m[3] = -s; m[5] = -s; m[6] = c; break;
            case 3: m[0] = c; m[2] = -s; m[4] = s; m[6] = c; break;
            case 4: m[0] = c; m[1] = -s; m[6] = c; m[7] = s; break;
            case 5: orc_cnot(psi, n, q0[k], q1[k]); continue;
            default: continue;
        }
        orc_gate1(psi, n, q0[k], m);
    }
}

/* sigma = sum_k c_k P_k psi  (reference fermionic_adapt_vqe.py:114: sig = H.dot(psi)) */
void orc_apply_paulisum(const double* psi, double* out, int n, int n_terms, const uint64_t* x, const uint64_t* z,
                        const int32_t* ny, const double* cre, const double* cim) {
    const uint64_t dim = 1ull << n;
    memset(out, 0, dim * 16);
    for (int k = 0; k < n_terms; ++k) {
        double cr = cre[k], ci = cim ? cim[k] : 0.0;
        if (cr == 0.0 && ci == 0.0) continue;
        double r = cr, m = ci;
        switch (ny[k] & 3) {
            case 1: cr = -m; ci = r; break;
            case 2: cr = -r; ci = -m; break;
            case 3: cr = m; ci = -r; break;
            default: break;
        }
        const uint64_t xk = x[k], zk = z[k];
#pragma omp parallel for schedule(static)
        for (uint64_t i = 0; i < dim; ++i) {
            uint64_t j = i ^ xk;
            double sg = (__builtin_popcountll(i & zk) & 1) ? -1.0 : 1.0;
            double vr = sg * psi[2 * i], vi = sg * psi[2 * i + 1];
            out[2 * j] += cr * vr - ci * vi;
            out[2 * j + 1] += cr * vi + ci * vr;
        }
    }
}

/* out[2k], out[2k+1] = <bra| A_k |ket> for every pool operator (reference fermionic_adapt_vqe.py:67-73 /
 * qubit_adapt_vqe.py:145-149: one matvec + one dot per operator) */
void orc_pool_overlaps(const double* bra, const double* ket, double* work, int n, int n_ops,
                       const int32_t* offsets, const uint64_t* x, const uint64_t* z, const int32_t* ny,
                       const double* cre, const double* cim, double* out) {
    const uint64_t dim = 1ull << n;
    for (int o = 0; o < n_ops; ++o) {
        int t0 = offsets[o], t1 = offsets[o + 1];
        orc_apply_paulisum(ket, work, n, t1 - t0, x + t0, z + t0, ny + t0, cre + t0, cim ? cim + t0 : 0);
        double re = 0.0, im = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : re, im)
        for (uint64_t i = 0; i < dim; ++i) {
            re += bra[2 * i] * work[2 * i] + bra[2 * i + 1] * work[2 * i + 1];
            im += bra[2 * i] * work[2 * i + 1] - bra[2 * i + 1] * work[2 * i];
        }
        out[2 * o] = re;
        out[2 * o + 1] = im;
    }
}
