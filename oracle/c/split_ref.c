/*
 * Plain-C (OpenMP) restatement of the reference's per-term product step and of the shifted-ket
 * estimator for Pauli-term (MaxCut) problems.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py):
 * used as the multi-core CPU baseline of bench.py and as a second checker; never linked into or
 * called from the product (diffquantum_b200/).
 *
 * What it restates (paths relative to the reference tree):
 *   oc_evolve_split   diffqc.cc:155-164 (and sim_plain.py:139,142): per step exp(-i dt H0), then
 *                     exp(-i dt u_h H_h) for h in list order.  ZZ terms and H0 are diagonal, so their
 *                     phases are accumulated in list order and applied in one sweep; an X term is the
 *                     2x2 rotation [[c,-is],[-is,c]] on bit (n-1-q) (qubit 0 = MSB: demo_maxcut.py:49-57).
 *   oc_shift_gate     (I +/- i r H_i)/sqrt(1+r^2) applied to phi: sim_plain.py:197-199.
 *   oc_energy_diag    <ket|M|ket> for diagonal M: sim_plain.py:205,215 with M = H_cost (demo_maxcut.py:60-61).
 * Cross-checked against oracle/restate.py (which is pinned to the reference's own outputs) by
 * tests/test_oracle_c.py.
 */
#include <complex.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex c128;

int oc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Launchers such as torchrun export OMP_NUM_THREADS=1 before the process starts; the CPU arm of bench.py asks for the
 * host's cores explicitly. */
void oc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* kind[h] 0: Z_qa Z_qb, 1: X_qa.  u: [n_steps][n_terms].  h0_diag: [2^n].  psi evolves in place. */
int oc_evolve_split(int n, int n_terms, const int32_t* kind, const int32_t* qa, const int32_t* qb,
                    const double* h0_diag, const double* u, int n_steps, double dt, c128* psi) {
    const size_t N = (size_t)1 << n;
    int n_zz = 0;
    for (int h = 0; h < n_terms; ++h) n_zz += kind[h] == 0;
    int* za = (int*)malloc(sizeof(int) * (n_zz ? n_zz : 1));
    int* zb = (int*)malloc(sizeof(int) * (n_zz ? n_zz : 1));
    double* th = (double*)malloc(sizeof(double) * (n_zz ? n_zz : 1));
    if (!za || !zb || !th) return -1;
    for (int k = 0; k < n_steps; ++k) {
        const double* uk = u + (size_t)k * n_terms;
        /* The diagonal factors commute with each other but not with the X rotations: the reference's
         * list order is ZZ controls before X controls (demo_maxcut.py:68-79), so all diagonal terms
         * that precede the first X term form one sweep.  Terms after an X term are handled in order. */
        int h = 0;
        int first = 1;
        while (h < n_terms || first) {
            int m = 0;
            while (h < n_terms && kind[h] == 0) {
                za[m] = n - 1 - qa[h];
                zb[m] = n - 1 - qb[h];
                th[m] = dt * uk[h];
                ++m; ++h;
            }
            if (m > 0 || first) {
                const int with_h0 = first;
#pragma omp parallel for schedule(static)
                for (ptrdiff_t x = 0; x < (ptrdiff_t)N; ++x) {
                    double a = with_h0 ? dt * h0_diag[x] : 0.0;
                    for (int e = 0; e < m; ++e) {
                        const int par = (int)(((x >> za[e]) ^ (x >> zb[e])) & 1);
                        a += par ? -th[e] : th[e];
                    }
                    const double c = cos(a), s = sin(a);
                    const double re = creal(psi[x]), im = cimag(psi[x]);
                    psi[x] = (re * c + im * s) + (im * c - re * s) * I;      /* psi * exp(-i a) */
                }
            }
            first = 0;
            while (h < n_terms && kind[h] == 1) {
                const int bit = n - 1 - qa[h];
                const double t = dt * uk[h];
                const double c = cos(t), s = sin(t);
                const size_t low = ((size_t)1 << bit) - 1;
#pragma omp parallel for schedule(static)
                for (ptrdiff_t i = 0; i < (ptrdiff_t)(N >> 1); ++i) {
                    const size_t x0 = (((size_t)i & ~low) << 1) | ((size_t)i & low);
                    const size_t x1 = x0 | ((size_t)1 << bit);
                    const double ar = creal(psi[x0]), ai = cimag(psi[x0]);
                    const double br = creal(psi[x1]), bi = cimag(psi[x1]);
                    psi[x0] = (c * ar + s * bi) + (c * ai - s * br) * I;     /* c a - i s b */
                    psi[x1] = (c * br + s * ai) + (c * bi - s * ar) * I;     /* c b - i s a */
                }
                ++h;
            }
        }
    }
    free(za); free(zb); free(th);
    return 0;
}

/* out = (phi + sign * i r P phi) / sqrt(1 + r^2), P = Z_qa Z_qb (kind 0) or X_qa (kind 1). */
void oc_shift_gate(int n, int kind, int qa, int qb, double sign, double r, const c128* phi, c128* out) {
    const size_t N = (size_t)1 << n;
    const double inv = 1.0 / sqrt(1.0 + r * r);
    const double g = sign * r;
    if (kind == 0) {
        const int ba = n - 1 - qa, bb = n - 1 - qb;
#pragma omp parallel for schedule(static)
        for (ptrdiff_t x = 0; x < (ptrdiff_t)N; ++x) {
            const double z = (((x >> ba) ^ (x >> bb)) & 1) ? -1.0 : 1.0;
            const double re = creal(phi[x]), im = cimag(phi[x]);
            out[x] = ((re - g * z * im) + (im + g * z * re) * I) * inv;
        }
    } else {
        const size_t flip = (size_t)1 << (n - 1 - qa);
#pragma omp parallel for schedule(static)
        for (ptrdiff_t x = 0; x < (ptrdiff_t)N; ++x) {
            const c128 w = phi[(size_t)x ^ flip];
            const double re = creal(phi[x]), im = cimag(phi[x]);
            out[x] = ((re - g * cimag(w)) + (im + g * creal(w)) * I) * inv;
        }
    }
}

double oc_energy_diag(int n, const double* m_diag, const c128* psi) {
    const size_t N = (size_t)1 << n;
    double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
    for (ptrdiff_t x = 0; x < (ptrdiff_t)N; ++x) {
        const double re = creal(psi[x]), im = cimag(psi[x]);
        acc += m_diag[x] * (re * re + im * im);
    }
    return acc;
}
