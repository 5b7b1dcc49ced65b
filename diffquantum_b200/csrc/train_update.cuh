// Gradient assembly + Adam update shared by the two device-resident training loops (dense_train.cu, ising_train.cu).
//   grad[i][j] = mean_k ps_k[i] dDdv_k[i][j],  ps = (1 + r^2) / (2 r) (E- - E+)                       sim_plain.py:220,227
//   dDdv[i][j] = omega_i 2 sigma(A_i) (1 - sigma(A_i)) phi_j(s/T),  A_i = sum_j c_ij phi_j(s/T)        sim_plain.py:169-184 (closed form)
//   torch.optim.Adam, single-tensor path: exp_avg.lerp_(g, 1 - b1); exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2);
//   denom = sqrt(exp_avg_sq) / sqrt(bc2) + eps; param += -(lr / bc1) exp_avg / denom                   sim_plain.py:266,291-292
#pragma once
#include "common.cuh"

namespace dq {

__device__ __forceinline__ double train_bump(double x, double l, double r, double norm_factor) {
    if (x >= r || x <= l) return 0.0;                       // open support, sim_plain.py:62
    return __dmul_rn(__dadd_rn(x, -l), __dadd_rn(x, -r)) / norm_factor;
}

// one thread per coefficient (i, j); a single block of >= n_H n_basis threads
static __global__ void k_train_update(const double* __restrict__ energies, const double* __restrict__ s_vals, int K, double* __restrict__ coeff,
                               double* __restrict__ m1, double* __restrict__ m2, const double* __restrict__ omegas, double T,
                               int n_H, int n_basis, const double* __restrict__ bl, const double* __restrict__ br, double norm_factor,
                               double r, double beta1, double beta2, double eps, double step_size, double bc2_sqrt,
                               const double* __restrict__ e_full, double e0, double* __restrict__ losses, int epoch) {
    const int idx = threadIdx.x;
    const int i = idx / n_basis, j = idx % n_basis;
    double newc = 0.0;
    if (idx < n_H * n_basis) {
        double g = 0.0;
        for (int k = 0; k < K; ++k) {
            const double x = s_vals[k] / T;
            double a = 0.0, phij = 0.0;
            for (int jj = 0; jj < n_basis; ++jj) {
                const double ph = train_bump(x, bl[jj], br[jj], norm_factor);
                a = __dadd_rn(a, __dmul_rn(coeff[i * n_basis + jj], ph));
                if (jj == j) phij = ph;
            }
            const double sg = 1.0 / (1.0 + exp(-a));
            const double dudc = omegas[i] * 2.0 * sg * (1.0 - sg) * phij;
            const double* e = energies + ((size_t)k * n_H + i) * 2;
            const double ps = (1.0 + r * r) / 2.0 / r * (e[1] - e[0]);
            g += ps * dudc;
        }
        g /= (double)K;
        const double a1 = m1[idx] + (1.0 - beta1) * (g - m1[idx]);
        const double a2 = m2[idx] * beta2 + (1.0 - beta2) * g * g;
        m1[idx] = a1;
        m2[idx] = a2;
        const double denom = sqrt(a2) / bc2_sqrt + eps;
        newc = coeff[idx] - step_size * (a1 / denom);
    }
    __syncthreads();                                         // every thread has read the old coefficients of its row
    if (idx < n_H * n_basis) coeff[idx] = newc;
    if (idx == 0) losses[epoch] = *e_full - e0;              // loss_energy - M.eigenenergies()[0], sim_plain.py:281,294
}

}  // namespace dq
