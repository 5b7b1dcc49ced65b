// Dense path (n <= 10 qubits, dim <= 1024): live `exact` step semantics of the reference,
//   psi <- expm(-i dt (H0 + sum_h u_h(t_k) H_h)) psi      (sim_plain.py:135-150, diffqc.cc:190-200)
// and the disabled per-term product (diffqc.cc:155-164) as mode 1.
//
// Device layout: every complex matrix / ket block is PLANAR float64 — a real plane followed by an
// imaginary plane — row-major with leading dimension Dp (dim padded to a multiple of 8) or Ncp
// (ket count padded to a multiple of 8).  Planar is what the FP64 tensor-core fragments want:
// a complex product is four real DMMA (mma.sync.m8n8k4.f64) accumulations.
#pragma once
#include <array>
#include <vector>
#include "common.cuh"

namespace dq {
namespace dense {

struct Problem {
    int dim = 0, Dp = 0, n_H = 0;
    DevBuf H;                          // [(1 + n_H)][2][Dp*Dp]
    DevBuf M;                          // [2][Dp*Dp] observable (gradient entry point only)
    std::vector<double> norm1;         // induced 1-norms of H0, H_1..H_nH (scaling choice)
    std::vector<double> host_copy;     // last uploaded H0|Hs (interleaved c128), to skip identical uploads
    // diffqc.set_H pulse model (diffqc.cc:21-25 globals)
    std::vector<std::vector<std::array<double, 4>>> channels;
    double duration = 1.0;
    int func_type = 0;
    bool is_set = false;
    size_t plane() const { return (size_t)Dp * Dp; }
};

struct State {                         // per context
    Problem global_H;                  // what dq_dense_set_H stored
    Problem scratch_H;                 // what dq_dense_evolve / dq_dense_grad were last called with
    DevBuf A, P0, P1, U, K0, K1, K2, u_dev, meta, phi, out;
    DevBuf small_H, small_traj;        // resident engine (dim <= 16): [(2 + n_H)][16][16] c128 (H0, H_h, M) and descriptors
    DevBuf train;                      // device-resident training loop (dense_train.cu): coefficients, Adam state, descriptors
    double last_gemm_flops = 0;        // real flops issued to the DMMA GEMM in the last call
    int last_strategy = 0;             // 0 block-Taylor, 1 per-step propagator, 2 chained propagator, 3 resident (dim <= 16)
    int last_squarings = 0, last_degree = 0;
    double last_kernel_ms = 0;         // resident engine: device time of its launches in the last call (CUDA events)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

struct Gemm {                          // C[z] = alpha * A[z] B[z] (+ Add[z]) (+ I),  z < batch
    const double* A; long long strideA, planeA; int lda;
    const double* B; long long strideB, planeB; int ldb;
    double* C; long long strideC, planeC; int ldc;
    const double* Add;                 // same geometry as C, may be NULL
    int M, N, K, batch;
    double alpha;
    int add_identity;
};

int zgemm(dq_context* ctx, const Gemm& g);
int build_generator(dq_context* ctx, const Problem& P, int nb, const double* d_u, const long long* d_rows,
                    const double* d_scale, int k, int term, double* d_A, double* d_P, double inv_m);
int gather_blocks(dq_context* ctx, const double* src, double* dst, const int* d_order, int nb, size_t block_doubles,
                  int scatter);
int fanout(dq_context* ctx, const Problem& P, int B, const double* d_phi, int ncp_phi, double* d_K, int ncp, double r);
int energies(dq_context* ctx, const Problem& P, int B, const double* d_K, int ncp, int n_cols, double* d_out);

// Resident engine for dim <= 16 (dense_small.cu): one warp per trajectory, all steps in one launch.
struct SmallTraj {
    long long row;                     // first pulse row of the trajectory in the packed table
    double scale;                      // dt / 2^s
    double shift;                      // shift gate (I + i shift H_term) / sqrt(1 + r^2) first; 0 = none.  With NK kets per
                                       // trajectory ket g uses control term + g/2 and the sign pattern +shift, -shift, ...
    int steps;
    int src;                           // index of the start ket
    int term;                          // control index of the (first) shift gate
    int out;                           // index of the (first) result (ket or energy); ket g writes out + g
};
bool small_fits(const Problem& P);
int small_upload(dq_context* ctx, const Problem& P, const double* M);
int small_run(dq_context* ctx, const Problem& P, int mode, int s, int m, int kets_per_traj, const std::vector<SmallTraj>& traj,
              const double* d_u, const double* d_src, double* d_dst_kets, double* d_dst_energy, double inv_norm);
int small_enqueue(dq_context* ctx, const Problem& P, int mode, int s, int m, int kets_per_traj, const SmallTraj* d_traj, int n,
                  const double* d_u, const double* d_src, double* d_dst_kets, double* d_dst_energy, double inv_norm, const int* d_s);
int upload_problem(dq_context* ctx, Problem& P, int dim, const double* H0, int n_H, const double* Hs);
State* state_of(dq_context* ctx);
void release(dq_context* ctx);

}  // namespace dense
}  // namespace dq
