// Structured (Pauli-term) problem handle shared by the generic and fused engines.
#pragma once
#include "common.cuh"

struct dq_ising {
    dq_context* ctx = nullptr;
    int n = 0;                     // qubits
    int n_zz = 0;                  // ZZ pairs
    int row_len = 0;               // 1 + n_zz + n doubles per angle row
    std::vector<int> qa, qb;       // pair endpoints (qubit ids)
    int bitpos[40];                // physical bit position of qubit q in the device layout
    std::vector<int> pa, pb;       // pair endpoints as physical bit positions
    bool identity_layout = true;   // bitpos[q] == n-1-q for all q
    dq::DevBuf mdiag;              // double[2^n], physical order
    dq::DevBuf mdiag_ref;          // staging: a caller-supplied table in reference order
    std::vector<double> m_zz_host, m_diag_host;   // observable as given at creation (the layout can change later)
    double m_const_host = 0.0;
    int layout_mode = 1;           // 0 reference order (bit n-1-q), 1 automatic (fused engine: no ZZ pair inside a register set)
    dq::DevBuf pairs_dev;          // int2[n_zz] physical bit positions

    int engine = 1;                // 0 generic, 1 fused TMA engine (32 amplitudes/thread, two teams x three tile buffers per SM, default)
                                   // 2 fused v3 (16 amplitudes/thread, LDGSTS); 12 <= n <= 20
    int ket_group = 0;             // states co-resident in a fused launch (the work ring a chained launch cycles through); 0 = automatic
    int item_tiles_log2 = 0;       // fused v2: a work item is 2^k consecutive tiles (one atomic / poll / release per item); measured: 0 is best (56.0 / 54.5 / 48.3 / 28.9 samples/s for k = 0..3 at n = 20, finer items pipeline better across pass boundaries)
    int grid_per_sm = 0;           // experiment knob: CTAs per SM in the persistent grid (0 = occupancy)
    int linear = 0;                // 1: estimator by linearity (n_H + 1 suffix trajectories per sample instead of 2 n_H)
    int time_launches = 0;         // 1: CUDA-event pairs around every pass-kernel launch (bench.py's roofline)

    // work buffers
    dq::DevBuf states, phi, rows_a, rows_b, trig_a, trig_b, energies, scratch, io, shift_desc;
    dq::DevBuf exact_diag, exact_t0, exact_t1;     // exact-step workspaces (generic engine)
    dq::DevBuf train;                              // device-resident training loop (ising_train.cu)
    dq::DevBuf pair_out;                           // shot sampling: <Z_a Z_b> of every shifted ket, [n_samples][2 n_shift][n_zz]
    bool want_pairs = false;                       // the staged run also fills pair_out (generic engine: the kets must exist)
    int step_mode = 0;             // 0 split (per-term product, diffqc.cc:155-164), 1 exact (live semantics, sim_plain.py:135-150)
    std::vector<double> host_rows_a, host_rows_b;  // exact step: host copy of the staged rows (norm bound per step)

    // staged gradient batch
    struct Staged {
        bool valid = false;
        int n_samples = 0, n_shift = 0;
        double r = 0.5;
        std::vector<int> prefix_steps, suffix_steps, shift_kind, shift_index;
        std::vector<int64_t> prefix_off, suffix_off;     // row offsets
        bool uniform_psi0 = true;
        bool scaled_ok = true;                            // every |x angle| and atan(r) <= 1 rad (all samples)
        std::vector<char> scaled_sample;                  // the same per sample; empty: scaled_ok for all
        double exact_bound = -1.0;                        // >= 0: rows live on the device only; bound on every ||dt H(t_k)|| (exact step)
        dq::DevBuf psi0;                                  // physical order, when not uniform
    } st;

    double stat_steps = 0, stat_launches = 0, stat_alg_bytes = 0;

    size_t dim() const { return (size_t)1 << n; }
};

namespace dq {
// States co-resident per fused launch: as many as keep the work ring inside ~72 MiB of the 126 MB L2
// (measured with the warp-specialised kernel: 4 x 16 MiB best at n = 20 -- 72.3 vs 70.0 / 69.9 samples/s for 3 / 5 --,
// 16-20 at n = 18, 80-96 at n = 16), at least 4, at most 96 (small states need many kets to fill 148 SMs).
inline int auto_ket_group(const dq_ising* p) {
    if (p->ket_group > 0) return p->ket_group > 96 ? 96 : p->ket_group;
    const size_t state_bytes = sizeof(double2) << p->n;
    const size_t g = ((size_t)72 << 20) / state_bytes;
    return (int)(g < 4 ? 4 : (g > 96 ? 96 : g));
}
// one shifted ket of the estimator: exp(sign * i * atan(r) * P), P = Z_b0 Z_b1 (kind 0) or X_b0 (kind 1)
struct ShiftDesc { int kind; int b0; int b1; double sign; };
// generic engine (any n >= 1): one kernel per term group, used for small n and as cross-check
int gen_fill_uniform(dq_ising* p, c128* psi, int batch);
int gen_permute_in(dq_ising* p, const c128* src_ref_order, c128* dst_phys, int batch);
int gen_permute_out(dq_ising* p, const c128* src_phys, c128* dst_ref_order, int batch);
int gen_permute_real_in(dq_ising* p, const double* src_ref_order, double* dst_phys);
int gen_trig(dq_ising* p, const double* d_rows, int64_t n_rows, double2* d_trig);
int gen_evolve(dq_ising* p, c128* d_states, int batch, const double* d_rows, const double2* d_trig,
               int n_steps);
int gen_evolve_exact(dq_ising* p, c128* d_states, int batch, const double* d_rows, const double* h_rows, int n_steps,
                     double uniform_bound = -1.0);
int gen_fanout(dq_ising* p, const c128* d_phi, c128* d_kets, int n_kets, const ShiftDesc* d_desc,
               double r);
int gen_energy(dq_ising* p, const c128* d_states, int batch, double* d_out);
int gen_build_mdiag(dq_ising* p, const double* m_zz, double m_const);
int gen_pair_expect(dq_ising* p, const c128* d_states, int batch, double* d_out);   // <Z_a Z_b> of every ZZ pair, [batch][n_zz]

// staging pieces shared by dq_ising_grad_stage and the device-resident training loop (ising_api.cu)
int stage_meta(dq_ising* p, int n_samples, const int32_t* prefix_steps, const int32_t* suffix_steps, int n_shift,
               const int32_t* shift_kind, const int32_t* shift_index, double r, const double* psi0);
int stage_trig(dq_ising* p);
bool engine_is_fused(const dq_ising* p);

// fused persistent engine (n >= 12)
int fused_supported(const dq_ising* p);
void fused_j_sets(int n, int* jl, int* jh);   // physical bits held in registers around the phase, L pass and H pass
void fused_release(dq_ising* p);
int fused_launch_times(dq_ising* p, double* total_ms, double* n_launches);
int fused_grad_run(dq_ising* p);
int fused_evolve(dq_ising* p, c128* d_states, int batch, const double* h_rows, int n_steps,
                 double* d_energies, bool want_states, const double* d_rows = nullptr, int scaled_hint = -1);

}  // namespace dq
