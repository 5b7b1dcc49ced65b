/*
 * diffqc_b200.h — C ABI of the B200-native replacement for diffquantum's evolution hot path.
 *
 * Everything the Python host layer (diffquantum_b200/) binds with ctypes is declared here and
 * nowhere else.  Plain pointers, sizes and scalars only; no C++ or torch types cross this line.
 * Complex numbers are interleaved (re, im) IEEE float64 pairs ("c128").  Matrices are row-major.
 * Every function returns 0 on success and a negative dq_status on failure; the message for the
 * calling thread's last failure is dq_last_error().  No C++ exception crosses the ABI.
 *
 * Which reference interface each entry point replaces (paths relative to the reference tree):
 *
 *   dq_dense_set_H      diffqc.set_H            diffqc.cc:43-73    (+ pybind11 list casters, stl.h:129-142)
 *   dq_dense_trotter    diffqc.trotter + f_u    diffqc.cc:95-135, 173-205
 *   dq_pulse_f_u_table  f_u / my_expit / bspline diffqc.cc:75-135  (host only)
 *   dq_dense_evolve     SimulatorPlain.trotter  sim_plain.py:119-153   (solver hook, sim_plain.py:43)
 *   dq_dense_evolve_many  forward runs of compute_energy_grad_FD / train_fidelity  sim_plain.py:322-351, 441-449
 *   dq_dense_grad       compute_energy_grad_MC  sim_plain.py:186-220   (prefix + 2*n_H shifted suffixes)
 *   dq_dense_grad_probs / dq_dense_outcome_probs  stochastic_measure  sim_plain.py:101-117 (outcome distributions; draws on the host)
 *   dq_dense_grad_times the same + generate_u   sim_plain.py:52-99,186-220 (pulse rows evaluated on the device)
 *   dq_dense_train      train_energy            sim_plain.py:245-305   (whole loop on the device, dim <= 16)
 *   dq_ising_train      train_energy            sim_plain.py:245-305   (whole loop on the device, Pauli-term problems)
 *   dq_ising_*          the same two paths for Pauli-term (MaxCut/QAOA) Hamiltonians that the dense
 *                       nested-list API cannot express beyond n~13 (demo_maxcut.py:19-85 builds
 *                       them with np.kron) — step semantics of diffqc.cc:155-164.
 *
 * Qubit/bit convention is the reference's: qubit j of n is bit (n-1-j) of the basis index
 * (demo_maxcut.py:49-57, sim_plain.py:477-482).
 */
#ifndef DIFFQC_B200_H
#define DIFFQC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum dq_status {
    DQ_OK = 0,
    DQ_ERR_INVALID = -1,   /* bad argument (sizes, indices, NULL) — the reference has UB here */
    DQ_ERR_CUDA = -2,      /* CUDA runtime failure, message carries cudaGetErrorString */
    DQ_ERR_STATE = -3,     /* call order (e.g. trotter before set_H, diffqc.cc:21-25 globals) */
    DQ_ERR_UNSUPPORTED = -4,
    DQ_ERR_NOMEM = -5
} dq_status;

typedef struct dq_context dq_context;   /* one per process x device; owns stream + workspaces */
typedef struct dq_ising dq_ising;       /* a structured (Pauli-term) problem bound to a context */

/* ---- library / context ------------------------------------------------------------------ */
const char* dq_version(void);                       /* "dev", as diffqc.__version__ (diffqc.cc:227) */
const char* dq_last_error(void);                    /* thread-local, never NULL */
int dq_device_count(int* count);
int dq_context_create(int device, dq_context** out);
int dq_context_destroy(dq_context* ctx);
int dq_context_synchronize(dq_context* ctx);
/* cudaStream_t of the context as an integer, so the host can record torch/CUDA events on it. */
int dq_context_stream(dq_context* ctx, uint64_t* stream_out);
/* kernels launched by this context since creation (bench.py's gpu_launches claim). */
int dq_context_launch_count(dq_context* ctx, uint64_t* count_out);

/* ---- dense path (n <= 10 qubits, D = dim <= 1024): live `exact` step semantics --------------
 * psi <- expm(-i dt (H0 + sum_h u_h(t_k) H_h)) psi per step  (sim_plain.py:135-150, diffqc.cc:190-200),
 * evaluated on the device with FP64 tensor-core (DMMA) complex matmuls. */

/* diffqc.set_H: H0 [dim*dim] c128, Hs [n_H*dim*dim] c128; channels flattened:
 * chan_counts[h] channels for term h, each 4 doubles {unused, omega, w, idx} (diffqc.cc:108-111).
 * func_type 0 = Legendre, otherwise the B-spline bump basis (diffqc.cc:25,115-125). */
int dq_dense_set_H(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs,
                   const int32_t* chan_counts, const double* channels, double duration,
                   int func_type);

/* diffqc.trotter: vv [2][n_param][n_basis] (A then B coefficients); psi0/psi_out [dim] c128 on the
 * HOST.  n_steps = (int)(per_step*(|T-T0|+1)), dt = (T-T0)/n_steps, t accumulated (diffqc.cc:182-199).
 * u_out, if not NULL, receives the [n_steps][n_H] envelope values f_u (for parity tests). */
int dq_dense_trotter(dq_context* ctx, const double* psi0, double T0, double T, int per_step,
                     const double* vv, int n_param, int n_basis, double* psi_out, double* u_out);

/* The pulse envelope f_u of diffqc.cc:95-135 alone, on the HOST (no context, no device): u_out[i][h] = f_u(h, ts[i], vv)
 * for the channel table of dq_dense_set_H.  This is the routine dq_dense_trotter evaluates its step grid with; it is
 * exported so that parity tests can compare it with the reference's own compiled f_u (oracle/_ref/libfu.so). */
int dq_pulse_f_u_table(int n_H, const int32_t* chan_counts, const double* channels, double duration,
                       int func_type, const double* vv, int n_param, int n_basis, int n_t, const double* ts,
                       double* u_out);

/* Solver-hook form: the host has already evaluated u[k][h] (Python closures, sim_plain.py:81-98).
 * mode 0 = exact (live reference code), 1 = split (per-term product, diffqc.cc:155-164).
 * psi_in / psi_out: [batch][dim] c128 host buffers; all batch members share H and u. */
int dq_dense_evolve(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs,
                    const double* u, int n_steps, double dt, int mode, int batch,
                    const double* psi_in, double* psi_out);

/* n_traj independent trajectories, each with its OWN start ket, step count, dt and pulse rows (u packed [sum_b steps_b][n_H]):
 * the forward runs of the finite-difference comparator (compute_energy_grad_FD, sim_plain.py:322-351: one evolution per
 * perturbed coefficient) and of train_fidelity's state batch (:441-449).  energies_out [n_traj] = Re<psi|M|psi> (needs M),
 * psi_out [n_traj][dim] c128; either may be NULL. */
int dq_dense_evolve_many(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* M, int n_traj,
                         const double* psi_in, const int32_t* steps, const double* dts, const double* u, int mode,
                         double* energies_out, double* psi_out);

/* Batched stochastic parameter-shift samples on a dense problem (sim_plain.py:186-220).
 * For each sample b: phi = U(prefix_b) psi0; for each term i and sign +/-:
 * ket = U(suffix_b) (I +/- i r H_i) phi / sqrt(1+r^2); energies[b][i][0|1] = Re <ket|M|ket> (+, -).
 * u_prefix / u_suffix are packed [sum_b steps_b][n_H] tables with per-sample dt. */
int dq_dense_grad(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs,
                  const double* M, const double* psi0, double r, int n_samples,
                  const int32_t* prefix_steps, const double* prefix_dt, const double* u_prefix,
                  const int32_t* suffix_steps, const double* suffix_dt, const double* u_suffix,
                  int mode, double* energies_out);

/* Shot-sampling support (SimulatorPlain.stochastic_measure, sim_plain.py:101-117, called at :202-203,212-213,278-279): the same
 * batch of shifted trajectories, but instead of Re<ket|M|ket> the device returns, for every shifted ket and every measurement
 * basis m, the outcome distribution distr[j] = |<e_mj|ket>|^2 that stochastic_measure builds at :105-109 and hands to
 * np.random.choice (:112; the draws stay on the host, in the reference's order).  bases [n_meas][dim][dim] c128: component x of
 * eigenvector j of basis m at bases[m][j][x] (sim.Pauli_M[m][2][1][j]).  probs_out [n_samples][n_H][2][n_meas][dim]. */
int dq_dense_grad_probs(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* psi0, double r,
                        int n_samples, const int32_t* prefix_steps, const double* prefix_dt, const double* u_prefix,
                        const int32_t* suffix_steps, const double* suffix_dt, const double* u_suffix, int mode, int n_meas,
                        const double* bases, double* probs_out);
/* The same distributions for n_kets host kets [n_kets][dim] c128 (the final state of train_energy, :278-279).
 * probs_out [n_kets][n_meas][dim]. */
int dq_dense_outcome_probs(dq_context* ctx, int dim, int n_kets, const double* kets, int n_meas, const double* bases,
                           double* probs_out);

/* The same batch from SAMPLE TIMES: the device evaluates the step grids' pulse rows itself -- the B-spline ansatz of the Python
 * twin, u_i(t) = omega_i (2 sigma(sum_j coeff[i][j] phi_j(t/T)) - 1) on the grid of sim_plain.py:123-150 (sim_plain.py:52-99),
 * in the reference's operation order -- so nothing but s_list [n_samples] and coeff [n_H][n_basis] crosses the bus.  Resident
 * engine only (dim <= 16; DQ_ERR_UNSUPPORTED otherwise: use dq_dense_grad with host tables).  u_out, if not NULL, receives the
 * device-built table [prefix rows of all samples | suffix rows of all samples][n_H] (parity tests against the host table). */
int dq_dense_grad_times(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* M,
                        const double* psi0, double r, int n_samples, const double* s_list, double T, int per_step,
                        const double* coeff, const double* omegas, int n_basis, int mode, double* energies_out,
                        double* u_out);

/* Device-resident SimulatorPlain.train_energy (sim_plain.py:245-305) for dim <= 16 and the B-spline ansatz: n_epoch epochs of
 * (full evolution + energy, K stochastic gradient samples, their mean, torch-style Adam) enqueued without a host round trip.
 * s_all [n_epoch][K]: the sample times, drawn by the caller from the reference's stream (one np.random.uniform() * T per
 * sample, sim_plain.py:167).  coeff_inout [n_H][n_basis]: start coefficients in, trained coefficients out.  e0: the
 * observable's lowest eigenvalue (the reference recomputes it every epoch, :294).  losses_out [n_epoch] = loss_energy - e0
 * as train_energy logs it; final_state_out [dim] c128 (may be NULL) = the state of the last epoch's full evolution (:303). */
int dq_dense_train(dq_context* ctx, int dim, const double* H0, int n_H, const double* Hs, const double* M,
                   const double* psi0, const double* omegas, double T, int per_step, int n_basis, double* coeff_inout,
                   int n_epoch, int K, const double* s_all, double lr, double beta1, double beta2, double eps, double r,
                   double e0, int mode, double* losses_out, double* final_state_out);

/* Counters of the last dense call: "gemm_flops" (real flops issued to the DMMA GEMM), "strategy"
 * (0 block-Taylor, 1 per-step propagator, 2 chained propagator, 3 resident warp-per-trajectory engine, dim <= 16),
 * "squarings", "degree", "kernel_ms" (strategy 3: device time of its launches, CUDA events). */
int dq_dense_last_stat(dq_context* ctx, const char* name, double* value);
/* "strategy": -1 automatic (resident engine for dim <= 16, else by flop count), or force 0/1/2/3 (parity tests cover all);
 * "small_mma": 1 (default) the shifted kets of a sample run on the FP64 tensor cores in the resident engine, 0 the DFMA kernel. */
int dq_dense_set_option(dq_context* ctx, const char* name, int64_t value);

/* ---- structured path: H(t) = c0 + sum_e (w_e + u_e(t)) Z_a Z_b + sum_q u_q(t) X_q ------------
 * One product-formula step = one diagonal phase exp(-i(angle_c + sum_e angle_e z_a z_b)) followed by
 * exp(-i angle_q X_q) on every qubit (term order of diffqc.cc:155-164: H0, ZZ controls, X controls).
 * The HOST evaluates pulses and supplies per-step ANGLES (already multiplied by dt):
 *   row layout  [ angle_c | angle_e for e < n_zz | angle_q for q < n_qubits ]   (1 + n_zz + n doubles)
 */
int dq_ising_create(dq_context* ctx, int n_qubits, int n_zz, const int32_t* zz_pairs /*[n_zz][2]*/,
                    const double* m_zz /*[n_zz] or NULL*/, double m_const,
                    const double* m_diag /*[2^n] host, overrides m_zz when not NULL*/,
                    dq_ising** out);
int dq_ising_destroy(dq_ising* p);
/* Tunables: "ket_group" (states co-resident in L2 per launch), "engine" (0 generic = one kernel per term group, 1 fused TMA pass engine = default when 12 <= n <= 20),
 * "step" (0 = per-term product step, diffqc.cc:155-164; 1 = the reference's live exact step
 * exp(-i dt H(t_k)) psi as a matrix-free scaled Taylor series, sim_plain.py:135-150 with :147; generic engine),
 * "linear" (1, fused v2 only: evolve one shifted ket per term plus the unshifted suffix state and obtain the
 * other sign from ket- = 2 a/sqrt(1+r^2) - ket+; same outputs to rounding, about half the trajectory-steps),
 * "time_launches" (1: bracket every pass-kernel launch with CUDA events, read back via dq_ising_last_stat). */
int dq_ising_set_option(dq_ising* p, const char* name, int64_t value);
int dq_ising_get_info(dq_ising* p, const char* name, int64_t* value);

/* Evolve `batch` states through `n_steps` rows of `angles`.  psi is [batch][2^n] c128 in the
 * REFERENCE index order; psi_is_device selects host or device pointers (device pointers come from
 * torch.Tensor.data_ptr()).  energies_out (host, [batch]) may be NULL; psi_out may be NULL when only
 * energies are wanted.  psi_in == NULL means the uniform superposition (demo_maxcut.py:12-17). */
int dq_ising_evolve(dq_ising* p, int batch, int n_steps, const double* angles,
                    const void* psi_in, void* psi_out, int psi_is_device, double* energies_out);

/* Batched gradient samples.  For each sample b: prefix rows evolve psi0 to phi_b; then for every
 * shift term i (kind 0: ZZ pair index, kind 1: X on qubit index) and sign s in (+,-):
 *   ket = U(suffix rows of b) exp(+/- i atan(r) P_i) phi_b   [== (I +/- i r P_i) phi / sqrt(1+r^2)]
 *   energies_out[b][i][0|1] = <ket|M|ket>.
 * prefix_angles / suffix_angles are packed row tables, sample b occupying prefix_steps[b] /
 * suffix_steps[b] consecutive rows.  psi0 == NULL means the uniform superposition (host c128 otherwise).
 * All pointers are HOST pointers; copies are part of the call. */
int dq_ising_grad(dq_ising* p, int n_samples, const int32_t* prefix_steps,
                  const double* prefix_angles, const int32_t* suffix_steps,
                  const double* suffix_angles, int n_shift, const int32_t* shift_kind,
                  const int32_t* shift_index, double r, const double* psi0, double* energies_out);

/* Same work with inputs already staged on the device by dq_ising_grad_stage (bench.py times this
 * as the HBM-resident `value`; dq_ising_grad is the host-buffer `e2e` path). */
int dq_ising_grad_stage(dq_ising* p, int n_samples, const int32_t* prefix_steps,
                        const double* prefix_angles, const int32_t* suffix_steps,
                        const double* suffix_angles, int n_shift, const int32_t* shift_kind,
                        const int32_t* shift_index, double r, const double* psi0);
int dq_ising_grad_run_staged(dq_ising* p);                      /* asynchronous on the ctx stream */
int dq_ising_grad_fetch(dq_ising* p, double* energies_out);     /* synchronises, copies D2H */

/* Shot-sampling support on the structured path (stochastic_measure, sim_plain.py:101-117, for Z-string observables as
 * demo_maxcut.py:47-65 builds them): <Z_a Z_b> of every ZZ pair of the problem in every shifted ket of the batch,
 * zz_out [n_samples][n_shift][2][n_zz]; the outcome distribution of a pair collapses to P(-1) = (1 - <ZZ>) / 2 and the draws stay
 * on the host.  The kets themselves are needed, so this call runs one kernel per term group (not the fused pass engine). */
int dq_ising_grad_pairs(dq_ising* p, int n_samples, const int32_t* prefix_steps, const double* prefix_angles,
                        const int32_t* suffix_steps, const double* suffix_angles, int n_shift, const int32_t* shift_kind,
                        const int32_t* shift_index, double r, const double* psi0, double* zz_out);
/* The same expectations for `batch` given states (host or device c128[batch][2^n], reference bit order): zz_out [batch][n_zz]. */
int dq_ising_pair_expect(dq_ising* p, int batch, const void* psi, int psi_is_device, double* zz_out);

/* Device-resident SimulatorPlain.train_energy (sim_plain.py:245-305) for a Pauli-term problem and the B-spline ansatz: n_epoch
 * epochs of (full evolution + energy, K stochastic gradient samples, their mean, torch-style Adam) enqueued on the context stream
 * with no host round trip of data in between; the device evaluates the pulses and angle rows itself (sim_plain.py:52-99 and the
 * row layout of dq_ising_evolve).  Controls: term i is Z_a Z_b of pair term_index[i] (term_kind 0) or X of qubit term_index[i]
 * (term_kind 1), ZZ controls first; drift H0 = h0_const + sum_e h0_zz[e] Z_a Z_b (h0_zz may be NULL).  s_all [n_epoch][K]: sample
 * times drawn by the caller from the reference's stream (np.random.uniform() * T each, :167).  coeff_inout [n_terms][n_basis].
 * e0: the observable's lowest eigenvalue (min of its diagonal; the reference recomputes it densely every epoch, :294).
 * psi0 (host c128[2^n], reference bit order) or NULL = uniform superposition.  losses_out [n_epoch] = loss_energy - e0;
 * final_state_out (host c128[2^n], may be NULL) = state of the last epoch's full evolution (:304); device_ms_out (may be NULL)
 * = device time of the whole loop (CUDA events).  Step semantics and engine follow the handle's options. */
int dq_ising_train(dq_ising* p, int n_terms, const int32_t* term_kind, const int32_t* term_index, const double* omegas,
                   const double* h0_zz, double h0_const, double T, int per_step, int n_basis, double* coeff_inout,
                   int n_epoch, int K, const double* s_all, double lr, double beta1, double beta2, double eps, double r,
                   double e0, const double* psi0, double* losses_out, double* final_state_out, double* device_ms_out);

/* Counters of the last run: "steps" (trajectory-steps executed), "launches", "alg_bytes",
 * "pass_kernel_ms" / "pass_kernel_launches" (event-timed fused pass kernel; needs time_launches=1). */
int dq_ising_last_stat(dq_ising* p, const char* name, double* value);

/* ---- one state distributed over ranks on its high-order index bits (BASELINE configs[4]) -------------
 * A rank owns 2^L consecutive amplitudes (device pointer, interleaved c128); the global basis index of
 * local x is (high_bits << L) | x.  pair_bits are the CURRENT physical bit positions (0 = least
 * significant bit of the global index) of each ZZ pair; the host tracks them across the global<->local
 * qubit swaps it performs with an NCCL all-to-all (diffquantum_b200/distributed.py).  Asynchronous on
 * the context stream except dq_slice_energy.  Step semantics: diffqc.cc:155-164. */
int dq_slice_fill_uniform(dq_context* ctx, void* psi_dev, int L, int n_total);
/* psi[x] *= exp(-i (angles[0] + sum_e angles[1+e] z_a z_b)) */
int dq_slice_phase(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                   const int32_t* pair_bits, const double* angles);
/* exp(-i theta X) on local bit `bit` < L */
int dq_slice_rx(dq_context* ctx, void* psi_dev, int L, int bit, double theta);
/* exp(-i thetas[k] X) on `count` distinct local bits in as few passes over the slice as their positions allow (the
 * rotations of one product-formula step commute, diffqc.cc:155-164): up to 12 bits per read + write. */
int dq_slice_rx_many(dq_context* ctx, void* psi_dev, int L, int count, const int32_t* bits, const double* thetas);
/* dq_slice_phase followed by dq_slice_rx_many as ONE call: when the rotations start with a contiguous 12-bit tile pass the
 * phase rides on that pass (applied to the tile in shared memory before the first rotation round), one pass over the slice
 * less per product-formula step. */
int dq_slice_phase_rx_many(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                           const int32_t* pair_bits, const double* angles, int count, const int32_t* bits, const double* thetas);
/* dq_slice_rx_many whose LAST pass also performs the global<->local qubit exchange: instead of writing the slice back in place
 * it stores element x at peer_recv[x >> (L-g)][(rank << (L-g)) | (x & (2^(L-g) - 1))] -- exactly where an all_to_all_single of
 * the 2^g contiguous chunks would put it -- through peer memory (NVLink P2P, or CUDA IPC mappings of the peers' receive
 * buffers: dq_ipc_export / dq_ipc_open).  peer_recv[rank] is this rank's own receive buffer.  The caller synchronises the
 * ranks (a barrier after the pass) before anybody reads its receive buffer. */
int dq_slice_rx_many_scatter(dq_context* ctx, void* psi_dev, int L, int count, const int32_t* bits, const double* thetas,
                             int g, int rank, void* const* peer_recv);
/* The whole local part of a step -- phase, every local rotation, exchange -- in one call (dq_slice_phase_rx_many + scatter). */
int dq_slice_phase_rx_many_scatter(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                                   const int32_t* pair_bits, const double* angles, int count, const int32_t* bits,
                                   const double* thetas, int g, int rank, void* const* peer_recv);
/* One product-formula step of a slice with rotations still owed to the PREVIOUS step (diffqc.cc:155-164 over a step boundary):
 *   [exp(-i pre_thetas[k] X) on pre_bits]  [phase of this step]  [exp(-i thetas[k] X) on bits]
 * When one tile of the pass plan holds every pre bit, all three ride on ONE pass over the slice (TMA tile kernel: rotation
 * rounds, phase and rotation rounds on the tile in shared memory).  A distributed step uses it to fold the rotations of the
 * qubits that became local in the last exchange into the first pass of the next step (5 -> 3 launches per step at n = 32 on
 * 8 GPUs); n_pre = 0 is dq_slice_phase_rx_many.  The _scatter form ends with the exchange (see dq_slice_rx_many_scatter). */
int dq_slice_step(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz, const int32_t* pair_bits,
                  const double* angles, int n_pre, const int32_t* pre_bits, const double* pre_thetas, int count,
                  const int32_t* bits, const double* thetas);
int dq_slice_step_scatter(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                          const int32_t* pair_bits, const double* angles, int n_pre, const int32_t* pre_bits,
                          const double* pre_thetas, int count, const int32_t* bits, const double* thetas, int g, int rank,
                          void* const* peer_recv);
/* n_steps product-formula steps of a slice that needs no exchange (every rotated bit local), chained: step p leaves the
 * rotations of one tile set undone and the first pass of step p + 1 applies them, the phase of step p + 1 and that set's
 * rotations of step p + 1 -- (tile sets - 1) passes per step instead of (tile sets), i.e. ONE read + write of the state per
 * step for n <= 21 and two for n <= 30.  angles: row k at angles + k * ld_angles = [c | n_zz pair angles]; thetas: row k at
 * thetas + k * ld_thetas = one angle per entry of bits.  The loop of sim_plain.py:135-150 in the product form of
 * diffqc.cc:155-164. */
int dq_slice_evolve_steps(dq_context* ctx, void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                          const int32_t* pair_bits, int count, const int32_t* bits, int n_steps, const double* angles,
                          int64_t ld_angles, const double* thetas, int64_t ld_thetas);
/* The launches dq_slice_evolve_steps would make for n_steps steps on these bits and pairs, written down instead of made (host
 * only, no device: planning introspection for callers and tests).  9 int32 per launch: step, tile bits T, contiguous low bits lo,
 * rotations owed to the previous step, 1 if the pass carries the phase, rotations of this step, 1 if it carries an exchange,
 * the tile's physical-bit mask (low, high word); T = 0: a stand-alone phase pass.  assume_tma: plan as on a device where the
 * TMA tile kernel is available.  At most rows_cap launches are written, *n_rows_out is the full count. */
int dq_slice_plan(int L, int n_total, int n_zz, const int32_t* pair_bits, int count, const int32_t* bits, int n_steps,
                  int assume_tma, int32_t* rows_out, int64_t rows_cap, int64_t* n_rows_out);
/* The same for ONE step of a distributed slice (dq_slice_step / dq_slice_step_scatter): rotations owed on pre_bits, the phase,
 * rotations on bits, and -- scatter_g > 0 -- the exchange with 2^scatter_g ranks on the last launch. */
int dq_slice_plan_step(int L, int n_total, int n_zz, const int32_t* pair_bits, int n_pre, const int32_t* pre_bits, int count,
                       const int32_t* bits, int scatter_g, int assume_tma, int32_t* rows_out, int64_t rows_cap, int64_t* n_rows_out);
/* CUDA IPC plumbing for the above: a 64-byte handle + byte offset for a device pointer (the handle names the allocation the
 * pointer lives in), and the mapping of such a handle in another process (same or peer device). */
int dq_ipc_export(dq_context* ctx, void* dev_ptr, void* handle64_out, uint64_t* offset_out);
int dq_ipc_open(dq_context* ctx, const void* handle64, uint64_t offset, void** ptr_out);
int dq_ipc_close(dq_context* ctx, void* ptr, uint64_t offset);
/* this rank's part of <psi| m_const + sum_e m_zz[e] Z_a Z_b |psi> (sim_plain.py:205,215,281) */
int dq_slice_energy(dq_context* ctx, const void* psi_dev, int L, uint64_t high_bits, int n_total, int n_zz,
                    const int32_t* pair_bits, const double* m_zz, double m_const, double* partial_out);

/* ---- micro-benchmarks used by bench.py/profiles to calibrate the roofline --------------------- */
/* kind: 0 = device copy GB/s over `bytes`, 1 = FP64 FMA TFLOP/s, 2 = L2-resident read+write GB/s. */
int dq_microbench(dq_context* ctx, int kind, int64_t bytes, int iters, double* result);

#ifdef __cplusplus
}
#endif
#endif /* DIFFQC_B200_H */
