// C-ABI entry points of the structured (Pauli-term) path: problem handle, single/batched
// evolution and the batched stochastic parameter-shift sample driver (sim_plain.py:186-220).
#include <string.h>
#include <math.h>
#include "ising.cuh"

using dq::c128;

namespace {

int check_rows(const dq_ising* p, const double* rows, int64_t n_rows, const char* what) {
    DQ_REQUIRE(n_rows == 0 || rows != nullptr, "%s: NULL angle table", what);
    for (int64_t i = 0; i < n_rows * p->row_len; ++i)
        DQ_REQUIRE(isfinite(rows[i]), "%s: non-finite angle at flat index %lld", what, (long long)i);
    return DQ_OK;
}

// the fused engines implement the product-formula step only; the exact step runs on the generic engine
bool use_fused(const dq_ising* p) { return p->step_mode == 0 && p->engine >= 1 && !p->want_pairs && dq::fused_supported(p); }

// n > 20 (the persistent pass engine plans 12 <= n <= 20): the product step runs on the TMA tile passes of slice.cu, chained
// over the step boundaries -- one pass per step at n = 21, two at n <= 30 -- instead of n + 1 per-term kernels.
bool use_slice_passes(const dq_ising* p) { return p->step_mode == 0 && p->engine >= 1 && !p->want_pairs && p->n > 20; }

// One product-formula step sequence on ONE state of n qubits held whole on this device (L = n, no high bits).
int slice_evolve(dq_ising* p, c128* psi, const double* h_rows, int n_steps) {
    std::vector<int32_t> pair_bits(2 * (size_t)std::max(1, p->n_zz)), xbits(p->n);
    for (int e = 0; e < p->n_zz; ++e) { pair_bits[2 * e] = p->pa[e]; pair_bits[2 * e + 1] = p->pb[e]; }
    for (int q = 0; q < p->n; ++q) xbits[q] = p->bitpos[q];
    // chained passes: (tile sets - 1) read + writes of the state per step (dq_slice_evolve_steps)
    return dq_slice_evolve_steps(p->ctx, psi, p->n, 0, p->n, p->n_zz, pair_bits.data(), p->n, xbits.data(), n_steps, h_rows,
                                 p->row_len, h_rows + 1 + p->n_zz, p->row_len);
}

// Batched gradient samples for n > 20: per sample the prefix state, then the 2 n_shift shifted kets in chunks that fit the
// device (a ket is 16 * 2^n bytes: 1 GiB at n = 26), each evolved by the slice passes, energies by the generic reduction.
int slice_grad_run(dq_ising* p) {
    auto& s = p->st;
    cudaStream_t st = p->ctx->stream;
    const size_t N = p->dim();
    const int kets = 2 * s.n_shift;
    DQ_REQUIRE(!p->host_rows_a.empty() || s.prefix_off[s.n_samples] == 0, "slice passes: host angle rows missing");
    DQ_TRY(p->phi.reserve(N * sizeof(c128)));
    size_t free_b = 0, total_b = 0;
    DQ_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t have = p->states.cap + free_b / 2;              // what the chunk buffer may grow to
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)kets, have / (N * sizeof(c128))));
    DQ_TRY(p->states.reserve((size_t)chunk * N * sizeof(c128)));
    c128* phi = p->phi.as<c128>();
    for (int b = 0; b < s.n_samples; ++b) {
        if (s.uniform_psi0) DQ_TRY(dq::gen_fill_uniform(p, phi, 1));
        else DQ_CUDA(cudaMemcpyAsync(phi, s.psi0.p, N * sizeof(c128), cudaMemcpyDeviceToDevice, st));
        DQ_TRY(slice_evolve(p, phi, p->host_rows_a.data() + s.prefix_off[b] * p->row_len, s.prefix_steps[b]));
        for (int k0 = 0; k0 < kets; k0 += chunk) {
            const int cnt = std::min(chunk, kets - k0);
            DQ_TRY(dq::gen_fanout(p, phi, p->states.as<c128>(), cnt, p->shift_desc.as<dq::ShiftDesc>() + k0, s.r));
            for (int g = 0; g < cnt; ++g)
                DQ_TRY(slice_evolve(p, p->states.as<c128>() + (size_t)g * N, p->host_rows_b.data() + s.suffix_off[b] * p->row_len,
                                    s.suffix_steps[b]));
            DQ_TRY(dq::gen_energy(p, p->states.as<c128>(), cnt, p->energies.as<double>() + (size_t)b * kets + k0));
        }
    }
    return DQ_OK;
}

// ---- qubit -> bit layout ---------------------------------------------------------------------------------
// Reference order is bit n-1-q for qubit q (np.kron order, demo_maxcut.py:53-57).  The fused engine keeps two
// sets of five index bits in registers (J_L and J_H, see Geo<> in ising_fused.cu and dq::fused_j_sets);
// a ZZ pair with both ends inside one set costs a 32-entry table multiply per amplitude (the `aj` factor).
// Automatic layout: put two disjoint independent sets of the ZZ graph on those bits, keep every other qubit
// in reference order.  Falls back to the reference order when no such sets are found.
bool choose_layout(const dq_ising* p, int* bitpos) {
    const int n = p->n;
    for (int q = 0; q < n; ++q) bitpos[q] = n - 1 - q;
    if (p->layout_mode == 0 || n < 12 || n > 20) return false;
    int jl[5], jh[5];
    dq::fused_j_sets(n, jl, jh);
    std::vector<unsigned> adj(n, 0u);
    for (int e = 0; e < p->n_zz; ++e) {
        adj[p->qa[e]] |= 1u << p->qb[e];
        adj[p->qb[e]] |= 1u << p->qa[e];
    }
    auto independent = [&](const int* pos) {
        unsigned set = 0;
        for (int i = 0; i < 5; ++i) set |= 1u << (n - 1 - pos[i]);     // qubit sitting on bit `pos` in reference order
        for (int q = 0; q < n; ++q)
            if (((set >> q) & 1u) && (adj[q] & set)) return false;
        return true;
    };
    if (independent(jl) && independent(jh)) return false;               // the reference order is already good
    // J_L and J_H may share one position (bit 2 for n <= 19): the qubit sitting there belongs to both sets
    int shared_pos = -1;
    for (int i = 0; i < 5; ++i)
        for (int k = 0; k < 5; ++k)
            if (jl[i] == jh[k]) shared_pos = jl[i];
    unsigned long long rng = 0x9E3779B97F4A7C15ull;
    std::vector<int> order(n);
    for (int attempt = 0; attempt < 4000; ++attempt) {
        for (int q = 0; q < n; ++q) order[q] = q;
        for (int q = n - 1; q > 0; --q) {
            rng = rng * 6364136223846793005ull + 1442695040888963407ull;
            std::swap(order[q], order[(int)((rng >> 33) % (unsigned)(q + 1))]);
        }
        unsigned used = 0, sets[2] = {0u, 0u};
        int shared_q = -1;
        if (shared_pos >= 0) {
            shared_q = order[0];
            sets[0] = sets[1] = used = 1u << shared_q;
        }
        bool ok = true;
        for (int s = 0; s < 2 && ok; ++s) {
            int cnt = shared_q >= 0 ? 1 : 0;
            for (int k = 0; k < n && cnt < 5; ++k) {
                const int q = order[k];
                if (((used >> q) & 1u) || (adj[q] & sets[s])) continue;
                sets[s] |= 1u << q;
                used |= 1u << q;
                ++cnt;
            }
            ok = cnt == 5;
        }
        if (!ok) continue;
        unsigned taken_pos = 0;
        if (shared_q >= 0) { bitpos[shared_q] = shared_pos; taken_pos |= 1u << shared_pos; }
        for (int s = 0; s < 2; ++s) {
            const int* pos = s == 0 ? jl : jh;
            int i = 0;
            for (int q = n - 1; q >= 0; --q) {                          // ascending bit = descending qubit, as in the reference order
                if (!((sets[s] >> q) & 1u) || q == shared_q) continue;
                if (pos[i] == shared_pos) ++i;
                bitpos[q] = pos[i];
                taken_pos |= 1u << pos[i];
                ++i;
            }
        }
        int pos = 0;
        for (int q = n - 1; q >= 0; --q) {
            if ((used >> q) & 1u) continue;
            while ((taken_pos >> pos) & 1u) ++pos;
            bitpos[q] = pos++;
        }
        return true;
    }
    for (int q = 0; q < n; ++q) bitpos[q] = n - 1 - q;
    return false;
}

// (re)build everything that depends on the layout: pair endpoints, the observable diagonal, engine plans
int apply_layout(dq_ising* p) {
    DQ_TRY(p->ctx->set_device());
    DQ_CUDA(cudaStreamSynchronize(p->ctx->stream));
    dq::fused_release(p);
    p->st.valid = false;
    p->identity_layout = !choose_layout(p, p->bitpos);
    {
        unsigned seen = 0;                                              // a layout must be a permutation of the bits
        for (int q = 0; q < p->n; ++q) seen |= 1u << p->bitpos[q];
        if (p->n <= 30 && seen != ((1u << p->n) - 1u)) {
            for (int q = 0; q < p->n; ++q) p->bitpos[q] = p->n - 1 - q;
            p->identity_layout = true;
        }
    }
    const int n_zz = p->n_zz;
    p->pa.resize(n_zz);
    p->pb.resize(n_zz);
    std::vector<int2> pr(n_zz);
    for (int e = 0; e < n_zz; ++e) {
        p->pa[e] = p->bitpos[p->qa[e]];
        p->pb[e] = p->bitpos[p->qb[e]];
        pr[e] = make_int2(p->pa[e], p->pb[e]);
    }
    DQ_TRY(p->pairs_dev.reserve((n_zz ? n_zz : 1) * sizeof(int2)));
    DQ_TRY(p->mdiag.reserve(p->dim() * sizeof(double)));
    if (n_zz) DQ_CUDA(cudaMemcpy(p->pairs_dev.p, pr.data(), n_zz * sizeof(int2), cudaMemcpyHostToDevice));
    if (!p->m_diag_host.empty()) {
        // caller's table is in reference order
        if (p->identity_layout) {
            DQ_CUDA(cudaMemcpy(p->mdiag.p, p->m_diag_host.data(), p->dim() * sizeof(double), cudaMemcpyHostToDevice));
        } else {
            DQ_TRY(p->mdiag_ref.reserve(p->dim() * sizeof(double)));
            DQ_CUDA(cudaMemcpy(p->mdiag_ref.p, p->m_diag_host.data(), p->dim() * sizeof(double), cudaMemcpyHostToDevice));
            DQ_TRY(dq::gen_permute_real_in(p, p->mdiag_ref.as<double>(), p->mdiag.as<double>()));
            DQ_CUDA(cudaStreamSynchronize(p->ctx->stream));
            p->mdiag_ref.release();
        }
    } else {
        DQ_TRY(dq::gen_build_mdiag(p, p->m_zz_host.data(), p->m_const_host));
    }
    return DQ_OK;
}

}  // namespace

namespace dq {

// Everything of a staged gradient batch except the angle rows: step counts and row offsets, the shifted-ket descriptors,
// the initial state, buffer sizes.  The rows are written into p->rows_a / p->rows_b by the caller (H2D copy in
// dq_ising_grad_stage, a device kernel in the device-resident training loop).
int stage_meta(dq_ising* p, int n_samples, const int32_t* prefix_steps, const int32_t* suffix_steps, int n_shift,
               const int32_t* shift_kind, const int32_t* shift_index, double r, const double* psi0) {
    DQ_REQUIRE(p, "NULL problem");
    DQ_REQUIRE(n_samples >= 1 && n_shift >= 1, "dq_ising_grad: n_samples=%d n_shift=%d", n_samples, n_shift);
    DQ_REQUIRE(prefix_steps && suffix_steps && shift_kind && shift_index, "dq_ising_grad: NULL table");
    DQ_REQUIRE(r > 0 && isfinite(r), "dq_ising_grad: r must be positive");
    DQ_TRY(p->ctx->set_device());
    auto& s = p->st;
    s.valid = false;
    s.exact_bound = -1.0;
    s.scaled_sample.clear();
    s.n_samples = n_samples;
    s.n_shift = n_shift;
    s.r = r;
    s.prefix_steps.assign(prefix_steps, prefix_steps + n_samples);
    s.suffix_steps.assign(suffix_steps, suffix_steps + n_samples);
    s.shift_kind.assign(shift_kind, shift_kind + n_shift);
    s.shift_index.assign(shift_index, shift_index + n_shift);
    s.prefix_off.resize(n_samples + 1);
    s.suffix_off.resize(n_samples + 1);
    s.prefix_off[0] = s.suffix_off[0] = 0;
    for (int b = 0; b < n_samples; ++b) {
        DQ_REQUIRE(prefix_steps[b] >= 0 && suffix_steps[b] >= 0, "dq_ising_grad: negative step count (sample %d)", b);
        s.prefix_off[b + 1] = s.prefix_off[b] + prefix_steps[b];
        s.suffix_off[b + 1] = s.suffix_off[b] + suffix_steps[b];
    }
    for (int i = 0; i < n_shift; ++i) {
        if (shift_kind[i] == 0)
            DQ_REQUIRE(shift_index[i] >= 0 && shift_index[i] < p->n_zz, "dq_ising_grad: shift %d: ZZ pair %d of %d", i, shift_index[i], p->n_zz);
        else if (shift_kind[i] == 1)
            DQ_REQUIRE(shift_index[i] >= 0 && shift_index[i] < p->n, "dq_ising_grad: shift %d: qubit %d of %d", i, shift_index[i], p->n);
        else
            DQ_REQUIRE(false, "dq_ising_grad: shift %d: unknown kind %d", i, shift_kind[i]);
    }
    const int64_t np = s.prefix_off[n_samples], ns = s.suffix_off[n_samples];
    cudaStream_t st = p->ctx->stream;
    const size_t rb = p->row_len * sizeof(double);
    DQ_TRY(p->rows_a.reserve((np ? np : 1) * rb));
    DQ_TRY(p->rows_b.reserve((ns ? ns : 1) * rb));
    // shifted-ket descriptors, order [i][+,-]
    std::vector<dq::ShiftDesc> h(2 * n_shift);
    for (int i = 0; i < n_shift; ++i)
        for (int sg = 0; sg < 2; ++sg) {
            dq::ShiftDesc d;
            d.kind = shift_kind[i];
            if (d.kind == 0) { d.b0 = p->pa[shift_index[i]]; d.b1 = p->pb[shift_index[i]]; }
            else { d.b0 = p->bitpos[shift_index[i]]; d.b1 = 0; }
            d.sign = sg == 0 ? +1.0 : -1.0;
            h[2 * i + sg] = d;
        }
    DQ_TRY(p->shift_desc.reserve(h.size() * sizeof(dq::ShiftDesc)));
    DQ_CUDA(cudaMemcpyAsync(p->shift_desc.p, h.data(), h.size() * sizeof(dq::ShiftDesc), cudaMemcpyHostToDevice, st));
    s.uniform_psi0 = psi0 == nullptr;
    if (psi0) {
        DQ_TRY(s.psi0.reserve(p->dim() * sizeof(c128)));
        if (p->identity_layout) {
            DQ_CUDA(cudaMemcpyAsync(s.psi0.p, psi0, p->dim() * sizeof(c128), cudaMemcpyHostToDevice, st));
        } else {
            DQ_TRY(p->io.reserve(p->dim() * sizeof(c128)));
            DQ_CUDA(cudaMemcpyAsync(p->io.p, psi0, p->dim() * sizeof(c128), cudaMemcpyHostToDevice, st));
            DQ_TRY(dq::gen_permute_in(p, p->io.as<c128>(), s.psi0.as<c128>(), 1));
        }
    }
    DQ_TRY(p->energies.reserve((size_t)n_samples * 2 * n_shift * sizeof(double)));
    DQ_CUDA(cudaStreamSynchronize(st));          // `h` goes out of scope
    return DQ_OK;
}

// generic engine only: (cos, sin) of the X angles of the rows now in p->rows_a / p->rows_b
int stage_trig(dq_ising* p) {
    if (use_fused(p)) return DQ_OK;
    auto& s = p->st;
    const int64_t np = s.prefix_off[s.n_samples], ns = s.suffix_off[s.n_samples];
    DQ_TRY(p->trig_a.reserve((np ? np : 1) * p->n * sizeof(double2)));
    DQ_TRY(p->trig_b.reserve((ns ? ns : 1) * p->n * sizeof(double2)));
    DQ_TRY(dq::gen_trig(p, p->rows_a.as<double>(), np, p->trig_a.as<double2>()));
    DQ_TRY(dq::gen_trig(p, p->rows_b.as<double>(), ns, p->trig_b.as<double2>()));
    return DQ_OK;
}

bool engine_is_fused(const dq_ising* p) { return use_fused(p); }

}  // namespace dq

extern "C" {

int dq_ising_create(dq_context* ctx, int n_qubits, int n_zz, const int32_t* zz_pairs,
                    const double* m_zz, double m_const, const double* m_diag, dq_ising** out) {
    DQ_REQUIRE(ctx && out, "dq_ising_create: NULL argument");
    *out = nullptr;
    DQ_REQUIRE(n_qubits >= 1 && n_qubits <= 30, "dq_ising_create: n_qubits=%d outside [1,30]", n_qubits);
    DQ_REQUIRE(n_zz >= 0 && (n_zz == 0 || zz_pairs), "dq_ising_create: bad ZZ pair list");
    DQ_TRY(ctx->set_device());
    dq_ising* p = new dq_ising();
    p->ctx = ctx;
    p->n = n_qubits;
    p->n_zz = n_zz;
    p->row_len = 1 + n_zz + n_qubits;
    for (int e = 0; e < n_zz; ++e) {
        int a = zz_pairs[2 * e], b = zz_pairs[2 * e + 1];
        if (a < 0 || b < 0 || a >= n_qubits || b >= n_qubits || a == b) {
            dq::set_error("dq_ising_create: pair %d = (%d,%d) invalid for %d qubits", e, a, b, n_qubits);
            delete p;
            return DQ_ERR_INVALID;
        }
        p->qa.push_back(a);
        p->qb.push_back(b);
    }
    if (m_diag) p->m_diag_host.assign(m_diag, m_diag + p->dim());
    p->m_zz_host.assign(n_zz ? n_zz : 1, 0.0);
    if (m_zz) for (int e = 0; e < n_zz; ++e) p->m_zz_host[e] = m_zz[e];
    p->m_const_host = m_const;
    const int s = apply_layout(p);
    if (s != DQ_OK) { dq_ising_destroy(p); return s; }
    *out = p;
    return DQ_OK;
}

int dq_ising_destroy(dq_ising* p) {
    if (!p) return DQ_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    dq::fused_release(p);
    dq::DevBuf* bufs[] = {&p->mdiag, &p->mdiag_ref, &p->pairs_dev, &p->states, &p->phi, &p->rows_a, &p->rows_b, &p->trig_a,
                          &p->trig_b, &p->energies, &p->scratch, &p->io, &p->shift_desc, &p->st.psi0, &p->exact_diag, &p->exact_t0,
                          &p->exact_t1, &p->train, &p->pair_out};
    for (auto* b : bufs) b->release();
    delete p;
    return DQ_OK;
}

int dq_ising_set_option(dq_ising* p, const char* name, int64_t value) {
    DQ_REQUIRE(p && name, "NULL argument");
    if (!strcmp(name, "engine")) {
        DQ_REQUIRE(value >= 0 && value <= 1, "engine must be 0 (generic, one kernel per term group) or 1 (fused pass engine)");
        p->engine = (int)value;
    } else if (!strcmp(name, "item_tiles_log2")) {
        DQ_REQUIRE(value >= 0 && value <= 6, "item_tiles_log2 out of range");
        p->item_tiles_log2 = (int)value;
    } else if (!strcmp(name, "grid_per_sm")) {
        p->grid_per_sm = (int)value;
    } else if (!strcmp(name, "step")) {
        DQ_REQUIRE(value == 0 || value == 1, "step must be 0 (split) or 1 (exact)");
        p->step_mode = (int)value;
    } else if (!strcmp(name, "layout")) {
        DQ_REQUIRE(value == 0 || value == 1, "layout must be 0 (reference bit order) or 1 (automatic)");
        if (p->layout_mode != (int)value) {
            p->layout_mode = (int)value;
            DQ_TRY(apply_layout(p));
        }
    } else if (!strcmp(name, "linear")) {
        p->linear = value != 0;
    } else if (!strcmp(name, "time_launches")) {
        p->time_launches = value != 0;
    } else if (!strcmp(name, "ket_group")) {
        DQ_REQUIRE(value >= 0 && value <= 1024, "ket_group out of range (0 = automatic)");
        p->ket_group = (int)value;
    } else {
        dq::set_error("dq_ising_set_option: unknown option '%s'", name);
        return DQ_ERR_INVALID;
    }
    return DQ_OK;
}

int dq_ising_get_info(dq_ising* p, const char* name, int64_t* value) {
    DQ_REQUIRE(p && name && value, "NULL argument");
    if (!strcmp(name, "engine")) *value = use_fused(p) ? 1 : (use_slice_passes(p) ? 2 : 0);   // 2: fused slice passes (n > 20)
    else if (!strcmp(name, "ket_group")) *value = dq::auto_ket_group(p);
    else if (!strcmp(name, "row_len")) *value = p->row_len;
    else if (!strcmp(name, "identity_layout")) *value = p->identity_layout ? 1 : 0;
    else if (!strcmp(name, "n_qubits")) *value = p->n;
    else { dq::set_error("dq_ising_get_info: unknown name '%s'", name); return DQ_ERR_INVALID; }
    return DQ_OK;
}

int dq_ising_last_stat(dq_ising* p, const char* name, double* value) {
    DQ_REQUIRE(p && name && value, "NULL argument");
    if (!strcmp(name, "steps")) *value = p->stat_steps;
    else if (!strcmp(name, "launches")) *value = p->stat_launches;
    else if (!strcmp(name, "alg_bytes")) *value = p->stat_alg_bytes;
    else if (!strcmp(name, "pass_kernel_ms") || !strcmp(name, "pass_kernel_launches")) {
        double ms = 0, nl = 0;
        DQ_TRY(dq::fused_launch_times(p, &ms, &nl));
        *value = !strcmp(name, "pass_kernel_ms") ? ms : nl;
    }
    else { dq::set_error("dq_ising_last_stat: unknown name '%s'", name); return DQ_ERR_INVALID; }
    return DQ_OK;
}

int dq_ising_evolve(dq_ising* p, int batch, int n_steps, const double* angles, const void* psi_in,
                    void* psi_out, int psi_is_device, double* energies_out) {
    DQ_REQUIRE(p, "NULL problem");
    DQ_REQUIRE(batch >= 1 && n_steps >= 0, "dq_ising_evolve: batch=%d n_steps=%d", batch, n_steps);
    DQ_REQUIRE(batch <= 65535, "dq_ising_evolve: batch=%d exceeds 65535 states per call (the batch is a grid dimension)", batch);
    DQ_REQUIRE(psi_out || energies_out, "dq_ising_evolve: nothing requested");
    DQ_TRY(check_rows(p, angles, n_steps, "dq_ising_evolve"));
    DQ_TRY(p->ctx->set_device());
    cudaStream_t st = p->ctx->stream;
    const size_t N = p->dim(), bytes = N * batch * sizeof(c128);
    const uint64_t l0 = p->ctx->launches;
    DQ_TRY(p->states.reserve(bytes));
    c128* d = p->states.as<c128>();
    if (!psi_in) {
        DQ_TRY(dq::gen_fill_uniform(p, d, batch));
    } else if (p->identity_layout) {
        DQ_CUDA(cudaMemcpyAsync(d, psi_in, bytes, psi_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    } else {
        DQ_TRY(p->io.reserve(bytes));
        DQ_CUDA(cudaMemcpyAsync(p->io.p, psi_in, bytes, psi_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        DQ_TRY(dq::gen_permute_in(p, p->io.as<c128>(), d, batch));
    }
    DQ_TRY(p->energies.reserve(batch * sizeof(double)));
    if (use_fused(p)) {
        DQ_TRY(dq::fused_evolve(p, d, batch, angles, n_steps, energies_out ? p->energies.as<double>() : nullptr,
                                psi_out != nullptr));
    } else if (use_slice_passes(p)) {
        for (int g = 0; g < batch; ++g) DQ_TRY(slice_evolve(p, d + (size_t)g * N, angles, n_steps));
        if (energies_out) DQ_TRY(dq::gen_energy(p, d, batch, p->energies.as<double>()));
    } else {
        if (n_steps) {
            DQ_TRY(p->rows_a.reserve((size_t)n_steps * p->row_len * sizeof(double)));
            DQ_TRY(p->trig_a.reserve((size_t)n_steps * p->n * sizeof(double2)));
            DQ_CUDA(cudaMemcpyAsync(p->rows_a.p, angles, (size_t)n_steps * p->row_len * sizeof(double),
                                    cudaMemcpyHostToDevice, st));
            if (p->step_mode == 1) {
                DQ_TRY(dq::gen_evolve_exact(p, d, batch, p->rows_a.as<double>(), angles, n_steps));
            } else {
                DQ_TRY(dq::gen_trig(p, p->rows_a.as<double>(), n_steps, p->trig_a.as<double2>()));
                DQ_TRY(dq::gen_evolve(p, d, batch, p->rows_a.as<double>(), p->trig_a.as<double2>(), n_steps));
            }
        }
        if (energies_out) DQ_TRY(dq::gen_energy(p, d, batch, p->energies.as<double>()));
    }
    if (energies_out)
        DQ_CUDA(cudaMemcpyAsync(energies_out, p->energies.p, batch * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (psi_out) {
        const c128* src = d;
        if (!p->identity_layout) {
            DQ_TRY(p->io.reserve(bytes));
            DQ_TRY(dq::gen_permute_out(p, d, p->io.as<c128>(), batch));
            src = p->io.as<c128>();
        }
        DQ_CUDA(cudaMemcpyAsync(psi_out, src, bytes, psi_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    }
    DQ_CUDA(cudaStreamSynchronize(st));
    p->stat_steps = (double)n_steps * batch;
    p->stat_alg_bytes = p->stat_steps * 2.0 * sizeof(c128) * N;
    p->stat_launches = (double)(p->ctx->launches - l0);
    return DQ_OK;
}

int dq_ising_grad_stage(dq_ising* p, int n_samples, const int32_t* prefix_steps, const double* prefix_angles,
                        const int32_t* suffix_steps, const double* suffix_angles, int n_shift,
                        const int32_t* shift_kind, const int32_t* shift_index, double r, const double* psi0) {
    DQ_TRY(dq::stage_meta(p, n_samples, prefix_steps, suffix_steps, n_shift, shift_kind, shift_index, r, psi0));
    auto& s = p->st;
    const int64_t np = s.prefix_off[n_samples], ns = s.suffix_off[n_samples];
    DQ_TRY(check_rows(p, prefix_angles, np, "dq_ising_grad(prefix)"));
    DQ_TRY(check_rows(p, suffix_angles, ns, "dq_ising_grad(suffix)"));
    cudaStream_t st = p->ctx->stream;
    const size_t rb = p->row_len * sizeof(double);
    if (np) DQ_CUDA(cudaMemcpyAsync(p->rows_a.p, prefix_angles, np * rb, cudaMemcpyHostToDevice, st));
    if (ns) DQ_CUDA(cudaMemcpyAsync(p->rows_b.p, suffix_angles, ns * rb, cudaMemcpyHostToDevice, st));
    {
        const int off_x = 1 + p->n_zz;
        s.scaled_sample.assign(n_samples, 0);
        s.scaled_ok = true;
        for (int b = 0; b < n_samples; ++b) {
            double mx = fabs(atan(r));
            for (int64_t k = s.prefix_off[b]; k < s.prefix_off[b + 1]; ++k)
                for (int q = 0; q < p->n; ++q) mx = fmax(mx, fabs(prefix_angles[k * p->row_len + off_x + q]));
            for (int64_t k = s.suffix_off[b]; k < s.suffix_off[b + 1]; ++k)
                for (int q = 0; q < p->n; ++q) mx = fmax(mx, fabs(suffix_angles[k * p->row_len + off_x + q]));
            s.scaled_sample[b] = mx <= 1.0;
            s.scaled_ok = s.scaled_ok && mx <= 1.0;
        }
    }
    DQ_TRY(dq::stage_trig(p));
    if (p->step_mode == 1 || use_slice_passes(p)) {
        p->host_rows_a.assign(prefix_angles, prefix_angles + np * p->row_len);
        p->host_rows_b.assign(suffix_angles, suffix_angles + ns * p->row_len);
    } else {
        p->host_rows_a.clear();
        p->host_rows_b.clear();
    }
    DQ_CUDA(cudaStreamSynchronize(st));          // the caller's tables may go out of scope
    s.valid = true;
    return DQ_OK;
}

int dq_ising_grad_run_staged(dq_ising* p) {
    DQ_REQUIRE(p && p->st.valid, "dq_ising_grad_run_staged: nothing staged");
    DQ_TRY(p->ctx->set_device());
    auto& s = p->st;
    const uint64_t l0 = p->ctx->launches;
    const size_t N = p->dim();
    const int kets = 2 * s.n_shift;
    double steps = 0;
    if (use_fused(p)) {
        DQ_TRY(dq::fused_grad_run(p));
    } else if (use_slice_passes(p) && !(p->host_rows_a.empty() && p->host_rows_b.empty() && s.prefix_off[s.n_samples] + s.suffix_off[s.n_samples] > 0)) {
        DQ_TRY(slice_grad_run(p));
    } else {
        cudaStream_t st = p->ctx->stream;
        DQ_TRY(p->phi.reserve(N * sizeof(c128)));
        DQ_TRY(p->states.reserve(N * kets * sizeof(c128)));
        for (int b = 0; b < s.n_samples; ++b) {
            c128* phi = p->phi.as<c128>();
            if (s.uniform_psi0) DQ_TRY(dq::gen_fill_uniform(p, phi, 1));
            else DQ_CUDA(cudaMemcpyAsync(phi, s.psi0.p, N * sizeof(c128), cudaMemcpyDeviceToDevice, st));
            if (p->step_mode == 1) {
                const bool dev_rows = s.exact_bound >= 0.0;
                DQ_TRY(dq::gen_evolve_exact(p, phi, 1, p->rows_a.as<double>() + s.prefix_off[b] * p->row_len,
                                            dev_rows ? nullptr : p->host_rows_a.data() + s.prefix_off[b] * p->row_len, s.prefix_steps[b],
                                            s.exact_bound));
                DQ_TRY(dq::gen_fanout(p, phi, p->states.as<c128>(), kets, p->shift_desc.as<dq::ShiftDesc>(), s.r));
                DQ_TRY(dq::gen_evolve_exact(p, p->states.as<c128>(), kets, p->rows_b.as<double>() + s.suffix_off[b] * p->row_len,
                                            dev_rows ? nullptr : p->host_rows_b.data() + s.suffix_off[b] * p->row_len, s.suffix_steps[b],
                                            s.exact_bound));
                DQ_TRY(dq::gen_energy(p, p->states.as<c128>(), kets, p->energies.as<double>() + (size_t)b * kets));
                if (p->want_pairs) DQ_TRY(dq::gen_pair_expect(p, p->states.as<c128>(), kets, p->pair_out.as<double>() + (size_t)b * kets * p->n_zz));
                continue;
            }
            DQ_TRY(dq::gen_evolve(p, phi, 1, p->rows_a.as<double>() + s.prefix_off[b] * p->row_len,
                                  p->trig_a.as<double2>() + s.prefix_off[b] * p->n, s.prefix_steps[b]));
            DQ_TRY(dq::gen_fanout(p, phi, p->states.as<c128>(), kets, p->shift_desc.as<dq::ShiftDesc>(), s.r));
            DQ_TRY(dq::gen_evolve(p, p->states.as<c128>(), kets, p->rows_b.as<double>() + s.suffix_off[b] * p->row_len,
                                  p->trig_b.as<double2>() + s.suffix_off[b] * p->n, s.suffix_steps[b]));
            DQ_TRY(dq::gen_energy(p, p->states.as<c128>(), kets, p->energies.as<double>() + (size_t)b * kets));
            if (p->want_pairs) DQ_TRY(dq::gen_pair_expect(p, p->states.as<c128>(), kets, p->pair_out.as<double>() + (size_t)b * kets * p->n_zz));
        }
    }
    const bool lin = p->linear && use_fused(p);
    for (int b = 0; b < s.n_samples; ++b)
        steps += s.prefix_steps[b] + (double)(lin ? s.n_shift + 1 : kets) * s.suffix_steps[b];
    p->stat_steps = steps;
    p->stat_alg_bytes = steps * 2.0 * sizeof(c128) * N;
    p->stat_launches = (double)(p->ctx->launches - l0);
    return DQ_OK;
}

int dq_ising_grad_fetch(dq_ising* p, double* energies_out) {
    DQ_REQUIRE(p && p->st.valid && energies_out, "dq_ising_grad_fetch: nothing staged or NULL output");
    DQ_TRY(p->ctx->set_device());
    DQ_CUDA(cudaMemcpyAsync(energies_out, p->energies.p, (size_t)p->st.n_samples * 2 * p->st.n_shift * sizeof(double),
                            cudaMemcpyDeviceToHost, p->ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(p->ctx->stream));
    return DQ_OK;
}

int dq_ising_grad_pairs(dq_ising* p, int n_samples, const int32_t* prefix_steps, const double* prefix_angles,
                        const int32_t* suffix_steps, const double* suffix_angles, int n_shift, const int32_t* shift_kind,
                        const int32_t* shift_index, double r, const double* psi0, double* zz_out) {
    DQ_REQUIRE(p && zz_out, "dq_ising_grad_pairs: NULL argument");
    DQ_REQUIRE(n_samples >= 1 && n_shift >= 1, "dq_ising_grad_pairs: n_samples=%d n_shift=%d", n_samples, n_shift);
    DQ_TRY(p->ctx->set_device());
    const size_t count = (size_t)n_samples * 2 * n_shift * p->n_zz;
    DQ_TRY(p->pair_out.reserve((count ? count : 1) * sizeof(double)));
    p->want_pairs = true;            // the shifted kets themselves are needed: one kernel per term group, no fused passes
    int rc = dq_ising_grad_stage(p, n_samples, prefix_steps, prefix_angles, suffix_steps, suffix_angles, n_shift, shift_kind,
                                 shift_index, r, psi0);
    if (rc == DQ_OK) rc = dq_ising_grad_run_staged(p);
    p->want_pairs = false;
    p->st.valid = false;
    DQ_TRY(rc);
    DQ_CUDA(cudaMemcpyAsync(zz_out, p->pair_out.p, count * sizeof(double), cudaMemcpyDeviceToHost, p->ctx->stream));
    DQ_CUDA(cudaStreamSynchronize(p->ctx->stream));
    return DQ_OK;
}

int dq_ising_pair_expect(dq_ising* p, int batch, const void* psi, int psi_is_device, double* zz_out) {
    DQ_REQUIRE(p && psi && zz_out && batch >= 1, "dq_ising_pair_expect: bad argument");
    DQ_TRY(p->ctx->set_device());
    cudaStream_t st = p->ctx->stream;
    const size_t bytes = p->dim() * batch * sizeof(c128);
    DQ_TRY(p->states.reserve(bytes));
    c128* d = p->states.as<c128>();
    if (p->identity_layout) {
        DQ_CUDA(cudaMemcpyAsync(d, psi, bytes, psi_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    } else {
        DQ_TRY(p->io.reserve(bytes));
        DQ_CUDA(cudaMemcpyAsync(p->io.p, psi, bytes, psi_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        DQ_TRY(dq::gen_permute_in(p, p->io.as<c128>(), d, batch));
    }
    const size_t count = (size_t)batch * p->n_zz;
    DQ_TRY(p->pair_out.reserve((count ? count : 1) * sizeof(double)));
    DQ_TRY(dq::gen_pair_expect(p, d, batch, p->pair_out.as<double>()));
    DQ_CUDA(cudaMemcpyAsync(zz_out, p->pair_out.p, count * sizeof(double), cudaMemcpyDeviceToHost, st));
    DQ_CUDA(cudaStreamSynchronize(st));
    return DQ_OK;
}

int dq_ising_grad(dq_ising* p, int n_samples, const int32_t* prefix_steps, const double* prefix_angles,
                  const int32_t* suffix_steps, const double* suffix_angles, int n_shift, const int32_t* shift_kind,
                  const int32_t* shift_index, double r, const double* psi0, double* energies_out) {
    DQ_REQUIRE(energies_out, "dq_ising_grad: NULL output");
    DQ_TRY(dq_ising_grad_stage(p, n_samples, prefix_steps, prefix_angles, suffix_steps, suffix_angles, n_shift,
                               shift_kind, shift_index, r, psi0));
    DQ_TRY(dq_ising_grad_run_staged(p));
    return dq_ising_grad_fetch(p, energies_out);
}

}  // extern "C"
