// Fused persistent engine v3 for the structured path (12 <= n <= 20), sm_100a: 16 amplitudes per thread.
//
// Same pass algebra as ising_fused.cu — one product-formula step (diffqc.cc:155-164) is a diagonal phase D(k)
// followed by X rotations on every qubit; rotations of one step commute and D is elementwise, so the work is
// regrouped into passes that touch one bit set each (L = physical bits [0,10), H = bits [10,n)):
//
//     pass p :  [ mixer S_p of step p ]  ->  D(p+1)  ->  [ mixer S_p of step p+1 ]        S_p alternates L,H
//
// i.e. ONE global read + write of the state per step.  What changed is the shape of the CTA: the v2 engine keeps
// 32 amplitudes (128 registers) per thread, which caps an SM at 8 warps and leaves every barrier, shared-memory
// and L2 latency exposed (ncu r01: FP64 pipe 37 % busy, 12 % warp occupancy).  Here a 2^12-amplitude tile
// (64 KiB of shared memory) is owned by 256 threads x 16 amplitudes, <= 128 registers, two CTAs per SM = 16 warps.
// A thread then rotates 4 qubits per register round, so the 10 active bits of a pass take five rounds
//     A(4 bits) -> B(2 bits) -> M(4 bits, phase, 4 bits) -> B(2 bits) -> A(4 bits)
// with a shared-memory exchange between rounds (XOR-swizzled: every round's 16-byte accesses are conflict free).
// Round A is also the global layout: each thread fetches its own 16 slots with cp.async while the previous item
// finishes, and stores (or reduces <M>) straight from registers.
//
// Rotations use the scaled form a' = a - i tan(theta) b (2 DFMA per amplitude per qubit), the product of cosines is
// folded into the phase tables.  The diagonal phase never costs a per-amplitude sincos: per thread it is
//     TC[tile] * TT[thread bits] * prod_m XT_m[neighbour pattern]      (base phase, <= 10 complex multiplies)
// followed by a doubling product over the 4 register bits with per-bit factors F_k (15 + 16 complex multiplies per
// 16 amplitudes).  Tables are built per (trajectory, pass) by k_setup from the host-evaluated angle rows.
// The estimator's shift gates (sim_plain.py:197-199) ride on pass 0: a ZZ gate is one more phase factor, an X gate is
// one more butterfly; the energy <ket|M|ket> (sim_plain.py:205,215) is reduced inside the last pass.
#include <algorithm>
#include <math.h>
#include <string.h>
#include "ising.cuh"

namespace dq {
namespace f16 {

constexpr int kTileBits = 12;
constexpr int kTile = 1 << kTileBits;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRegs = 16;
constexpr int kMaxNbr = 3;
constexpr int kMaxPairs = 128;
#ifndef DQ16_CTAS_PER_SM
#define DQ16_CTAS_PER_SM 2
#endif

enum : int { F_ENERGY = 2, F_STORE = 4 };
enum : int { AP = 0, BP = 1, MP = 2, MN = 3, BN = 4, AN = 5 };       // rotation slots of a pass

struct __align__(16) PassStep {
    double2 tt[256];                // base phase over the 8 thread bits of round M (constant, cosine scale, T-T, R-x at j = 0)
    double2 aj[kRegs];              // R-R pairs, relative to j = 0
    double2 xt[8][8];               // thread bit m x pattern of its column neighbours
    double2 fr[4][8];               // register bit k x pattern of its neighbours outside R: exp(+2 i w_k)
    double2 rot[6][4];              // (1, tan) or (cos, sin) per slot and bit
    int flags;
    int type;
    unsigned long long tc_offset;   // first entry of this pass-step's column (tile id) table
};

// Launch-constant geometry of a pass type (kernel parameter space).
struct TypeGeom {
    int a, lowmask;                 // tile bit t -> physical: t < a ? t : 10 + (t - a)      (type H; L is the identity)
    int tid_lo_bits, high_end;
    int fr_pos[4][kMaxNbr], fr_msk[4][kMaxNbr];
    int xt_pos[8][kMaxNbr], xt_msk[8][kMaxNbr];
    int offA[kRegs];                // physical offset of register j in round A
};

struct TypePlan {
    TypeGeom g;
    int start;
    int aq[4], bq[2], mq[4];        // x-angle column (qubit) rotated by that slot bit, -1 = spectator
    int fr_pair[4][kMaxNbr], xt_pair[8][kMaxNbr];
    int n_col_bits;
    int n_pairs;
    int has_aj;
    // pair classes: 0 RR, 1 TT, 2 RT, 3 RC, 4 TC, 5 CC ; i0/i1 local indices (R: 0..3, T: 0..7, C: column bit)
    signed char cls[kMaxPairs], i0[kMaxPairs], i1[kMaxPairs];
};

struct KetDesc {
    const c128* src;
    c128* buf;
    const PassStep* steps;
    double* partial;
    double sigma;
    double escale;
    int n_pass;
    int shift_kind;                 // -1 none, 0 ZZ on physical bits (sb0, sb1), 1 X on sb0
    int sb0, sb1;
    int cls;                        // pass type of its pass 0
    int pad_;
};

struct SetupJob {
    long long row_pre, row_cur;     // rows of the previous / current step in the angle table, -1 = none
    int type;
    int flags;
};

struct LaunchArgs {
    const KetDesc* kets;
    const double2* tc;
    const double* mdiag;
    unsigned* counters;             // [0] next item, [1 + g] tiles done of ket g
    int n_kets;
    int max_pass;
    int tiles_log2;
    double r, ca, sa, c2a, s2a;
    TypeGeom geom[2];
};

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int insert4(int t, int p) { return (t & ((1 << p) - 1)) | ((t >> p) << (p + 4)); }
__device__ __forceinline__ int swz(int e) { return e ^ ((e >> 4) & 7); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int gather3(size_t x, const int* pos, const int* msk) {
    return (int)(((x >> pos[0]) & msk[0]) | (((x >> pos[1]) & msk[1]) << 1) | (((x >> pos[2]) & msk[2]) << 2));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// exp(-i theta X) on register bit B.  SCALED: rc = (1, tan) -> a' = a - i t b.
template <bool SCALED, int B>
__device__ __forceinline__ void rot_bit(c128 (&v)[kRegs], const double2 rc) {
#pragma unroll
    for (int j = 0; j < kRegs; ++j) {
        if (j & (1 << B)) continue;
        const c128 a = v[j], b = v[j | (1 << B)];
        if (SCALED) {
            v[j] = make_double2(fma(rc.y, b.y, a.x), fma(-rc.y, b.x, a.y));
            v[j | (1 << B)] = make_double2(fma(rc.y, a.y, b.x), fma(-rc.y, a.x, b.y));
        } else {
            v[j] = make_double2(fma(rc.y, b.y, rc.x * a.x), fma(-rc.y, b.x, rc.x * a.y));
            v[j | (1 << B)] = make_double2(fma(rc.y, a.y, rc.x * b.x), fma(-rc.y, a.x, rc.x * b.y));
        }
    }
}

// Straight-line on purpose (no branch around a butterfly block); an inactive bit holds the identity.
template <bool SCALED>
__device__ __forceinline__ void rot4(c128 (&v)[kRegs], const double2* rc, const int ov, const double2 ov_rc) {
    double2 r0 = rc[0], r1 = rc[1], r2 = rc[2], r3 = rc[3];
    if (ov == 0) r0 = ov_rc;
    if (ov == 1) r1 = ov_rc;
    if (ov == 2) r2 = ov_rc;
    if (ov == 3) r3 = ov_rc;
    rot_bit<SCALED, 0>(v, r0);
    rot_bit<SCALED, 1>(v, r1);
    rot_bit<SCALED, 2>(v, r2);
    rot_bit<SCALED, 3>(v, r3);
}
template <bool SCALED, int B0>
__device__ __forceinline__ void rot2(c128 (&v)[kRegs], const double2* rc, const int ov, const double2 ov_rc) {
    double2 r0 = rc[0], r1 = rc[1];
    if (ov == 0) r0 = ov_rc;
    if (ov == 1) r1 = ov_rc;
    rot_bit<SCALED, B0>(v, r0);
    rot_bit<SCALED, B0 + 1>(v, r1);
}

// Register sets of the three rounds, as the lowest of their 4 contiguous tile bits.
//   TYPE 0 (L): tile = physical bits [0,12): A = 6..9, B = 2..5 (rotates 4,5), M = 0..3; bits 10,11 are spectators.
//   TYPE 1 (H): tile = `a` low spectator bits + physical [10, 10+12-a): A = 4..7, B = 2..5 (rotates 2,3), M = 8..11.
// The two M sets are disjoint in physical bits for every n, so a ZZ shift gate always finds a pass type in which
// at most one of its operands is a register bit of round M.
template <int TYPE> struct Geo;
template <> struct Geo<0> { static constexpr int rA = 6, rB = 2, rM = 0, bB = 2; };
template <> struct Geo<1> { static constexpr int rA = 4, rB = 2, rM = 8, bB = 0; };

struct ItemInfo {
    unsigned item;
    int p, g, t_id;
    int valid;
    int ready;
};

__device__ __forceinline__ void decode_item(const LaunchArgs& A, unsigned item, ItemInfo& I) {
    I.item = item;
    I.t_id = (int)(item & ((1u << A.tiles_log2) - 1u));
    const unsigned rest = item >> A.tiles_log2;
    I.g = (int)(rest % (unsigned)A.n_kets);
    I.p = (int)(rest / (unsigned)A.n_kets);
}

__device__ __forceinline__ size_t tile_base(const TypeGeom& T, const int type, const int t_id) {
    if (type == 0) return (size_t)t_id << kTileBits;
    return ((size_t)(t_id & ((1 << T.tid_lo_bits) - 1)) << T.a) | ((size_t)(t_id >> T.tid_lo_bits) << T.high_end);
}
__device__ __forceinline__ size_t phys_off(const TypeGeom& T, const int type, const int e) {
    if (type == 0) return (size_t)e;
    return (size_t)(e & T.lowmask) | ((size_t)(e >> T.a) << 10);
}

// Each thread copies the 16 amplitudes it reads back in round A into the slots it reads them from.
__device__ __forceinline__ void prefetch_tile(const LaunchArgs& A, const KetDesc* __restrict__ kd, const int type,
                                              const int p, const int t_id, c128* __restrict__ tile) {
    const int tid = threadIdx.x;
    const c128* __restrict__ src = (p == 0 ? kd->src : kd->buf);
    if (type == 0) {
        const int eA = insert4(tid, Geo<0>::rA);
        src += ((size_t)t_id << kTileBits) + eA;
#pragma unroll
        for (int j = 0; j < kRegs; ++j) cp_async16(tile + swz(eA | (j << Geo<0>::rA)), src + (j << Geo<0>::rA));
    } else {
        const TypeGeom& T = A.geom[1];
        const int eA = insert4(tid, Geo<1>::rA);
        src += tile_base(T, 1, t_id) + phys_off(T, 1, eA);
#pragma unroll
        for (int j = 0; j < kRegs; ++j) cp_async16(tile + swz(eA | (j << Geo<1>::rA)), src + T.offA[j]);
    }
}

__device__ __forceinline__ void prefetch_tables(const PassStep* __restrict__ ps, PassStep* __restrict__ slot) {
    const char* s = reinterpret_cast<const char*>(ps);
    char* d = reinterpret_cast<char*>(slot);
    for (int i = threadIdx.x; i < (int)(sizeof(PassStep) / 16); i += kThreads) cp_async16(d + 16 * i, s + 16 * i);
}

struct Shared {
    ItemInfo info[2];
    double red[2][kWarps];
};

struct Pending {
    double* partial;
    double escale;
    int g;
    int slot;
};

__device__ __forceinline__ void flush_pending(const LaunchArgs& A, Shared& sh, Pending& pd) {
    if (pd.g < 0) return;
    if (pd.partial) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += sh.red[pd.slot][w];
        *pd.partial = s * pd.escale;
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    atomicAdd(&A.counters[1 + pd.g], 1u);
    pd.g = -1;
}

template <bool SCALED, bool AJ, int TYPE>
__device__ __forceinline__ void process_tile(const LaunchArgs& A, const KetDesc* __restrict__ kd, const PassStep& P,
                                             c128* __restrict__ tile, Shared& sh, const int p, const int t_id,
                                             const unsigned nxt_raw, const int nb, const unsigned total,
                                             PassStep* __restrict__ cache, const PassStep* (&cached_ps)[2],
                                             const int cb, int& next_cb, bool& next_tables_new, Pending& pd) {
    using G = Geo<TYPE>;
    const TypeGeom& T = A.geom[TYPE];
    const int tid = threadIdx.x;
    const int flags = P.flags;
    const int eA = insert4(tid, G::rA), eB = insert4(tid, G::rB), eM = insert4(tid, G::rM);
    const size_t tbase = tile_base(T, TYPE, t_id);
    const size_t xA = tbase + phys_off(T, TYPE, eA);
    const size_t xM = tbase + phys_off(T, TYPE, eM);

    // shift gate of the estimator (pass 0 only), expressed as data so that the amplitude code stays straight-line
    const int shift_kind = (p == 0) ? kd->shift_kind : -1;
    const double sigma = kd->sigma;
    int ovA = -1, ovB = -1, ovM = -1;
    int tb0 = -1, tb1 = -1;
    if (shift_kind >= 0) {
        const int s0 = kd->sb0, s1 = kd->sb1;
        if (TYPE == 0) { tb0 = s0 < kTileBits ? s0 : -1; tb1 = s1 < kTileBits ? s1 : -1; }
        else {
            tb0 = s0 < T.a ? s0 : (s0 >= 10 ? s0 - 10 + T.a : -1);
            tb1 = s1 < T.a ? s1 : (s1 >= 10 ? s1 - 10 + T.a : -1);
        }
        if (shift_kind == 1) {
            if (tb0 >= G::rA && tb0 < G::rA + 4) ovA = tb0 - G::rA;
            if (tb0 >= G::rB + G::bB && tb0 < G::rB + G::bB + 2) ovB = tb0 - G::rB - G::bB;
            if (tb0 >= G::rM && tb0 < G::rM + 4) ovM = tb0 - G::rM;
        }
    }
    // (I + i sigma r X) = exp(-i theta X) / cos(theta) with tan(theta) = -sigma r
    const double2 shift_rc = SCALED ? make_double2(1.0, -sigma * A.r) : make_double2(A.ca, -sigma * A.sa);

    const c128 phi_tc = __ldg(A.tc + P.tc_offset + (unsigned)t_id);
    c128 v[kRegs];
    // ---- round A (previous step) : own slots -> registers -----------------------------------------------
#pragma unroll
    for (int j = 0; j < kRegs; ++j) v[j] = tile[swz(eA | (j << G::rA))];
    rot4<SCALED>(v, P.rot[AP], ovA, shift_rc);
#pragma unroll
    for (int j = 0; j < kRegs; ++j) tile[swz(eA | (j << G::rA))] = v[j];
    unsigned polled = 0;
    ItemInfo nI;
    nI.valid = 0;
    nI.ready = 0;
    if (tid == 0) {
        decode_item(A, nxt_raw, nI);
        if (nxt_raw < total) {
            nI.valid = nI.p < A.kets[nI.g].n_pass;
            polled = nI.p > 0 ? ld_acquire(&A.counters[1 + nI.g]) : 0u;
        }
    }
    __syncthreads();
    // ---- round B (previous step) ---------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < kRegs; ++j) v[j] = tile[swz(eB | (j << G::rB))];
    rot2<SCALED, G::bB>(v, P.rot[BP], ovB, shift_rc);
#pragma unroll
    for (int j = 0; j < kRegs; ++j) tile[swz(eB | (j << G::rB))] = v[j];
    __syncthreads();
    if (tid == 0) flush_pending(A, sh, pd);      // previous item: its stores were ordered by the barriers above
    // ---- round M : rotations of the previous step, phase, rotations of the next step ----------------------
#pragma unroll
    for (int j = 0; j < kRegs; ++j) v[j] = tile[swz(eM | (j << G::rM))];
    rot4<SCALED>(v, P.rot[MP], ovM, shift_rc);
    {
        c128 phi = cmul(phi_tc, P.tt[tid]);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            if (T.xt_msk[m][0]) {                // launch-uniform
                c128 w = P.xt[m][gather3(xM, T.xt_pos[m], T.xt_msk[m])];
                if ((tid >> m) & 1) w.y = -w.y;
                phi = cmul(phi, w);
            }
        }
        c128 F[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) F[k] = P.fr[k][gather3(xM, T.fr_pos[k], T.fr_msk[k])];
        if (shift_kind == 0) {                   // ZZ shift gate exp(i sigma alpha z0 z1); at most one operand in R
            const int j0b = (tb0 >= G::rM && tb0 < G::rM + 4) ? tb0 - G::rM : -1;
            const int j1b = (tb1 >= G::rM && tb1 < G::rM + 4) ? tb1 - G::rM : -1;
            const double z0 = ((xM >> kd->sb0) & 1) ? -1.0 : 1.0;
            const double z1 = ((xM >> kd->sb1) & 1) ? -1.0 : 1.0;
            phi = cmul(phi, make_double2(A.ca, sigma * A.sa * z0 * z1));     // register operand at 0 (z = +1)
            const int jb = j0b >= 0 ? j0b : j1b;
            if (jb >= 0) {                       // flipping that bit multiplies by exp(-2 i sigma alpha z_other)
                const double zo = j0b >= 0 ? z1 : z0;
                const c128 f = make_double2(A.c2a, -sigma * A.s2a * zo);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k == jb) F[k] = cmul(F[k], f);
            }
        }
#pragma unroll
        for (int b3 = 0; b3 < 2; ++b3) {
            const c128 p3 = b3 ? cmul(phi, F[3]) : phi;
#pragma unroll
            for (int b2 = 0; b2 < 2; ++b2) {
                const c128 p2 = b2 ? cmul(p3, F[2]) : p3;
#pragma unroll
                for (int b1 = 0; b1 < 2; ++b1) {
                    const c128 p1 = b1 ? cmul(p2, F[1]) : p2;
                    const int j = (b3 << 3) | (b2 << 2) | (b1 << 1);
                    v[j] = cmul(v[j], p1);
                    v[j | 1] = cmul(v[j | 1], cmul(p1, F[0]));
                }
            }
        }
        if (AJ) {
#pragma unroll
            for (int j = 0; j < kRegs; ++j) v[j] = cmul(v[j], P.aj[j]);
        }
    }
    rot4<SCALED>(v, P.rot[MN], -1, shift_rc);
#pragma unroll
    for (int j = 0; j < kRegs; ++j) tile[swz(eM | (j << G::rM))] = v[j];
    if (tid == 0) {                              // publish the following item
        nI.ready = nI.valid && (nI.p == 0 || polled >= (unsigned)nI.p << A.tiles_log2);
        sh.info[nb] = nI;
    }
    __syncthreads();
    // ---- round B (next step) -----------------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < kRegs; ++j) v[j] = tile[swz(eB | (j << G::rB))];
    rot2<SCALED, G::bB>(v, P.rot[BN], -1, shift_rc);
#pragma unroll
    for (int j = 0; j < kRegs; ++j) tile[swz(eB | (j << G::rB))] = v[j];
    __syncthreads();
    // ---- round A (next step) : prefetch of the next tile into the slots just read; store or reduce ---------
#pragma unroll
    for (int j = 0; j < kRegs; ++j) v[j] = tile[swz(eA | (j << G::rA))];
    {
        const ItemInfo& N = sh.info[nb];
        next_cb = cb;
        next_tables_new = false;
        if (N.ready) {
            const KetDesc* __restrict__ nkd = A.kets + N.g;
            const PassStep* nps = nkd->steps + N.p;
            // a tile of the other pass type lands in OTHER threads' round-A slots: wait until everyone has read
            if ((((N.p + nkd->cls) & 1) != TYPE)) __syncthreads();
            if (nps != cached_ps[cb]) {
                next_cb = cb ^ 1;
                if (nps != cached_ps[next_cb]) {
                    prefetch_tables(nps, cache + next_cb);
                    cached_ps[next_cb] = nps;
                    next_tables_new = true;
                }
            }
            prefetch_tile(A, nkd, (N.p + nkd->cls) & 1, N.p, N.t_id, tile);
            cp_async_commit();
        }
    }
    rot_bit<SCALED, 0>(v, P.rot[AN][0]);
    rot_bit<SCALED, 1>(v, P.rot[AN][1]);
    rot_bit<SCALED, 2>(v, P.rot[AN][2]);
    {
        const double2 rc = P.rot[AN][3];
        const bool do_store = (flags & F_STORE) != 0, do_energy = (flags & F_ENERGY) != 0;
        c128* __restrict__ dst = kd->buf + xA;
        const double* __restrict__ md = A.mdiag + xA;
        double e = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const c128 a = v[j], b = v[j + 8];
            c128 na, nbv;
            if (SCALED) {
                na = make_double2(fma(rc.y, b.y, a.x), fma(-rc.y, b.x, a.y));
                nbv = make_double2(fma(rc.y, a.y, b.x), fma(-rc.y, a.x, b.y));
            } else {
                na = make_double2(fma(rc.y, b.y, rc.x * a.x), fma(-rc.y, b.x, rc.x * a.y));
                nbv = make_double2(fma(rc.y, a.y, rc.x * b.x), fma(-rc.y, a.x, rc.x * b.y));
            }
            const size_t o0 = TYPE == 0 ? (size_t)(j << G::rA) : (size_t)T.offA[j];
            const size_t o1 = TYPE == 0 ? (size_t)((j + 8) << G::rA) : (size_t)T.offA[j + 8];
            if (do_store) { __stcg(dst + o0, na); __stcg(dst + o1, nbv); }
            if (do_energy) {
                e = fma(__ldg(md + o0), fma(na.x, na.x, na.y * na.y), e);
                e = fma(__ldg(md + o1), fma(nbv.x, nbv.x, nbv.y * nbv.y), e);
            }
        }
        if (do_energy) {
            for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
            if ((tid & 31) == 0) sh.red[nb ^ 1][tid >> 5] = e;
        }
    }
}

template <bool SCALED, bool AJ>
__global__ void __launch_bounds__(kThreads, DQ16_CTAS_PER_SM) k_f16_passes(const __grid_constant__ LaunchArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* tile = reinterpret_cast<c128*>(smem_raw);
    PassStep* cache = reinterpret_cast<PassStep*>(smem_raw + sizeof(c128) * kTile);     // two slots
    __shared__ Shared sh;

    const int tid = threadIdx.x;
    const unsigned total = ((unsigned)A.max_pass * (unsigned)A.n_kets) << A.tiles_log2;
    const PassStep* cached_ps[2] = {nullptr, nullptr};
    int cur = 0, cb = 0;
    bool prefetched = false, tables_new = false;
    Pending pd;
    pd.g = -1;
    pd.partial = nullptr;
    pd.escale = 0.0;
    pd.slot = 0;

    if (tid == 0) {
        ItemInfo I;
        decode_item(A, atomicAdd(&A.counters[0], 1u), I);
        I.valid = I.item < total && I.p < A.kets[I.g].n_pass;
        I.ready = 0;
        sh.info[0] = I;
    }
    __syncthreads();

    for (;;) {
        const ItemInfo I = sh.info[cur];
        if (I.item >= total) break;
        if (!I.valid) {                        // ragged group: this ket has no such pass
            __syncthreads();
            if (tid == 0) {
                flush_pending(A, sh, pd);
                ItemInfo N;
                decode_item(A, atomicAdd(&A.counters[0], 1u), N);
                N.valid = N.item < total && N.p < A.kets[N.g].n_pass;
                N.ready = 0;
                sh.info[cur ^ 1] = N;
            }
            __syncthreads();
            cur ^= 1;
            prefetched = false;
            continue;
        }
        const KetDesc* __restrict__ kd = A.kets + I.g;
        if (!prefetched) {                     // cold path: wait for the dependency, then fetch tile and tables
            __syncthreads();
            if (tid == 0) {
                flush_pending(A, sh, pd);      // always before spinning: the dependency may be our own tile
                if (I.p > 0) {
                    const unsigned need = (unsigned)I.p << A.tiles_log2;
                    while (ld_acquire(&A.counters[1 + I.g]) < need) __nanosleep(32);
                }
            }
            __syncthreads();
            const PassStep* ps = kd->steps + I.p;
            tables_new = false;
            if (ps != cached_ps[cb]) {
                cb ^= 1;
                if (ps != cached_ps[cb]) {
                    prefetch_tables(ps, cache + cb);
                    cached_ps[cb] = ps;
                    tables_new = true;
                }
            }
            prefetch_tile(A, kd, (I.p + kd->cls) & 1, I.p, I.t_id, tile);
            cp_async_commit();
        }
        unsigned nxt_raw = 0;
        if (tid == 0) nxt_raw = atomicAdd(&A.counters[0], 1u);
        cp_async_wait_all();
        if (tables_new) __syncthreads();
        int next_cb = cb;
        bool next_tables_new = false;
        const PassStep& P = cache[cb];
        const int flags = P.flags;
        if (P.type == 0)
            process_tile<SCALED, AJ, 0>(A, kd, P, tile, sh, I.p, I.t_id, nxt_raw, cur ^ 1, total, cache, cached_ps, cb,
                                        next_cb, next_tables_new, pd);
        else
            process_tile<SCALED, AJ, 1>(A, kd, P, tile, sh, I.p, I.t_id, nxt_raw, cur ^ 1, total, cache, cached_ps, cb,
                                        next_cb, next_tables_new, pd);
        if (tid == 0) {
            pd.g = I.g;
            pd.partial = (flags & F_ENERGY) ? kd->partial + I.t_id : nullptr;
            pd.escale = kd->escale;
            pd.slot = cur;
        }
        prefetched = sh.info[cur ^ 1].ready != 0;
        cur ^= 1;
        cb = next_cb;
        tables_new = next_tables_new;
    }
    cp_async_wait_all();
    __syncthreads();
    if (tid == 0) flush_pending(A, sh, pd);
}

// ------------------------------------------------------------------------------------------
// table setup: one CTA per pass-step (y = column-table chunk)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double zsign(int bits, int i) { return ((bits >> i) & 1) ? -1.0 : 1.0; }

__global__ void __launch_bounds__(256) k_setup(const SetupJob* __restrict__ jobs, const double* __restrict__ rows,
                                               int row_len, int n_zz, const TypePlan* __restrict__ types,
                                               PassStep* __restrict__ steps, double2* __restrict__ tc,
                                               int n_col_bits, int scaled) {
    const SetupJob job = jobs[blockIdx.x];
    const TypePlan& T = types[job.type];
    PassStep& P = steps[blockIdx.x];
    const double* pre = job.row_pre >= 0 ? rows + job.row_pre * row_len : nullptr;
    const double* cur = job.row_cur >= 0 ? rows + job.row_cur * row_len : nullptr;
    const int off_x = 1 + n_zz;
    const unsigned long long tc_off = (unsigned long long)blockIdx.x << n_col_bits;
    const int tid = threadIdx.x;

    double ang[6][4];
    double scale = 1.0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int qa = T.aq[b], qm = T.mq[b], qb = b < 2 ? T.bq[b] : -1;
        ang[AP][b] = (pre && qa >= 0) ? pre[off_x + qa] : 0.0;
        ang[BP][b] = (pre && qb >= 0) ? pre[off_x + qb] : 0.0;
        ang[MP][b] = (pre && qm >= 0) ? pre[off_x + qm] : 0.0;
        ang[MN][b] = (cur && qm >= 0) ? cur[off_x + qm] : 0.0;
        ang[BN][b] = (cur && qb >= 0) ? cur[off_x + qb] : 0.0;
        ang[AN][b] = (cur && qa >= 0) ? cur[off_x + qa] : 0.0;
    }
    if (scaled) {
#pragma unroll
        for (int s = 0; s < 6; ++s)
#pragma unroll
            for (int b = 0; b < 4; ++b) scale *= cos(ang[s][b]);
    }
    if (blockIdx.y == 0) {
        if (tid < 24) {
            const int s = tid / 4, b = tid % 4;
            double a = 0.0;
#pragma unroll
            for (int ss = 0; ss < 6; ++ss)
#pragma unroll
                for (int bb = 0; bb < 4; ++bb)
                    if (ss == s && bb == b) a = ang[ss][bb];
            double sn, cs;
            sincos(a, &sn, &cs);
            P.rot[s][b] = scaled ? make_double2(1.0, sn / cs) : make_double2(cs, sn);
        }
        if (tid == 0) {
            P.flags = job.flags;
            P.type = job.type;
            P.tc_offset = tc_off;
        }
        {                                       // tt[tid]: every thread one entry
            double a = cur ? cur[0] : 0.0;
            if (cur)
                for (int e = 0; e < T.n_pairs; ++e) {
                    const double g = cur[1 + e];
                    switch (T.cls[e]) {
                        case 0: a += g; break;
                        case 1: a += g * zsign(tid, T.i0[e]) * zsign(tid, T.i1[e]); break;
                        case 2: a += g * zsign(tid, T.i1[e]); break;
                        default: break;
                    }
                }
            double sn, cs;
            sincos(a, &sn, &cs);
            P.tt[tid] = make_double2(scale * cs, -scale * sn);
        }
        if (tid < 16) {                         // aj[j]
            double aa = 0.0;
            if (cur)
                for (int e = 0; e < T.n_pairs; ++e)
                    if (T.cls[e] == 0) aa += cur[1 + e] * (zsign(tid, T.i0[e]) * zsign(tid, T.i1[e]) - 1.0);
            double sn, cs;
            sincos(aa, &sn, &cs);
            P.aj[tid] = make_double2(cs, -sn);
        } else if (tid >= 32 && tid < 96) {     // xt[m][pat]
            const int m = (tid - 32) >> 3, pat = (tid - 32) & 7;
            double a = 0.0;
            if (cur)
                for (int t = 0; t < kMaxNbr; ++t)
                    if (T.g.xt_msk[m][t]) a += cur[1 + T.xt_pair[m][t]] * zsign(pat, t);
            double sn, cs;
            sincos(a, &sn, &cs);
            P.xt[m][pat] = make_double2(cs, -sn);
        } else if (tid >= 96 && tid < 128) {    // fr[k][pat] = exp(+2i w_k)
            const int k = (tid - 96) >> 3, pat = (tid - 96) & 7;
            double a = 0.0;
            if (cur)
                for (int t = 0; t < kMaxNbr; ++t)
                    if (T.g.fr_msk[k][t]) a += cur[1 + T.fr_pair[k][t]] * zsign(pat, t);
            double sn, cs;
            sincos(2.0 * a, &sn, &cs);
            P.fr[k][pat] = make_double2(cs, sn);
        }
    }
    // column (tile id) table: C-C pairs + fields from register-bit neighbours
    const unsigned ncol = 1u << n_col_bits;
    for (unsigned col = blockIdx.y * blockDim.x + tid; col < ncol; col += gridDim.y * blockDim.x) {
        double a = 0.0;
        if (cur)
            for (int e = 0; e < T.n_pairs; ++e) {
                const double g = cur[1 + e];
                if (T.cls[e] == 5) a += g * ((((col >> T.i0[e]) ^ (col >> T.i1[e])) & 1) ? -1.0 : 1.0);
                else if (T.cls[e] == 3) a += g * (((col >> T.i1[e]) & 1) ? -1.0 : 1.0);
            }
        double sn, cs;
        sincos(a, &sn, &cs);
        tc[tc_off + col] = make_double2(cs, -sn);
    }
}

__global__ void k_sum_partials(const double* __restrict__ partial, int tiles, double* __restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < tiles; i += 32) acc += partial[(size_t)blockIdx.x * tiles + i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct Plan {
    bool ok = false;
    int n = 0, n_col_bits = 0, tiles_log2 = 0;
    bool has_aj = false;
    TypePlan types[2];
    int rphys[2][4];
    DevBuf d_types, jobs, steps, tc, kets, counters, partials, work, rows, phi, uniform;
    size_t smem_bytes = 0;
    int ctas_per_sm = 0;
    int counter_slots = 0, counter_cursor = 0;
    std::vector<cudaEvent_t> ev;
    int ev_used = 0;
};

static std::vector<std::pair<dq_ising*, Plan*>> g_plans;

static Plan* find_plan(const dq_ising* p) {
    for (auto& kv : g_plans)
        if (kv.first == p) return kv.second;
    return nullptr;
}

static bool build_type(const dq_ising* p, int type, TypePlan& T, int* rphys_out) {
    const int n = p->n;
    memset(&T, 0, sizeof(T));
    int a, b, rA, rB, rM, bB;
    if (type == 0) { a = kTileBits; T.start = kTileBits; b = 0; rA = 6; rB = 2; rM = 0; bB = 2; }
    else { b = n - 10; a = kTileBits - b; T.start = 10; rA = 4; rB = 2; rM = 8; bB = 0; }
    T.g.a = a;
    T.g.lowmask = (1 << a) - 1;
    T.g.tid_lo_bits = T.start - a;
    T.g.high_end = T.start + b;
    T.n_col_bits = n - kTileBits;
    auto phys_of_tile_bit = [&](int t) { return t < a ? t : T.start + (t - a); };
    auto phys_off = [&](int i) { return (i & T.g.lowmask) | ((i >> a) << T.start); };
    auto tile_bit_of_phys = [&](int pos) { return pos < a ? pos : (pos >= T.start && pos < T.start + b ? pos - T.start + a : -1); };
    for (int j = 0; j < kRegs; ++j) T.g.offA[j] = phys_off(j << rA);
    auto active = [&](int pos) { return type == 0 ? (pos < 10) : (pos >= 10); };
    int qubit_of_pos[64];
    for (int q = 0; q < n; ++q) qubit_of_pos[p->bitpos[q]] = q;
    int rphys[4], tphys[8];
    for (int i = 0; i < 4; ++i) {
        rphys[i] = phys_of_tile_bit(rM + i);
        rphys_out[i] = rphys[i];
        T.mq[i] = active(rphys[i]) ? qubit_of_pos[rphys[i]] : -1;
        const int pa = phys_of_tile_bit(rA + i);
        T.aq[i] = active(pa) ? qubit_of_pos[pa] : -1;
    }
    for (int i = 0; i < 2; ++i) {
        const int pb = phys_of_tile_bit(rB + bB + i);
        T.bq[i] = active(pb) ? qubit_of_pos[pb] : -1;
    }
    // every active bit of this type must be rotated by exactly one slot
    {
        int cover[64] = {0};
        for (int i = 0; i < 4; ++i) { cover[phys_of_tile_bit(rM + i)]++; cover[phys_of_tile_bit(rA + i)]++; }
        for (int i = 0; i < 2; ++i) cover[phys_of_tile_bit(rB + bB + i)]++;
        for (int pos = 0; pos < n; ++pos)
            if (active(pos) && cover[pos] != 1) return false;
    }
    for (int i = 0, t = 0; t < kTileBits; ++t)       // thread bits of round M: tile bits outside R, ascending
        if (t < rM || t >= rM + 4) tphys[i++] = phys_of_tile_bit(t);
    std::vector<int> colphys;
    for (int pos = 0; pos < n; ++pos)
        if (tile_bit_of_phys(pos) < 0) colphys.push_back(pos);
    if ((int)colphys.size() != n - kTileBits) return false;
    // the tile id enumerates the column bits in ascending physical order (tile_base): low part then high part
    auto local = [&](int pos, int& kind, int& idx) {
        for (int i = 0; i < 4; ++i) if (rphys[i] == pos) { kind = 0; idx = i; return; }
        for (int i = 0; i < 8; ++i) if (tphys[i] == pos) { kind = 1; idx = i; return; }
        for (size_t i = 0; i < colphys.size(); ++i) if (colphys[i] == pos) { kind = 2; idx = (int)i; return; }
        kind = -1; idx = -1;
    };
    if (p->n_zz > kMaxPairs) return false;
    T.n_pairs = p->n_zz;
    int fr_cnt[4] = {0, 0, 0, 0}, xt_cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int e = 0; e < p->n_zz; ++e) {
        int ka, ia, kb2, ib;
        local(p->pa[e], ka, ia);
        local(p->pb[e], kb2, ib);
        if (ka < 0 || kb2 < 0) return false;
        if (ka > kb2) { std::swap(ka, kb2); std::swap(ia, ib); }
        const int posb = (kb2 == 1) ? tphys[ib] : (kb2 == 2 ? colphys[ib] : rphys[ib]);
        if (ka == 0 && kb2 == 0) { T.cls[e] = 0; T.has_aj = 1; }
        else if (ka == 1 && kb2 == 1) T.cls[e] = 1;
        else if (ka == 0 && kb2 == 1) T.cls[e] = 2;
        else if (ka == 0 && kb2 == 2) T.cls[e] = 3;
        else if (ka == 1 && kb2 == 2) T.cls[e] = 4;
        else T.cls[e] = 5;
        T.i0[e] = (signed char)ia;
        T.i1[e] = (signed char)ib;
        if (ka == 0 && kb2 != 0) {              // neighbour of register bit ia outside R
            if (fr_cnt[ia] >= kMaxNbr) return false;
            const int t = fr_cnt[ia]++;
            T.g.fr_pos[ia][t] = posb; T.g.fr_msk[ia][t] = 1; T.fr_pair[ia][t] = e;
        }
        if (ka == 1 && kb2 == 2) {              // column neighbour of thread bit ia
            if (xt_cnt[ia] >= kMaxNbr) return false;
            const int t = xt_cnt[ia]++;
            T.g.xt_pos[ia][t] = posb; T.g.xt_msk[ia][t] = 1; T.xt_pair[ia][t] = e;
        }
    }
    return true;
}

template <typename K> static bool prep_kernel(K kern, size_t smem, int* occ) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, kThreads, smem) == cudaSuccess && *occ >= 1;
}

static Plan* get_plan(dq_ising* p) {
    Plan* pl = find_plan(p);
    if (pl) return pl;
    pl = new Plan();
    g_plans.push_back({p, pl});
    pl->n = p->n;
    if (p->ctx->set_device() != DQ_OK) return pl;
    if (p->n < kTileBits || p->n > 20) return pl;
    for (int t = 0; t < 2; ++t)
        if (!build_type(p, t, pl->types[t], pl->rphys[t])) return pl;
    pl->has_aj = pl->types[0].has_aj || pl->types[1].has_aj;
    pl->n_col_bits = p->n - kTileBits;
    pl->tiles_log2 = p->n - kTileBits;
    pl->smem_bytes = sizeof(c128) * kTile + 2 * sizeof(PassStep);
    if (pl->d_types.reserve(sizeof(TypePlan) * 2) != DQ_OK) return pl;
    if (cudaMemcpy(pl->d_types.p, pl->types, sizeof(TypePlan) * 2, cudaMemcpyHostToDevice) != cudaSuccess) return pl;
    int occ = 0, o2 = 0;
    bool good = pl->has_aj ? (prep_kernel(k_f16_passes<true, true>, pl->smem_bytes, &occ) &&
                              prep_kernel(k_f16_passes<false, true>, pl->smem_bytes, &o2))
                           : (prep_kernel(k_f16_passes<true, false>, pl->smem_bytes, &occ) &&
                              prep_kernel(k_f16_passes<false, false>, pl->smem_bytes, &o2));
    if (!good) { cudaGetLastError(); return pl; }
    pl->ctas_per_sm = std::min(occ, o2);
    pl->counter_slots = 1 << 16;
    if (pl->counters.reserve(pl->counter_slots * sizeof(unsigned)) != DQ_OK) return pl;
    pl->ok = true;
    return pl;
}

struct Traj {
    long long row0;
    int n_steps;
    int cls;
    bool final_energy;
    size_t step0;
};

static void add_traj(std::vector<SetupJob>& jobs, Traj& t) {
    t.step0 = jobs.size();
    for (int p = 0; p <= t.n_steps; ++p) {
        SetupJob j;
        j.row_pre = p >= 1 ? t.row0 + p - 1 : -1;
        j.row_cur = p < t.n_steps ? t.row0 + p : -1;
        j.type = (p + t.cls) & 1;
        j.flags = p < t.n_steps ? F_STORE : (t.final_energy ? F_ENERGY : F_STORE);
        jobs.push_back(j);
    }
}

static int run_setup(dq_ising* p, Plan* pl, const std::vector<SetupJob>& jobs, const double* d_rows, bool scaled) {
    cudaStream_t st = p->ctx->stream;
    DQ_TRY(pl->jobs.reserve(jobs.size() * sizeof(SetupJob)));
    DQ_TRY(pl->steps.reserve(jobs.size() * sizeof(PassStep)));
    DQ_TRY(pl->tc.reserve((jobs.size() << pl->n_col_bits) * sizeof(double2)));
    DQ_CUDA(cudaMemcpyAsync(pl->jobs.p, jobs.data(), jobs.size() * sizeof(SetupJob), cudaMemcpyHostToDevice, st));
    const unsigned ncol = 1u << pl->n_col_bits;
    dim3 grid((unsigned)jobs.size(), std::max(1u, std::min(8u, ncol / 256)));
    k_setup<<<grid, 256, 0, st>>>(pl->jobs.as<SetupJob>(), d_rows, p->row_len, p->n_zz, pl->d_types.as<TypePlan>(),
                                  pl->steps.as<PassStep>(), pl->tc.as<double2>(), pl->n_col_bits, scaled ? 1 : 0);
    p->ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

static int launch_group(dq_ising* p, Plan* pl, const KetDesc* d_kets, int n_kets, int max_pass, bool scaled, double r) {
    cudaStream_t st = p->ctx->stream;
    if (pl->counter_cursor + 1 + n_kets > pl->counter_slots) pl->counter_cursor = 0;
    unsigned* ctr = pl->counters.as<unsigned>() + pl->counter_cursor;
    pl->counter_cursor += 1 + n_kets;
    DQ_CUDA(cudaMemsetAsync(ctr, 0, (1 + n_kets) * sizeof(unsigned), st));
    LaunchArgs A;
    A.kets = d_kets;
    A.tc = pl->tc.as<double2>();
    A.mdiag = p->mdiag.as<double>();
    A.counters = ctr;
    A.n_kets = n_kets;
    A.max_pass = max_pass;
    A.tiles_log2 = pl->tiles_log2;
    const double alpha = atan(r);
    A.r = r; A.ca = cos(alpha); A.sa = sin(alpha); A.c2a = cos(2 * alpha); A.s2a = sin(2 * alpha);
    A.geom[0] = pl->types[0].g;
    A.geom[1] = pl->types[1].g;
    const long long all_items = ((long long)n_kets << pl->tiles_log2) * max_pass;
    long long grid = (long long)p->ctx->prop.multiProcessorCount * (p->grid_per_sm > 0 ? std::min(p->grid_per_sm, pl->ctas_per_sm) : pl->ctas_per_sm);
    if (grid > all_items) grid = all_items;
    const bool timed = p->time_launches != 0;
    if (timed) {
        while ((int)pl->ev.size() < pl->ev_used + 2) {
            cudaEvent_t e;
            DQ_CUDA(cudaEventCreate(&e));
            pl->ev.push_back(e);
        }
        DQ_CUDA(cudaEventRecord(pl->ev[pl->ev_used], st));
    }
    if (scaled) {
        if (pl->has_aj) k_f16_passes<true, true><<<(unsigned)grid, kThreads, pl->smem_bytes, st>>>(A);
        else k_f16_passes<true, false><<<(unsigned)grid, kThreads, pl->smem_bytes, st>>>(A);
    } else {
        if (pl->has_aj) k_f16_passes<false, true><<<(unsigned)grid, kThreads, pl->smem_bytes, st>>>(A);
        else k_f16_passes<false, false><<<(unsigned)grid, kThreads, pl->smem_bytes, st>>>(A);
    }
    p->ctx->launches++;
    if (timed) {
        DQ_CUDA(cudaEventRecord(pl->ev[pl->ev_used + 1], st));
        pl->ev_used += 2;
    }
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

static bool rows_allow_scaled(const dq_ising* p, const double* rows, long long n_rows, double r) {
    double mx = fabs(atan(r));
    for (long long k = 0; k < n_rows; ++k)
        for (int q = 0; q < p->n; ++q) mx = std::max(mx, fabs(rows[k * p->row_len + 1 + p->n_zz + q]));
    return mx <= 1.0;
}

static bool in_r(const Plan* pl, int type, int pos) {
    for (int i = 0; i < 4; ++i)
        if (pl->rphys[type][i] == pos) return true;
    return false;
}

}  // namespace f16

int f16_launch_times(dq_ising* p, double* total_ms, double* n_launches) {
    f16::Plan* pl = f16::find_plan(p);
    *total_ms = 0.0;
    *n_launches = 0.0;
    if (!pl || pl->ev_used == 0) return DQ_OK;
    DQ_CUDA(cudaStreamSynchronize(p->ctx->stream));
    for (int i = 0; i + 1 < pl->ev_used; i += 2) {
        float ms = 0.f;
        DQ_CUDA(cudaEventElapsedTime(&ms, pl->ev[i], pl->ev[i + 1]));
        *total_ms += ms;
        *n_launches += 1.0;
    }
    return DQ_OK;
}

int f16_supported(const dq_ising* p) {
    f16::Plan* pl = f16::get_plan(const_cast<dq_ising*>(p));
    return pl->ok ? 1 : 0;
}

void f16_release(dq_ising* p) {
    for (size_t i = 0; i < f16::g_plans.size(); ++i)
        if (f16::g_plans[i].first == p) {
            f16::Plan* pl = f16::g_plans[i].second;
            DevBuf* bufs[] = {&pl->d_types, &pl->jobs, &pl->steps, &pl->tc, &pl->kets, &pl->counters, &pl->partials,
                              &pl->work, &pl->rows, &pl->phi, &pl->uniform};
            for (auto* b : bufs) b->release();
            for (cudaEvent_t e : pl->ev) cudaEventDestroy(e);
            delete pl;
            f16::g_plans.erase(f16::g_plans.begin() + i);
            return;
        }
}

int f16_evolve(dq_ising* p, c128* d_states, int batch, const double* h_rows, int n_steps, double* d_energies,
               bool want_states) {
    using namespace f16;
    Plan* pl = get_plan(p);
    DQ_REQUIRE(pl->ok, "f16 engine unavailable for this problem");
    pl->ev_used = 0;
    cudaStream_t st = p->ctx->stream;
    const bool scaled = rows_allow_scaled(p, h_rows, n_steps, 0.0);
    const size_t N = p->dim();
    const int tiles = 1 << pl->tiles_log2;
    DQ_TRY(pl->rows.reserve((size_t)std::max(1, n_steps) * p->row_len * sizeof(double)));
    if (n_steps)
        DQ_CUDA(cudaMemcpyAsync(pl->rows.p, h_rows, (size_t)n_steps * p->row_len * sizeof(double), cudaMemcpyHostToDevice, st));
    std::vector<SetupJob> jobs;
    Traj tr{0, n_steps, 0, !want_states, 0};
    add_traj(jobs, tr);
    DQ_TRY(run_setup(p, pl, jobs, pl->rows.as<double>(), scaled));
    std::vector<KetDesc> kets(batch);
    DQ_TRY(pl->partials.reserve((size_t)batch * tiles * sizeof(double)));
    for (int g = 0; g < batch; ++g) {
        KetDesc& k = kets[g];
        k.src = d_states + (size_t)g * N;
        k.buf = d_states + (size_t)g * N;
        k.steps = pl->steps.as<PassStep>();
        k.partial = pl->partials.as<double>() + (size_t)g * tiles;
        k.sigma = 0.0;
        k.escale = 1.0;
        k.n_pass = n_steps + 1;
        k.shift_kind = -1;
        k.sb0 = k.sb1 = 0;
        k.cls = 0;
        k.pad_ = 0;
    }
    DQ_TRY(pl->kets.reserve(kets.size() * sizeof(KetDesc)));
    DQ_CUDA(cudaMemcpyAsync(pl->kets.p, kets.data(), kets.size() * sizeof(KetDesc), cudaMemcpyHostToDevice, st));
    const int G = auto_ket_group(p);
    for (int g0 = 0; g0 < batch; g0 += G)
        DQ_TRY(launch_group(p, pl, pl->kets.as<KetDesc>() + g0, std::min(G, batch - g0), n_steps + 1, scaled, 0.5));
    if (d_energies) {
        if (want_states) {
            DQ_TRY(gen_energy(p, d_states, batch, d_energies));
        } else {
            k_sum_partials<<<batch, 32, 0, st>>>(pl->partials.as<double>(), tiles, d_energies);
            p->ctx->launches++;
        }
    }
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

int f16_grad_run(dq_ising* p) {
    using namespace f16;
    Plan* pl = get_plan(p);
    DQ_REQUIRE(pl->ok, "f16 engine unavailable for this problem");
    pl->ev_used = 0;
    auto& s = p->st;
    cudaStream_t st = p->ctx->stream;
    const size_t N = p->dim();
    const int tiles = 1 << pl->tiles_log2;
    const int B = s.n_samples, n_shift = s.n_shift, kets_per = 2 * n_shift;
    const int G = auto_ket_group(p);
    const bool scaled = s.scaled_ok;

    const long long np = s.prefix_off[B], ns = s.suffix_off[B];
    DQ_TRY(pl->rows.reserve((size_t)std::max<long long>(1, np + ns) * p->row_len * sizeof(double)));
    if (np) DQ_CUDA(cudaMemcpyAsync(pl->rows.p, p->rows_a.p, np * p->row_len * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (ns) DQ_CUDA(cudaMemcpyAsync(pl->rows.as<double>() + np * p->row_len, p->rows_b.p, ns * p->row_len * sizeof(double), cudaMemcpyDeviceToDevice, st));
    std::vector<SetupJob> jobs;
    std::vector<Traj> pre(B), sufL(B), sufH(B);
    for (int b = 0; b < B; ++b) {
        pre[b] = Traj{s.prefix_off[b], s.prefix_steps[b], 0, false, 0};
        sufL[b] = Traj{np + s.suffix_off[b], s.suffix_steps[b], 0, true, 0};
        sufH[b] = Traj{np + s.suffix_off[b], s.suffix_steps[b], 1, true, 0};
        add_traj(jobs, pre[b]);
        add_traj(jobs, sufL[b]);
        add_traj(jobs, sufH[b]);
    }
    DQ_TRY(run_setup(p, pl, jobs, pl->rows.as<double>(), scaled));

    DQ_TRY(pl->phi.reserve((size_t)B * N * sizeof(c128)));
    DQ_TRY(pl->work.reserve((size_t)G * N * sizeof(c128)));
    DQ_TRY(pl->partials.reserve((size_t)B * kets_per * tiles * sizeof(double)));
    const c128* psi0 = nullptr;
    if (s.uniform_psi0) {
        DQ_TRY(pl->uniform.reserve(N * sizeof(c128)));
        DQ_TRY(gen_fill_uniform(p, pl->uniform.as<c128>(), 1));
        psi0 = pl->uniform.as<c128>();
    } else {
        psi0 = s.psi0.as<c128>();
    }

    const double esc_x = scaled ? 1.0 / (1.0 + s.r * s.r) : 1.0;
    std::vector<KetDesc> kets;
    kets.reserve((size_t)B * (kets_per + 1));
    for (int b = 0; b < B; ++b) {
        KetDesc k;
        k.src = psi0;
        k.buf = pl->phi.as<c128>() + (size_t)b * N;
        k.steps = pl->steps.as<PassStep>() + pre[b].step0;
        k.partial = nullptr;
        k.sigma = 0.0;
        k.escale = 1.0;
        k.n_pass = pre[b].n_steps + 1;
        k.shift_kind = -1;
        k.sb0 = k.sb1 = 0;
        k.cls = 0;
        k.pad_ = 0;
        kets.push_back(k);
    }
    struct Group { size_t first; int count; int max_pass; };
    std::vector<Group> groups;
    for (int g0 = 0; g0 < B; g0 += G) {
        int cnt = std::min(G, B - g0), mp = 0;
        for (int g = 0; g < cnt; ++g) mp = std::max(mp, kets[g0 + g].n_pass);
        groups.push_back({(size_t)g0, cnt, mp});
    }
    for (int b = 0; b < B; ++b) {
        for (int cls = 0; cls < 2; ++cls) {
            std::vector<KetDesc> mine;
            for (int i = 0; i < n_shift; ++i) {
                int kcls = 0, b0 = 0, b1 = 0;
                if (s.shift_kind[i] == 0) {
                    // a ZZ shift rides on the phase of pass 0: start with the type in which the pair is not R-R
                    b0 = p->pa[s.shift_index[i]]; b1 = p->pb[s.shift_index[i]];
                    kcls = (in_r(pl, 0, b0) && in_r(pl, 0, b1)) ? 1 : 0;
                    if (kcls == 1 && in_r(pl, 1, b0) && in_r(pl, 1, b1)) {
                        set_error("f16 engine: ZZ shift pair (%d,%d) is R-R in both pass types", b0, b1);
                        return DQ_ERR_UNSUPPORTED;
                    }
                } else {
                    b0 = p->bitpos[s.shift_index[i]];
                    kcls = b0 >= 10 ? 1 : 0;
                }
                if (kcls != cls) continue;
                for (int sg = 0; sg < 2; ++sg) {
                    KetDesc k;
                    k.src = pl->phi.as<c128>() + (size_t)b * N;
                    k.buf = nullptr;
                    k.steps = pl->steps.as<PassStep>() + (cls == 0 ? sufL[b].step0 : sufH[b].step0);
                    const size_t kidx = (size_t)b * kets_per + 2 * i + sg;
                    k.partial = pl->partials.as<double>() + kidx * tiles;
                    k.sigma = sg == 0 ? +1.0 : -1.0;
                    k.escale = s.shift_kind[i] == 1 ? esc_x : 1.0;
                    k.n_pass = s.suffix_steps[b] + 1;
                    k.shift_kind = s.shift_kind[i];
                    k.sb0 = b0;
                    k.sb1 = b1;
                    k.cls = cls;
                    k.pad_ = 0;
                    mine.push_back(k);
                }
            }
            for (size_t g0 = 0; g0 < mine.size(); g0 += G) {
                const int cnt = (int)std::min<size_t>(G, mine.size() - g0);
                for (int g = 0; g < cnt; ++g) mine[g0 + g].buf = pl->work.as<c128>() + (size_t)g * N;
                groups.push_back({kets.size() + g0, cnt, s.suffix_steps[b] + 1});
            }
            kets.insert(kets.end(), mine.begin(), mine.end());
        }
    }
    DQ_TRY(pl->kets.reserve(kets.size() * sizeof(KetDesc)));
    DQ_CUDA(cudaMemcpyAsync(pl->kets.p, kets.data(), kets.size() * sizeof(KetDesc), cudaMemcpyHostToDevice, st));

    for (const Group& g : groups)
        DQ_TRY(launch_group(p, pl, pl->kets.as<KetDesc>() + g.first, g.count, g.max_pass, scaled, s.r));

    k_sum_partials<<<B * kets_per, 32, 0, st>>>(pl->partials.as<double>(), tiles, p->energies.as<double>());
    p->ctx->launches++;
    DQ_CUDA(cudaGetLastError());
    return DQ_OK;
}

}  // namespace dq
